// comm.cu -- frame-range shards of one clip over the GPUs of a box, inside the C ABI (no PyTorch, no MPI).
//
// The reference is single-adapter (dips/src/gpu/mod.rs:71-78 picks one wgpu adapter); the north star shards a long clip by
// frame range over the 8 GPUs of one box.  Rank r of R owns frames [r*N/R, (r+1)*N/R) (dipsb_shard_range).  Per clip:
//   * overall mode:   rank 0 builds the u16 reference plane from frame 0 (2 B/px, less than the raw frame) and it travels as
//                     scatter + all-gather over peer memory: rank 0's prime kernel stores slice k straight into rank k+1,
//                     the other ranks forward their slices to each other -- the only exchange before the pass;
//   * per-frame mode: NO exchange before the pass.  Every rank starts at once, primed from its own first frame (whose
//                     difference is therefore missing); meanwhile its copy engine pushes that first frame over NVLink
//                     into the previous rank's window, and the previous rank's clip kernel differences it as one extra
//                     trailing frame (its producer warp waits for the arrival stamp right before the last TMA fetch).
//                     The one-frame halo thus never sits on the critical path; the boundary frame's scalars travel back
//                     with the accumulator exchange;
//   * at the end:     the accumulators are combined by a reduce-scatter over peer memory: the clip kernel's last flush packs
//                     the elements this rank does not own (sum | count << bits in one u32 while a shard's totals fit, else
//                     two u32) and stores them straight into the owner's window over NVLink (xchg_push_kernel does the same
//                     as a separate launch when the clip kernel cannot), one block stamps the peers, and xchg_reduce_kernel
//                     waits for the peers' stamps and adds up the pixel range this rank owns; the totals stay sharded by pixel
//                     range until somebody reads them (dipsb_gather_accumulators: the same pattern as an all-gather).
//                     Integer sums: bit-exact, order independent.
// Fallback and comparison path (DIPSB_REDUCE_NCCL): ncclBroadcast of the plane, pack -> ncclAllReduce -> unpack, and ncclSend/ncclRecv for the halo,
// when peer memory cannot be mapped (cudaIpc* refused, no P2P) or on request.
//
// Two process models share all of this:
//   * one process per GPU (torchrun, MPI, ...): dipsb_comm_unique_id + dipsb_comm_init_rank; windows are mapped with
//     cudaIpcGetMemHandle / cudaIpcOpenMemHandle, the handles travel through an ncclAllGather;
//   * one process, all GPUs (a Rust host): dipsb_create_group -> ncclCommInitAll + cudaDeviceEnablePeerAccess.
// NCCL is loaded with dlopen("libnccl.so.2") when the first communicator is made, so the library has no link-time
// dependency on it and single-GPU users never need it.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "dipsb_ctx.h"
#include "peer_sync.cuh"

using namespace dipsb;

namespace dipsb {

constexpr uint32_t kXchgBlocks = 592;      // 4 blocks of 256 threads per SM
constexpr uint32_t kXchgThreads = 256;
constexpr size_t kCtrlBytes = 4096;

// ---- NCCL, loaded on demand ------------------------------------------------------------------------------------------
struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    int version = 0;
};

static NcclApi* nccl_api(std::string* why) {
    static NcclApi api;
    static std::string err;
    static bool tried = false;
    if (!tried) {
        tried = true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) err = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "?");
        else {
#define DIPSB_NCCL_SYM(field, sym)                                                  \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, #sym));      \
    if (!api.field && err.empty()) err = "libnccl has no " #sym;
            DIPSB_NCCL_SYM(GetUniqueId, ncclGetUniqueId)
            DIPSB_NCCL_SYM(CommInitRank, ncclCommInitRank)
            DIPSB_NCCL_SYM(CommInitAll, ncclCommInitAll)
            DIPSB_NCCL_SYM(CommDestroy, ncclCommDestroy)
            DIPSB_NCCL_SYM(Broadcast, ncclBroadcast)
            DIPSB_NCCL_SYM(AllReduce, ncclAllReduce)
            DIPSB_NCCL_SYM(AllGather, ncclAllGather)
            DIPSB_NCCL_SYM(Send, ncclSend)
            DIPSB_NCCL_SYM(Recv, ncclRecv)
            DIPSB_NCCL_SYM(GroupStart, ncclGroupStart)
            DIPSB_NCCL_SYM(GroupEnd, ncclGroupEnd)
            DIPSB_NCCL_SYM(GetErrorString, ncclGetErrorString)
            DIPSB_NCCL_SYM(GetVersion, ncclGetVersion)
#undef DIPSB_NCCL_SYM
            if (err.empty()) api.GetVersion(&api.version);
        }
    }
    if (!err.empty()) {
        if (why) *why = err;
        return nullptr;
    }
    return &api;
}

#define NK(c, api, call)                                                                                          \
    do {                                                                                                          \
        ncclResult_t r__ = (call);                                                                                \
        if (r__ != ncclSuccess)                                                                                   \
            return fail((c), DIPSB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, (api)->GetErrorString(r__), __FILE__, __LINE__); \
    } while (0)

// ---- the window: per-rank device memory its peers write into -----------------------------------------------------------
//   [0, 4096)                      control: counters, stamps, boundary-scalar mailboxes, status (struct Control)
//   [4096, + 2*halo_bytes)         halo frame buffers, one per epoch parity (per-frame mode; zero padded to 16 bytes)
//   then 2 * nranks * slot_bytes   receive slots of the accumulator exchange: [parity][source rank][chunk * 8 bytes]
struct Control {
    unsigned long long xchg_arrived[kMaxRanks];     // per source rank: the last pass whose partial sums it has pushed here (monotonic)
    unsigned long long gather_arrived[kMaxRanks];   // same for the all-gather of the totals
    unsigned long long halo_stamp[2];     // epoch whose halo frame sits in halo buffer [parity]
    unsigned long long plane_stamp;       // epoch whose reference plane (my slice of it) rank 0 has pushed into my state plane
    unsigned long long plane_stamp2[kMaxRanks];   // per source rank: epoch whose slice of the reference plane it forwarded here
    unsigned long long my_stamp[2];       // source words of the stamp copies this rank sends (written by stamp_kernel)
    unsigned long long mbox[2][2];        // [parity]{sad, cnt} of my first frame, computed by the previous rank
    uint32_t status;                      // != 0: a bounded wait timed out
    uint32_t blocks_done[4];              // local: finished blocks of the running exchange / gather / plane scatter / plane forward kernel
};
static_assert(sizeof(Control) <= kCtrlBytes, "control block");

struct Comm {
    int nranks = 1, rank = 0;
    bool single_process = false;
    bool loopback = false;                // several ranks on one device (test mode): clip kernels are serialised before the exchange
    NcclApi* nccl = nullptr;
    ncclComm_t comm = nullptr;
    bool p2p = false;                     // peers' windows are mapped
    int reduce_path = 0;                  // 0 automatic, 1 peer-memory kernel, 2 NCCL all-reduce
    uint8_t* win = nullptr;               // this rank's window
    size_t win_bytes = 0, halo_bytes = 0, slot_bytes = 0;
    uint64_t chunk = 0;                   // accumulator elements owned per rank (multiple of 4; the last rank's may be short)
    uint8_t* win_peer[kMaxRanks] = {};    // every rank's window in this process's address space ([rank] == win)
    uint32_t* acc_peer[kMaxRanks] = {};
    uint16_t* state_peer[kMaxRanks][2] = {};
    bool opened[kMaxRanks][4] = {};       // IPC mappings to close
    uint64_t epoch = 0;                   // sharded passes so far
    uint64_t gathers = 0;
    uint64_t timeout_ns = 5000000000ull;
    cudaEvent_t ev_start = nullptr, ev_halo[2] = {nullptr, nullptr};
    bool halo_used[2] = {false, false};
    uint8_t* halo_local = nullptr;        // NCCL-only halo path: receive buffer (no window)
    bool fuse_push = true;                // let the clip kernel's last flush push the partial sums (DIPSB_COMM_FUSE_PUSH=0: separate kernel)
    struct dipsb_group* group = nullptr;
    void* packed = nullptr; uint64_t packed_words = 0;   // NCCL path: the packed exchange buffer between its phases
};

}  // namespace dipsb

struct dipsb_group {
    std::vector<dipsb_ctx*> ctx;
    NcclApi* nccl = nullptr;
    bool loopback = false;
    std::string err;
};

namespace {

__global__ void wait_flag_kernel(const unsigned long long* flag, unsigned long long want, unsigned long long timeout_ns,
                                 uint32_t* status) {
    spin_until(flag, want, timeout_ns, status);
}
__global__ void stamp_kernel(unsigned long long* word, unsigned long long value) { *word = value; }
// ---- the accumulator exchange ----------------------------------------------------------------------------------------
struct XchgParams {
    uint32_t* acc;                // local planes: sum[n_elems] then cnt[n_elems], internal tile order
    uint64_t n_elems, chunk;      // chunk: elements owned per rank (multiple of 4)
    uint32_t nranks, rank;
    int fmt;                      // 1: one u32 per element (sum | cnt << sum_bits); 2: sum and cnt as two u32
    int sum_bits;
    uint64_t slot_bytes;          // bytes between two source ranks' slots in a receive area
    const uint8_t* recv_local;    // my receive area of this parity
    uint8_t* recv_peer[kMaxRanks];
    unsigned long long* arrived_local;        // my xchg_arrived[]: one stamp per source rank
    unsigned long long* stamp_peer[kMaxRanks]; // peer r's xchg_arrived[my rank] (null for myself)
    unsigned long long target, timeout_ns;
    uint32_t* status;
    uint32_t* blocks_done;
    // per-frame boundary scalars: mine for the next rank's first frame go into its mailbox; my own first frame's come in
    const unsigned long long* sad_src; const unsigned long long* cnt_src; unsigned long long* mbox_next;
    const unsigned long long* mbox_local; unsigned long long* sad_dst; unsigned long long* cnt_dst;
};

__device__ __forceinline__ uint4 ld4(const uint32_t* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void st4(uint32_t* p, uint4 v) { *reinterpret_cast<uint4*>(p) = v; }
__device__ __forceinline__ uint4 add4(uint4 a, uint4 b) { return make_uint4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// Reduce-scatter over peer memory, two launches.
// xchg_push_kernel: every element this rank does not own is packed and stored into its owner's receive slot (remote stores
// over NVLink, 16 bytes per thread, 512 contiguous bytes per warp); its last block then stamps this rank's slot in every
// peer's window with the pass number.  It never waits, so it always completes -- whatever else is resident on the GPU.
// xchg_reduce_kernel: waits until every peer's stamp has reached this pass (one stamp per source rank, monotonic over
// the passes: a fast peer's next pass cannot stand in for a slow peer's current one), then adds the N-1 received partials
// of the owned range to the local ones, in place.
__global__ void __launch_bounds__(kXchgThreads) xchg_push_kernel(const XchgParams P) {
    const uint32_t* sum = P.acc;
    const uint32_t* cnt = P.acc + P.n_elems;
    if (blockIdx.x == 0 && threadIdx.x == 0 && P.mbox_next) {
        P.mbox_next[0] = *P.sad_src;
        P.mbox_next[1] = *P.cnt_src;
    }
    // Owners are visited in the order rank+1, rank+2, ...: at any moment every rank sends to a different owner, so no GPU's
    // NVLink ingress takes the traffic of all the others at once.  Two 16-byte pieces per thread and step are in flight.
    const uint64_t chunk_units = P.chunk / 4, total_units = P.n_elems / 4;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, t0 = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    for (uint32_t k = 1; k < P.nranks; ++k) {
        const uint32_t owner = (P.rank + k) % P.nranks;
        const uint64_t lo = (uint64_t)owner * chunk_units, hi = min(lo + chunk_units, total_units);
        uint32_t* slot = reinterpret_cast<uint32_t*>(P.recv_peer[owner] + (uint64_t)P.rank * P.slot_bytes);
        for (uint64_t u = lo + t0; u < hi; u += 2 * stride) {
            const uint64_t u2 = u + stride;
            const bool two = u2 < hi;
            const uint4 s0 = ld4(sum + 4 * u), c0 = ld4(cnt + 4 * u);
            uint4 s1 = s0, c1 = c0;
            if (two) { s1 = ld4(sum + 4 * u2); c1 = ld4(cnt + 4 * u2); }
            const uint64_t k0 = 4 * (u - lo), k1 = 4 * (u2 - lo);
            if (P.fmt == 1) {
                st4(slot + k0, make_uint4(s0.x | (c0.x << P.sum_bits), s0.y | (c0.y << P.sum_bits), s0.z | (c0.z << P.sum_bits), s0.w | (c0.w << P.sum_bits)));
                if (two) st4(slot + k1, make_uint4(s1.x | (c1.x << P.sum_bits), s1.y | (c1.y << P.sum_bits), s1.z | (c1.z << P.sum_bits), s1.w | (c1.w << P.sum_bits)));
            } else {
                st4(slot + k0, s0);
                st4(slot + P.chunk + k0, c0);
                if (two) { st4(slot + k1, s1); st4(slot + P.chunk + k1, c1); }
            }
        }
    }
    stamp_when_last(P.blocks_done, P.stamp_peer, P.nranks, P.target);
}

__global__ void __launch_bounds__(kXchgThreads) xchg_reduce_kernel(const XchgParams P) {
    const uint64_t own_lo = (uint64_t)P.rank * P.chunk, own_hi = min(own_lo + P.chunk, P.n_elems);
    int mine_ok = 1;
    if (threadIdx.x < P.nranks && threadIdx.x != P.rank)
        mine_ok = spin_until(P.arrived_local + threadIdx.x, P.target, P.timeout_ns, P.status) ? 1 : 0;
    if (!__syncthreads_and(mine_ok)) return;
    if (blockIdx.x == 0 && threadIdx.x == 0 && P.mbox_local) {
        *P.sad_dst = P.mbox_local[0];
        *P.cnt_dst = P.mbox_local[1];
    }
    uint32_t* osum = P.acc;
    uint32_t* ocnt = P.acc + P.n_elems;
    const uint32_t mask = P.sum_bits >= 32 ? 0xFFFFFFFFu : ((1u << P.sum_bits) - 1u);
    for (uint64_t u = own_lo / 4 + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; u < own_hi / 4; u += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = 4 * u, k = i - own_lo;
        uint4 s = ld4(osum + i), c = ld4(ocnt + i);
        for (uint32_t r = 0; r < P.nranks; ++r) {
            if (r == P.rank) continue;
            const uint32_t* slot = reinterpret_cast<const uint32_t*>(P.recv_local + (uint64_t)r * P.slot_bytes);
            if (P.fmt == 1) {
                const uint4 w = ld4(slot + k);
                s = add4(s, make_uint4(w.x & mask, w.y & mask, w.z & mask, w.w & mask));
                c = add4(c, make_uint4(w.x >> P.sum_bits, w.y >> P.sum_bits, w.z >> P.sum_bits, w.w >> P.sum_bits));
            } else {
                s = add4(s, ld4(slot + k));
                c = add4(c, ld4(slot + P.chunk + k));
            }
        }
        st4(osum + i, s);
        st4(ocnt + i, c);
    }
}

// All-gather of the totals: every rank stores its owned range into the accumulator planes of all its peers, raises its
// counter there, and a one-warp kernel behind it holds the stream until every peer's range has arrived here.
struct GatherParams {
    uint32_t* acc; uint64_t n_elems, chunk; uint32_t nranks, rank;
    uint32_t* acc_peer[kMaxRanks];
    unsigned long long* arrived_local; unsigned long long* stamp_peer[kMaxRanks];
    unsigned long long target, timeout_ns; uint32_t* status; uint32_t* blocks_done;
};
__global__ void __launch_bounds__(kXchgThreads) gather_push_kernel(const GatherParams P) {
    const uint64_t own_lo = (uint64_t)P.rank * P.chunk, own_hi = min(own_lo + P.chunk, P.n_elems);
    for (uint64_t u = own_lo / 4 + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; u < own_hi / 4; u += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = 4 * u;
        const uint4 s = ld4(P.acc + i), c = ld4(P.acc + P.n_elems + i);
        for (uint32_t k = 1; k < P.nranks; ++k) {
            const uint32_t r = (P.rank + k) % P.nranks;
            st4(P.acc_peer[r] + i, s);
            st4(P.acc_peer[r] + P.n_elems + i, c);
        }
    }
    stamp_when_last(P.blocks_done, P.stamp_peer, P.nranks, P.target);
}

// Reference-plane broadcast over peer memory: ranges of the local u16 plane go to peers' planes.  Scatter (rank 0 without the
// fused prime kernel): slice d -> rank d.  Forward (ranks > 0): wait for rank 0's stamp, then my slice -> every rank but 0 and
// me.  Then the last block stamps the receivers.
struct PlanePush {
    const uint16_t* plane; uint16_t* dst[kMaxRanks];
    uint64_t lo[kMaxRanks], hi[kMaxRanks];      // per destination: range in 16-byte units (lo == hi: nothing)
    uint32_t nranks, rank;
    const unsigned long long* wait_flag; unsigned long long wait_value, timeout_ns; uint32_t* status;
    unsigned long long* stamp_peer[kMaxRanks]; unsigned long long stamp; uint32_t* blocks_done;
};
__global__ void __launch_bounds__(kXchgThreads) plane_push_kernel(const PlanePush P) {
    if (P.wait_flag) {
        __shared__ int ok;
        if (threadIdx.x == 0) ok = spin_until(P.wait_flag, P.wait_value, P.timeout_ns, P.status) ? 1 : 0;
        __syncthreads();
        if (!ok) return;
    }
    const uint4* src = reinterpret_cast<const uint4*>(P.plane);
    for (uint32_t k = 1; k < P.nranks; ++k) {
        const uint32_t d = (P.rank + k) % P.nranks;
        uint4* dst = reinterpret_cast<uint4*>(P.dst[d]);
        for (uint64_t u = P.lo[d] + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; u < P.hi[d]; u += (uint64_t)gridDim.x * blockDim.x)
            dst[u] = src[u];
    }
    stamp_when_last(P.blocks_done, P.stamp_peer, P.nranks, P.stamp);
}
// one warp holds the stream until the stamps of the source ranks in `mask` have reached `target`
__global__ void wait_sources_kernel(const unsigned long long* arrived, uint32_t mask, unsigned long long target,
                                    unsigned long long timeout_ns, uint32_t* status) {
    if ((mask >> threadIdx.x) & 1u) spin_until(arrived + threadIdx.x, target, timeout_ns, status);
}

Control* ctrl_of(uint8_t* win) { return reinterpret_cast<Control*>(win); }

uint64_t timeout_from_env() {
    if (const char* e = getenv("DIPSB_COMM_TIMEOUT_MS")) {
        const long v = strtol(e, nullptr, 10);
        if (v >= 1 && v <= 600000) return (uint64_t)v * 1000000ull;
    }
    return 5000000000ull;
}

}  // namespace

cudaError_t dipsb::launch_wait_flag(const unsigned long long* flag, unsigned long long want, unsigned long long timeout_ns,
                                    uint32_t* status, cudaStream_t s) {
    wait_flag_kernel<<<1, 1, 0, s>>>(flag, want, timeout_ns, status);
    count_launch();
    return cudaGetLastError();
}

// ---- layout ----------------------------------------------------------------------------------------------------------
static uint64_t chunk_of(uint64_t n_elems, int nranks) {
    const uint64_t units = n_elems / 4;                                   // n_elems is a multiple of 512
    return 4 * ((units + (uint64_t)nranks - 1) / (uint64_t)nranks);
}

extern "C" void dipsb_shard_range(uint64_t total_frames, uint32_t nranks, uint32_t rank, uint64_t* first, uint64_t* count) {
    if (!nranks || rank >= nranks) {
        if (first) *first = 0;
        if (count) *count = 0;
        return;
    }
    const uint64_t t0 = total_frames / nranks * rank + (total_frames % nranks) * rank / nranks;
    const uint64_t t1 = total_frames / nranks * (rank + 1) + (total_frames % nranks) * (rank + 1) / nranks;
    if (first) *first = t0;      // == floor(rank * total / nranks) without the 64-bit overflow
    if (count) *count = t1 - t0;
}

// the exchange format of the peer-memory reduce for a clip of total_frames over nranks ranks:
// out[0] = bytes per element (4 or 8), out[1] = sum bits, out[2] = per-rank frame bound, out[3] = elements owned per rank
extern "C" int32_t dipsb_xchg_plan_query(uint64_t total_frames, uint32_t nranks, uint64_t n_elems, uint64_t out[4]) {
    if (!out || !nranks || nranks > (uint32_t)kMaxRanks || !total_frames) return DIPSB_ERR_INVALID;
    // a rank differences at most ceil(total/nranks) frames of its own plus, in per-frame mode, its successor's first frame
    const uint64_t bound = (total_frames + nranks - 1) / nranks + 1;
    const int sum_bits = bit_length(510ull * bound), cnt_bits = bit_length(bound);
    out[0] = (sum_bits + cnt_bits <= 32) ? 4 : 8;
    out[1] = (uint64_t)sum_bits;
    out[2] = bound;
    out[3] = chunk_of(n_elems ? n_elems : 512, (int)nranks);
    return DIPSB_OK;
}

// ---- communicator set-up ---------------------------------------------------------------------------------------------
static int32_t alloc_window(dipsb_ctx* c, Comm* m) {
    const Geometry& g = c->g;
    m->chunk = chunk_of(g.n_elems, m->nranks);
    m->halo_bytes = (size_t)((g.npx * g.bpp + 15) & ~15ull);
    m->slot_bytes = (size_t)(m->chunk * 8);
    m->win_bytes = kCtrlBytes + 2 * m->halo_bytes + 2 * (size_t)m->nranks * m->slot_bytes;
    CK(c, cudaMalloc(&m->win, m->win_bytes));
    // control block and halo buffers start zeroed (the halo padding past a frame must stay zero: the clip kernel's last
    // bulk copy of an odd-sized frame is rounded up into it)
    CK(c, cudaMemsetAsync(m->win, 0, kCtrlBytes + 2 * m->halo_bytes, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    CK(c, cudaEventCreateWithFlags(&m->ev_start, cudaEventDisableTiming));
    for (int k = 0; k < 2; ++k) CK(c, cudaEventCreateWithFlags(&m->ev_halo[k], cudaEventDisableTiming));
    m->timeout_ns = timeout_from_env();
    if (const char* e = getenv("DIPSB_COMM_FUSE_PUSH")) m->fuse_push = atoi(e) != 0;
    return DIPSB_OK;
}

static uint8_t* halo_buf(Comm* m, uint8_t* win, int parity) { return win + kCtrlBytes + (size_t)parity * m->halo_bytes; }
static uint8_t* recv_area(Comm* m, uint8_t* win, int parity) {
    return win + kCtrlBytes + 2 * m->halo_bytes + (size_t)parity * m->nranks * m->slot_bytes;
}

void dipsb::comm_detach(dipsb_ctx* c) {
    Comm* m = c->comm;
    if (!m) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->copy_stream);
    for (int r = 0; r < m->nranks; ++r) {
        if (m->opened[r][0]) cudaIpcCloseMemHandle(m->win_peer[r]);
        if (m->opened[r][1]) cudaIpcCloseMemHandle(m->acc_peer[r]);
        if (m->opened[r][2]) cudaIpcCloseMemHandle(m->state_peer[r][0]);
        if (m->opened[r][3]) cudaIpcCloseMemHandle(m->state_peer[r][1]);
    }
    if (m->comm && m->nccl) m->nccl->CommDestroy(m->comm);
    if (m->win) cudaFree(m->win);
    if (m->halo_local) cudaFree(m->halo_local);
    if (m->ev_start) cudaEventDestroy(m->ev_start);
    for (auto& e : m->ev_halo) if (e) cudaEventDestroy(e);
    delete m;
    c->comm = nullptr;
    c->acc_sharded = false;
}

extern "C" int32_t dipsb_comm_unique_id(void* id128) {
    if (!id128) return DIPSB_ERR_INVALID;
    std::string why;
    NcclApi* api = nccl_api(&why);
    if (!api) return fail(nullptr, DIPSB_ERR_CUDA, "comm_unique_id: %s", why.c_str());
    static_assert(sizeof(ncclUniqueId) == DIPSB_UNIQUE_ID_BYTES, "ncclUniqueId size");
    NK(nullptr, api, api->GetUniqueId(reinterpret_cast<ncclUniqueId*>(id128)));
    return DIPSB_OK;
}

struct RankCard {                      // what the ranks tell each other at init (all-gathered through NCCL)
    cudaIpcMemHandle_t win, acc, state0, state1;
    uint64_t n_elems, npx, win_bytes;
    int32_t device, p2p_wanted, pad[2];
};

extern "C" int32_t dipsb_comm_init_rank(dipsb_ctx* c, uint32_t nranks, uint32_t rank, const void* id128) {
    if (!c || !id128 || nranks == 0 || nranks > (uint32_t)kMaxRanks || rank >= nranks) return DIPSB_ERR_INVALID;
    if (c->comm) return fail(c, DIPSB_ERR_STATE, "comm_init_rank: the context already has a communicator");
    if (c->cfg.flavor != DIPSB_FLAVOR_FRAME0 || c->cfg.spatial_window > 1)
        return fail(c, DIPSB_ERR_INVALID, "comm_init_rank: sharded passes need the FRAME0 flavour and spatial_window 1");
    CK(c, cudaSetDevice(c->device));
    std::string why;
    NcclApi* api = nccl_api(&why);
    if (!api) return fail(c, DIPSB_ERR_CUDA, "comm_init_rank: %s", why.c_str());
    Comm* m = new (std::nothrow) Comm();
    if (!m) return fail(c, DIPSB_ERR_NOMEM, "comm_init_rank: out of host memory");
    m->nranks = (int)nranks; m->rank = (int)rank; m->nccl = api;
    c->comm = m;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    ncclResult_t nr = api->CommInitRank(&m->comm, (int)nranks, id, (int)rank);
    if (nr != ncclSuccess) {
        m->comm = nullptr;
        comm_detach(c);
        return fail(c, DIPSB_ERR_CUDA, "comm_init_rank: ncclCommInitRank: %s", api->GetErrorString(nr));
    }
    int32_t rc = alloc_window(c, m);
    if (rc) { comm_detach(c); return rc; }
    m->win_peer[rank] = m->win; m->acc_peer[rank] = c->acc;
    m->state_peer[rank][0] = c->state[0]; m->state_peer[rank][1] = c->state[1];
    if (nranks == 1) return DIPSB_OK;

    // exchange the cards (IPC handles of window / accumulators / state planes) through the communicator itself
    const char* env = getenv("DIPSB_COMM_P2P");
    RankCard mine{};
    mine.n_elems = c->g.n_elems; mine.npx = c->g.npx; mine.win_bytes = m->win_bytes; mine.device = c->device;
    mine.p2p_wanted = (env && atoi(env) == 0) ? 0 : 1;
    if (mine.p2p_wanted) {
        if (cudaIpcGetMemHandle(&mine.win, m->win) != cudaSuccess || cudaIpcGetMemHandle(&mine.acc, c->acc) != cudaSuccess ||
            cudaIpcGetMemHandle(&mine.state0, c->state[0]) != cudaSuccess || cudaIpcGetMemHandle(&mine.state1, c->state[1]) != cudaSuccess) {
            cudaGetLastError();
            mine.p2p_wanted = 0;
        }
    }
    RankCard* d_cards = nullptr;
    std::vector<RankCard> cards(nranks);
    cudaError_t e = cudaMalloc(&d_cards, nranks * sizeof(RankCard));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_cards + rank, &mine, sizeof mine, cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) { comm_detach(c); return fail(c, DIPSB_ERR_CUDA, "comm_init_rank: %s", cudaGetErrorString(e)); }
    nr = api->AllGather(d_cards + rank, d_cards, sizeof(RankCard), ncclUint8, m->comm, c->stream);
    if (nr == ncclSuccess) e = cudaMemcpyAsync(cards.data(), d_cards, nranks * sizeof(RankCard), cudaMemcpyDeviceToHost, c->stream);
    if (nr == ncclSuccess && e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_cards);
    if (nr != ncclSuccess || e != cudaSuccess) {
        comm_detach(c);
        return fail(c, DIPSB_ERR_CUDA, "comm_init_rank: card exchange failed (%s / %s)", api->GetErrorString(nr), cudaGetErrorString(e));
    }
    int ok = 1;
    for (uint32_t r = 0; r < nranks; ++r) {
        if (cards[r].n_elems != mine.n_elems || cards[r].npx != mine.npx || cards[r].win_bytes != mine.win_bytes) {
            comm_detach(c);
            return fail(c, DIPSB_ERR_INVALID, "comm_init_rank: rank %u has another geometry than rank %u", r, rank);
        }
        if (!cards[r].p2p_wanted) ok = 0;
    }
    for (uint32_t r = 0; r < nranks && ok; ++r) {
        if (r == rank) continue;
        void* p[4] = {nullptr, nullptr, nullptr, nullptr};
        const cudaIpcMemHandle_t* h[4] = {&cards[r].win, &cards[r].acc, &cards[r].state0, &cards[r].state1};
        for (int k = 0; k < 4 && ok; ++k) {
            if (cudaIpcOpenMemHandle(&p[k], *h[k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; }
            else m->opened[r][k] = true;
        }
        m->win_peer[r] = (uint8_t*)p[0]; m->acc_peer[r] = (uint32_t*)p[1];
        m->state_peer[r][0] = (uint16_t*)p[2]; m->state_peer[r][1] = (uint16_t*)p[3];
    }
    // every rank must take the same path: agree on the minimum
    int* d_ok = nullptr;
    e = cudaMalloc(&d_ok, sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_ok, &ok, sizeof ok, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) nr = api->AllReduce(d_ok, d_ok, 1, ncclInt32, ncclMin, m->comm, c->stream);
    if (e == cudaSuccess && nr == ncclSuccess) e = cudaMemcpyAsync(&ok, d_ok, sizeof ok, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess && nr == ncclSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_ok);
    if (nr != ncclSuccess || e != cudaSuccess) {
        comm_detach(c);
        return fail(c, DIPSB_ERR_CUDA, "comm_init_rank: agreement failed (%s / %s)", api->GetErrorString(nr), cudaGetErrorString(e));
    }
    m->p2p = ok != 0;
    return DIPSB_OK;
}

extern "C" int32_t dipsb_comm_destroy(dipsb_ctx* c) {
    if (!c) return DIPSB_ERR_INVALID;
    if (c->comm && c->comm->group) return fail(c, DIPSB_ERR_STATE, "comm_destroy: the context belongs to a group (dipsb_destroy_group)");
    comm_detach(c);
    return DIPSB_OK;
}

extern "C" int32_t dipsb_comm_set_reduce(dipsb_ctx* c, int32_t path) {
    if (!c || !c->comm) return DIPSB_ERR_INVALID;
    if (path < 0 || path > 2) return fail(c, DIPSB_ERR_INVALID, "comm_set_reduce: %d is not 0 (automatic), 1 (peer memory) or 2 (NCCL)", path);
    if (path == DIPSB_REDUCE_P2P && !c->comm->p2p) return fail(c, DIPSB_ERR_STATE, "comm_set_reduce: peer memory is not mapped on this communicator");
    if (path == DIPSB_REDUCE_NCCL && !c->comm->comm) return fail(c, DIPSB_ERR_STATE, "comm_set_reduce: this communicator has no NCCL (loopback group)");
    c->comm->reduce_path = path;
    return DIPSB_OK;
}

extern "C" int32_t dipsb_comm_info(const dipsb_ctx* c, uint32_t out[8]) {
    if (!c || !out) return DIPSB_ERR_INVALID;
    memset(out, 0, 8 * sizeof(uint32_t));
    const Comm* m = c->comm;
    if (!m) { out[0] = 1; return DIPSB_OK; }
    out[0] = (uint32_t)m->nranks; out[1] = (uint32_t)m->rank; out[2] = m->p2p ? 1u : 0u;
    out[3] = m->nccl ? (uint32_t)m->nccl->version : 0u;
    out[4] = (uint32_t)((m->reduce_path == DIPSB_REDUCE_NCCL || !m->p2p) ? DIPSB_REDUCE_NCCL : DIPSB_REDUCE_P2P);
    out[5] = m->single_process ? 1u : 0u;
    out[6] = c->acc_sharded ? 1u : 0u;
    out[7] = m->comm ? 1u : 0u;
    return DIPSB_OK;
}

// a bounded wait timed out somewhere since the last check (a rank never arrived): the results are invalid
extern "C" int32_t dipsb_comm_check(dipsb_ctx* c) {
    if (!c) return DIPSB_ERR_INVALID;
    if (!c->comm || !c->comm->win) return DIPSB_OK;
    CK(c, cudaSetDevice(c->device));
    uint32_t st = 0;
    CK(c, cudaMemcpyAsync(&st, &ctrl_of(c->comm->win)->status, sizeof st, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    if (st) {
        CK(c, cudaMemsetAsync(&ctrl_of(c->comm->win)->status, 0, sizeof st, c->stream));
        return fail(c, DIPSB_ERR_STATE, "comm_check: a wait for a peer GPU timed out (DIPSB_COMM_TIMEOUT_MS); results since the last check are invalid");
    }
    return DIPSB_OK;
}

// ---- one sharded pass, in three phases (the group runner interleaves them over its devices) ----------------------------
namespace {

bool use_p2p_reduce(const Comm* m) { return m->p2p && m->reduce_path != DIPSB_REDUCE_NCCL; }   // both exchanges follow it

// slices of the reference plane for its broadcast over peer memory: the plane is cut into nranks-1 slices; slice j-1 (pixels
// [(j-1)*slice_px, j*slice_px)) is rank j's to receive from rank 0 and to forward to the other ranks > 0
uint64_t plane_slice_px(const Geometry& g, int nranks) {
    const uint64_t groups = (g.npx + 15) / 16, parts = (uint64_t)std::max(1, nranks - 1);
    return 16 * ((groups + parts - 1) / parts);
}
PlaneScatter make_scatter(dipsb_ctx* c) {
    Comm* m = c->comm;
    PlaneScatter sc;
    sc.nranks = (uint32_t)m->nranks;
    sc.slice_px = plane_slice_px(c->g, m->nranks);
    for (int r = 1; r < m->nranks; ++r) {
        sc.plane_peer[r] = m->state_peer[r][0];
        sc.stamp_peer[r] = &ctrl_of(m->win_peer[r])->plane_stamp;
    }
    sc.epoch = m->epoch;
    sc.blocks_done = &ctrl_of(m->win)->blocks_done[2];
    return sc;
}

struct Pass {
    const uint8_t* d_frames = nullptr;     // device clip (or nullptr for a host clip)
    const uint8_t* h_frames = nullptr;
    uint64_t n = 0, stride = 0, first = 0, total = 0;
    ShardExtra extra;
    bool has_extra = false;
    bool scattered = false;                // rank 0: the prime kernel already stored the slices of the plane into their owners
    bool pushed = false;                   // the clip kernel's last flush already handed the partial sums to their owners
};

int32_t validate_pass(dipsb_ctx* c, const Pass& p) {
    Comm* m = c->comm;
    if (!m) return fail(c, DIPSB_ERR_STATE, "run_clip_sharded: no communicator (dipsb_comm_init_rank / dipsb_create_group)");
    if (p.n == 0) return fail(c, DIPSB_ERR_INVALID, "run_clip_sharded: every rank needs at least one frame");
    if (p.first + p.n > p.total) return fail(c, DIPSB_ERR_INVALID, "run_clip_sharded: frames [%llu,%llu) outside the clip of %llu",
                                              (unsigned long long)p.first, (unsigned long long)(p.first + p.n), (unsigned long long)p.total);
    if (c->frames_processed != 0) return fail(c, DIPSB_ERR_STATE, "run_clip_sharded: one pass per dipsb_reset (the context already accumulated %llu frames)",
                                               (unsigned long long)c->frames_processed);
    if (p.n + 1 > (p.total + m->nranks - 1) / m->nranks + 1)
        return fail(c, DIPSB_ERR_INVALID, "run_clip_sharded: %llu frames on one rank exceed ceil(%llu / %d): shard with dipsb_shard_range",
                    (unsigned long long)p.n, (unsigned long long)p.total, m->nranks);
    if (p.stride < c->g.npx * c->g.bpp) return fail(c, DIPSB_ERR_INVALID, "run_clip_sharded: stride smaller than a frame");
    return DIPSB_OK;
}

// phase A (before the broadcast): rank 0 builds the reference plane (overall); per-frame: push my first frame backwards
int32_t pass_begin(dipsb_ctx* c, Pass& p) {
    Comm* m = c->comm;
    const Geometry& g = c->g;
    m->epoch += 1;
    const int parity = (int)(m->epoch & 1);
    c->shard_total_frames = p.total; c->shard_first = p.first; c->shard_n = p.n;
    if (m->nranks == 1) return DIPSB_OK;
    const uint64_t fb = g.npx * g.bpp;
    if (c->cfg.mode == DIPSB_MODE_OVERALL) {
        if (m->rank == 0) {
            if (c->state_cur != 0) return fail(c, DIPSB_ERR_STATE, "run_clip_sharded: overall mode expects the reference in state plane 0 (dipsb_reset first)");
            const bool fuse = use_p2p_reduce(m);
            if (p.d_frames) {
                if (fuse && prime_fast_path(g, p.d_frames)) {
                    const PlaneScatter sc = make_scatter(c);
                    CK(c, launch_prime(g, p.d_frames, c->state[0], c->stream, &sc));
                    p.scattered = true;
                } else {
                    CK(c, launch_prime(g, p.d_frames, c->state[0], c->stream));
                }
            } else {   // host clip: frame 0 goes up on its own first
                if (c->d_frame_bytes < fb) {
                    cudaFree(c->d_frame); c->d_frame = nullptr; c->d_frame_bytes = 0;
                    CK(c, cudaMalloc(&c->d_frame, fb));
                    c->d_frame_bytes = fb;
                }
                cudaPointerAttributes attr;
                const bool pinned = cudaPointerGetAttributes(&attr, p.h_frames) == cudaSuccess && attr.type == cudaMemoryTypeHost;
                if (!pinned) cudaGetLastError();
                if (pinned) {
                    CK(c, cudaMemcpyAsync(c->d_frame, p.h_frames, fb, cudaMemcpyHostToDevice, c->stream));
                } else {
                    if (c->h_pin_bytes < std::max<size_t>(fb, g.npx * 4)) {
                        if (c->h_pin) cudaFreeHost(c->h_pin);
                        c->h_pin = nullptr; c->h_pin_bytes = 0;
                        CK(c, cudaMallocHost(&c->h_pin, std::max<size_t>(fb, g.npx * 4)));
                        c->h_pin_bytes = std::max<size_t>(fb, g.npx * 4);
                    }
                    CK(c, cudaStreamSynchronize(c->stream));     // an earlier upload from the bounce buffer
                    host_copy2d(c->h_pin, fb, p.h_frames, fb, fb, 1);
                    CK(c, cudaMemcpyAsync(c->d_frame, c->h_pin, fb, cudaMemcpyHostToDevice, c->stream));
                }
                if (fuse && prime_fast_path(g, c->d_frame)) {
                    const PlaneScatter sc = make_scatter(c);
                    CK(c, launch_prime(g, c->d_frame, c->state[0], c->stream, &sc));
                    p.scattered = true;
                } else {
                    CK(c, launch_prime(g, c->d_frame, c->state[0], c->stream));
                }
            }
            c->state_valid = true;
        }
        return DIPSB_OK;
    }
    // ---- per-frame mode ----
    const bool p2p_halo = m->p2p;
    if (m->rank + 1 < m->nranks) {   // my successor's first frame will be my extra trailing frame
        p.has_extra = true;
        p.extra.timeout_ns = m->timeout_ns;
        p.extra.status = &ctrl_of(m->win)->status;
        if (p2p_halo) {
            p.extra.frame = halo_buf(m, m->win, parity);
            p.extra.flag = &ctrl_of(m->win)->halo_stamp[parity];
            p.extra.epoch = m->epoch;
        } else {
            if (!m->halo_local) {
                CK(c, cudaMalloc(&m->halo_local, m->halo_bytes));
                CK(c, cudaMemsetAsync(m->halo_local, 0, m->halo_bytes, c->stream));
            }
            p.extra.frame = m->halo_local;     // filled by ncclRecv in pass_exchange, stream ordered: no flag
        }
    }
    if (p2p_halo && m->rank > 0 && p.d_frames) {
        // device clip: my first frame -> the previous rank's halo buffer, by the copy engine, while the kernels run.  The
        // stamp (my_stamp[parity] := epoch) is written on the main stream before the clip kernel fills the SMs; the copy
        // stream then ships the frame and, behind it, the stamp.
        Control* mine = ctrl_of(m->win);
        if (m->halo_used[parity]) CK(c, cudaStreamWaitEvent(c->stream, m->ev_halo[parity], 0));   // stamp word still in flight?
        stamp_kernel<<<1, 1, 0, c->stream>>>(&mine->my_stamp[parity], m->epoch);
        count_launch();
        CK(c, cudaGetLastError());
        CK(c, cudaEventRecord(m->ev_start, c->stream));
        CK(c, cudaStreamWaitEvent(c->copy_stream, m->ev_start, 0));
        uint8_t* dst_win = m->win_peer[m->rank - 1];
        CK(c, cudaMemcpyAsync(halo_buf(m, dst_win, parity), p.d_frames, fb, cudaMemcpyDeviceToDevice, c->copy_stream));
        CK(c, cudaMemcpyAsync(&ctrl_of(dst_win)->halo_stamp[parity], &mine->my_stamp[parity], sizeof(unsigned long long),
                              cudaMemcpyDeviceToDevice, c->copy_stream));
        CK(c, cudaEventRecord(m->ev_halo[parity], c->copy_stream));
        m->halo_used[parity] = true;
    }
    return DIPSB_OK;
}

// host clip, per-frame mode, peer memory: the halo push happens once the first chunk is on the device
int32_t push_halo_after_upload(dipsb_ctx* c, const uint8_t* d_first_frame, void*) {
    Comm* m = c->comm;
    if (!m || !m->p2p || m->rank == 0 || c->cfg.mode != DIPSB_MODE_PERFRAME) return DIPSB_OK;
    const int parity = (int)(m->epoch & 1);
    const uint64_t fb = c->g.npx * c->g.bpp;
    Control* mine = ctrl_of(m->win);
    // the stamp word is written through the copy stream itself here (no clip kernel is running yet on this rank's first
    // chunk, and the copy stream already holds the upload this frame came with)
    if (m->halo_used[parity]) CK(c, cudaStreamWaitEvent(c->copy_stream, m->ev_halo[parity], 0));
    stamp_kernel<<<1, 1, 0, c->copy_stream>>>(&mine->my_stamp[parity], m->epoch);
    count_launch();
    CK(c, cudaGetLastError());
    uint8_t* dst_win = m->win_peer[m->rank - 1];
    CK(c, cudaMemcpyAsync(halo_buf(m, dst_win, parity), d_first_frame, fb, cudaMemcpyDeviceToDevice, c->copy_stream));
    CK(c, cudaMemcpyAsync(&ctrl_of(dst_win)->halo_stamp[parity], &mine->my_stamp[parity], sizeof(unsigned long long),
                          cudaMemcpyDeviceToDevice, c->copy_stream));
    CK(c, cudaEventRecord(m->ev_halo[parity], c->copy_stream));
    m->halo_used[parity] = true;
    return DIPSB_OK;
}

// phase B: the exchange before the pass that needs every rank (inside ncclGroupStart/End when one thread drives several)
int32_t pass_exchange(dipsb_ctx* c, Pass& p, int phase = 7) {
    Comm* m = c->comm;
    const Geometry& g = c->g;
    if (m->nranks == 1) return DIPSB_OK;
    if (c->cfg.mode == DIPSB_MODE_OVERALL) {
        if (!use_p2p_reduce(m)) {   // NCCL: one broadcast of the u16 plane from rank 0
            if (!m->comm) return fail(c, DIPSB_ERR_STATE, "run_clip_sharded: neither peer memory nor NCCL available");
            if (phase & 1) NK(c, m->nccl, m->nccl->Broadcast(c->state[0], c->state[0], g.npx * sizeof(uint16_t), ncclUint8, 0, m->comm, c->stream));
            return DIPSB_OK;
        }
        // Peer memory: scatter + all-gather.  Rank 0 stores slice j-1 of the plane into rank j (fused into its prime kernel when
        // the frame allows) and starts its pass at once; every other rank forwards its slice to the rest, then waits for theirs.
        // Rank 0 sends the plane once in total (not N-1 times), and nothing is passed on hop by hop as in a ring (measured at
        // 8 GPUs: the last rank of NCCL's broadcast ring started 0.26 ms after the first).
        const uint64_t slice_px = plane_slice_px(g, m->nranks), top = (g.npx + 15) / 16 * 16;
        PlanePush P{};
        P.plane = c->state[0]; P.nranks = (uint32_t)m->nranks; P.rank = (uint32_t)m->rank;
        P.timeout_ns = m->timeout_ns; P.status = &ctrl_of(m->win)->status;
        if ((phase & 1) && m->rank == 0 && !p.scattered) {
            for (int r = 1; r < m->nranks; ++r) {
                P.dst[r] = m->state_peer[r][0];
                P.lo[r] = std::min((uint64_t)(r - 1) * slice_px, top) / 8; P.hi[r] = std::min((uint64_t)r * slice_px, top) / 8;
                P.stamp_peer[r] = &ctrl_of(m->win_peer[r])->plane_stamp;
            }
            P.stamp = m->epoch; P.blocks_done = &ctrl_of(m->win)->blocks_done[2];
            plane_push_kernel<<<kXchgBlocks, kXchgThreads, 0, c->stream>>>(P);
            count_launch();
            CK(c, cudaGetLastError());
        }
        if ((phase & 2) && m->rank > 0) {
            for (int r = 1; r < m->nranks; ++r) {
                if (r == m->rank) continue;
                P.dst[r] = m->state_peer[r][0];
                P.lo[r] = std::min((uint64_t)(m->rank - 1) * slice_px, top) / 8; P.hi[r] = std::min((uint64_t)m->rank * slice_px, top) / 8;
                P.stamp_peer[r] = ctrl_of(m->win_peer[r])->plane_stamp2 + m->rank;
            }
            P.wait_flag = &ctrl_of(m->win)->plane_stamp; P.wait_value = m->epoch;
            P.stamp = m->epoch; P.blocks_done = &ctrl_of(m->win)->blocks_done[3];
            plane_push_kernel<<<kXchgBlocks / 2, kXchgThreads, 0, c->stream>>>(P);
            count_launch();
            CK(c, cudaGetLastError());
        }
        if ((phase & 4) && m->rank > 0 && m->nranks > 2) {
            const uint32_t mask = ((1u << m->nranks) - 1u) & ~1u & ~(1u << m->rank);
            wait_sources_kernel<<<1, 32, 0, c->stream>>>(ctrl_of(m->win)->plane_stamp2, mask, m->epoch, m->timeout_ns, &ctrl_of(m->win)->status);
            count_launch();
            CK(c, cudaGetLastError());
        }
        return DIPSB_OK;
    }
    if (!(phase & 1)) return DIPSB_OK;
    if (!m->p2p && m->comm) {   // halo by NCCL: my first frame -> previous rank (device clips only; host clips upload it first)
        const uint64_t fb = g.npx * g.bpp;
        const uint8_t* src = p.d_frames;
        if (!src && m->rank > 0) {
            if (c->d_frame_bytes < fb) {
                cudaFree(c->d_frame); c->d_frame = nullptr; c->d_frame_bytes = 0;
                CK(c, cudaMalloc(&c->d_frame, fb));
                c->d_frame_bytes = fb;
            }
            CK(c, cudaMemcpyAsync(c->d_frame, p.h_frames, fb, cudaMemcpyHostToDevice, c->stream));   // pageable: staged by the driver
            src = c->d_frame;
        }
        NK(c, m->nccl, m->nccl->GroupStart());
        if (m->rank > 0) NK(c, m->nccl, m->nccl->Send(src, fb, ncclUint8, m->rank - 1, m->comm, c->stream));
        if (m->rank + 1 < m->nranks) NK(c, m->nccl, m->nccl->Recv(m->halo_local, fb, ncclUint8, m->rank + 1, m->comm, c->stream));
        NK(c, m->nccl, m->nccl->GroupEnd());
    }
    return DIPSB_OK;
}

// phase C: the pass itself
int32_t pass_run(dipsb_ctx* c, Pass& p) {
    Comm* m = c->comm;
    if (m->nranks > 1 && c->cfg.mode == DIPSB_MODE_OVERALL && m->rank > 0) c->state_valid = true;   // the plane has arrived
    if (m->nranks > 1 && use_p2p_reduce(m) && m->fuse_push) {
        const int parity = (int)(m->epoch & 1);
        uint64_t plan[4];
        dipsb_xchg_plan_query(p.total, (uint32_t)m->nranks, c->g.n_elems, plan);
        ShardPush& x = p.extra.push;
        x.nranks = (uint32_t)m->nranks; x.rank = (uint32_t)m->rank; x.chunk = m->chunk;
        x.fmt = plan[0] == 4 ? 1 : 2; x.sum_bits = (int)plan[1];
        for (int r = 0; r < m->nranks; ++r)
            x.recv[r] = reinterpret_cast<uint32_t*>(recv_area(m, m->win_peer[r], parity) + (uint64_t)m->rank * m->slot_bytes);
        p.extra.pushed = &p.pushed;
        // overall mode: the scalars are nobody else's business -- finalise them after the peers have been told (pass_reduce);
        // per-frame mode hands the boundary frame's scalars over with the stamp, so they have to exist first
        p.extra.defer_finalize = c->cfg.mode == DIPSB_MODE_OVERALL && p.d_frames != nullptr;
    }
    const ShardExtra* extra = (p.has_extra || p.extra.push.nranks) ? &p.extra : nullptr;
    if (p.d_frames) return run_clip_on_stream(c, p.d_frames, p.n, p.stride, p.first, false, extra);
    HostClipHooks hooks;
    hooks.extra = extra;
    hooks.after_first_upload = push_halo_after_upload;
    return run_clip_host_impl(c, p.h_frames, p.n, p.stride, p.first, &hooks);
}

// phase D: combine the accumulators (and hand the boundary scalars to their owner).  Peer-memory path: phase bit 0 pushes
// the partial sums to their owners, bit 1 adds up the owned range (the group runner separates the two in loopback mode)
int32_t pass_reduce(dipsb_ctx* c, Pass& p, int phase = 15) {
    Comm* m = c->comm;
    const Geometry& g = c->g;
    if (m->nranks == 1) return finalize_pending(c);
    { int32_t zrc = ensure_acc_zero(c); if (zrc) return zrc; }   // (only a pass that never reached a clip kernel leaves it pending)
    if (!use_p2p_reduce(m)) { int32_t frc = finalize_pending(c); if (frc) return frc; }
    const bool perframe = c->cfg.mode == DIPSB_MODE_PERFRAME;
    if (use_p2p_reduce(m)) {
        const int parity = (int)(m->epoch & 1);
        uint64_t plan[4];
        dipsb_xchg_plan_query(p.total, (uint32_t)m->nranks, g.n_elems, plan);
        XchgParams X{};
        X.acc = c->acc; X.n_elems = g.n_elems; X.chunk = m->chunk; X.nranks = (uint32_t)m->nranks; X.rank = (uint32_t)m->rank;
        X.fmt = plan[0] == 4 ? 1 : 2; X.sum_bits = (int)plan[1]; X.slot_bytes = m->slot_bytes;
        X.recv_local = recv_area(m, m->win, parity);
        for (int r = 0; r < m->nranks; ++r) {
            X.recv_peer[r] = recv_area(m, m->win_peer[r], parity);
            X.stamp_peer[r] = r == m->rank ? nullptr : ctrl_of(m->win_peer[r])->xchg_arrived + m->rank;
        }
        X.arrived_local = ctrl_of(m->win)->xchg_arrived;
        X.target = m->epoch;                       // stamp of this pass
        X.timeout_ns = m->timeout_ns; X.status = &ctrl_of(m->win)->status;
        X.blocks_done = &ctrl_of(m->win)->blocks_done[0];
        if (perframe && m->rank + 1 < m->nranks) {
            X.sad_src = reinterpret_cast<const unsigned long long*>(c->d_sad + p.first + p.n);
            X.cnt_src = reinterpret_cast<const unsigned long long*>(c->d_cnt + p.first + p.n);
            X.mbox_next = ctrl_of(m->win_peer[m->rank + 1])->mbox[parity];
        }
        if (perframe && m->rank > 0) {
            X.mbox_local = ctrl_of(m->win)->mbox[parity];
            X.sad_dst = reinterpret_cast<unsigned long long*>(c->d_sad + p.first);
            X.cnt_dst = reinterpret_cast<unsigned long long*>(c->d_cnt + p.first);
        }
        if (phase & 1) {
            // the clip kernel's last flush may have stored the partials into their owners already: then one block hands
            // over the boundary scalars and stamps the peers (its launch follows the clip kernel's completion, hence its stores)
            if (p.pushed) { X.n_elems = 0; xchg_push_kernel<<<1, kXchgThreads, 0, c->stream>>>(X); X.n_elems = g.n_elems; }
            else xchg_push_kernel<<<kXchgBlocks, kXchgThreads, 0, c->stream>>>(X);
            count_launch();
            CK(c, cudaGetLastError());
            { int32_t frc = finalize_pending(c); if (frc) return frc; }   // behind the stamp, ahead of the wait for the peers
        }
        if (phase & 2) {
            xchg_reduce_kernel<<<kXchgBlocks, kXchgThreads, 0, c->stream>>>(X);
            count_launch();
            CK(c, cudaGetLastError());
            c->acc_sharded = true;
        }
        return DIPSB_OK;
    }
    if (!m->comm) return fail(c, DIPSB_ERR_STATE, "run_clip_sharded: neither peer memory nor NCCL available");
    // NCCL path: pack (bit 0) -> all-reduce (bit 1) -> unpack (bit 2) -> boundary scalars to their owner (bit 3); the totals
    // end up replicated on every rank.  A single-process group wraps the NCCL bits of all its ranks in ncclGroupStart/End.
    if (phase & 1) {
        int32_t rc = dipsb_pack_accumulators_device(c, p.total + (uint64_t)m->nranks, &m->packed, &m->packed_words);
        if (rc) return rc;
    }
    if (phase & 2) NK(c, m->nccl, m->nccl->AllReduce(m->packed, m->packed, m->packed_words, ncclInt32, ncclSum, m->comm, c->stream));
    if (phase & 4) {
        int32_t rc = dipsb_unpack_accumulators_device(c);
        if (rc) return rc;
        c->acc_sharded = false;
    }
    if ((phase & 8) && perframe) {   // the boundary frame's scalars go to the rank that owns the frame
        NK(c, m->nccl, m->nccl->GroupStart());
        if (m->rank + 1 < m->nranks) {
            NK(c, m->nccl, m->nccl->Send(c->d_sad + p.first + p.n, 1, ncclUint64, m->rank + 1, m->comm, c->stream));
            NK(c, m->nccl, m->nccl->Send(c->d_cnt + p.first + p.n, 1, ncclUint64, m->rank + 1, m->comm, c->stream));
        }
        if (m->rank > 0) {
            NK(c, m->nccl, m->nccl->Recv(c->d_sad + p.first, 1, ncclUint64, m->rank - 1, m->comm, c->stream));
            NK(c, m->nccl, m->nccl->Recv(c->d_cnt + p.first, 1, ncclUint64, m->rank - 1, m->comm, c->stream));
        }
        NK(c, m->nccl, m->nccl->GroupEnd());
    }
    return DIPSB_OK;
}

int32_t run_sharded(dipsb_ctx* c, Pass& p) {
    int32_t rc = validate_pass(c, p);
    if (rc) return rc;
    if (c->comm->single_process && c->comm->nranks > 1)
        return fail(c, DIPSB_ERR_STATE, "run_clip_sharded: this context belongs to a single-process group; use dipsb_group_run_clip_device");
    // optional phase timing (dipsb_enable_timing): begin | reference / halo exchange | pass | accumulator exchange
    cudaEvent_t* ev = nullptr;
    if (c->timing) {
        if (c->pev_used + 4 > c->pev.size())
            for (int k = 0; k < 4; ++k) {
                cudaEvent_t e;
                CK(c, cudaEventCreate(&e));
                c->pev.push_back(e);
            }
        ev = &c->pev[c->pev_used];
        c->pev_used += 4;
        CK(c, cudaEventRecord(ev[0], c->stream));
    }
    if ((rc = pass_begin(c, p))) return rc;
    if ((rc = pass_exchange(c, p))) return rc;
    if (ev) CK(c, cudaEventRecord(ev[1], c->stream));
    if ((rc = pass_run(c, p))) return rc;
    if (ev) CK(c, cudaEventRecord(ev[2], c->stream));
    rc = pass_reduce(c, p);
    if (ev) CK(c, cudaEventRecord(ev[3], c->stream));
    return rc;
}

}  // namespace

extern "C" int32_t dipsb_run_clip_sharded_device(dipsb_ctx* c, const void* d_frames, uint64_t n_frames, uint64_t stride,
                                                 uint64_t first_frame_index, uint64_t total_frames) {
    if (!c || !d_frames) return DIPSB_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    Pass p;
    p.d_frames = (const uint8_t*)d_frames; p.n = n_frames; p.stride = stride; p.first = first_frame_index; p.total = total_frames;
    return run_sharded(c, p);
}

extern "C" int32_t dipsb_run_clip_sharded_host(dipsb_ctx* c, const uint8_t* frames, uint64_t n_frames, uint64_t stride,
                                               uint64_t first_frame_index, uint64_t total_frames) {
    if (!c || !frames) return DIPSB_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    Pass p;
    p.h_frames = frames; p.n = n_frames; p.stride = stride; p.first = first_frame_index; p.total = total_frames;
    return run_sharded(c, p);
}

extern "C" int32_t dipsb_comm_phase_times(dipsb_ctx* c, double out_ms[3], uint64_t* passes) {
    if (!c || !out_ms || !passes) return DIPSB_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaStreamSynchronize(c->stream));
    out_ms[0] = out_ms[1] = out_ms[2] = 0.0;
    for (size_t i = 0; i + 3 < c->pev_used; i += 4)
        for (int k = 0; k < 3; ++k) {
            float ms = 0.f;
            CK(c, cudaEventElapsedTime(&ms, c->pev[i + k], c->pev[i + k + 1]));
            out_ms[k] += ms;
        }
    *passes = c->pev_used / 4;
    c->pev_used = 0;
    return DIPSB_OK;
}

static int32_t gather_enqueue(dipsb_ctx* c, int phase);

// Measurement aid (collective): mean milliseconds of `reps` back-to-back exchanges on this rank's stream, without a pass in
// between -- what = 0: accumulator reduce-scatter over peer memory (push + reduce kernels), 1: reference-plane ncclBroadcast
// from rank 0, 2: all-gather of the totals, 3: pack + ncclAllReduce + unpack, 4: reference-plane scatter + all-gather over
// peer memory.  Leaves the accumulators and the state plane undefined (dipsb_reset).
extern "C" int32_t dipsb_comm_probe(dipsb_ctx* c, int32_t what, uint64_t total_frames, uint32_t reps, float* ms) {
    if (!c || !ms || !reps || what < 0 || what > 4) return DIPSB_ERR_INVALID;
    Comm* m = c->comm;
    if (!m || m->nranks < 2 || m->single_process) return fail(c, DIPSB_ERR_STATE, "comm_probe: needs a multi-process communicator of at least 2 ranks");
    if ((what == 0 || what == 2 || what == 4) && !m->p2p) return fail(c, DIPSB_ERR_STATE, "comm_probe: peer memory is not mapped");
    if ((what == 1 || what == 3) && !m->comm) return fail(c, DIPSB_ERR_STATE, "comm_probe: no NCCL communicator");
    CK(c, cudaSetDevice(c->device));
    cudaEvent_t e0, e1;
    CK(c, cudaEventCreate(&e0));
    CK(c, cudaEventCreate(&e1));
    Pass p;
    p.total = total_frames ? total_frames : 1; p.n = 1; p.first = (uint64_t)m->rank;
    const int keep_path = m->reduce_path;
    for (uint32_t r = 0; r <= reps; ++r) {          // rep 0 is a warm-up
        if (r == 1) CK(c, cudaEventRecord(e0, c->stream));
        int32_t rc = DIPSB_OK;
        if (what == 0 || what == 3) {
            m->epoch += 1;
            m->reduce_path = what == 0 ? DIPSB_REDUCE_P2P : DIPSB_REDUCE_NCCL;
            const int mode = c->cfg.mode;
            c->cfg.mode = DIPSB_MODE_OVERALL;        // no boundary scalars in the probe
            const uint64_t fp = c->frames_processed;
            c->frames_processed = 0;
            rc = pass_reduce(c, p);
            c->frames_processed = fp;
            c->cfg.mode = mode;
            m->reduce_path = keep_path;
        } else if (what == 1) {
            uint16_t* plane = c->state[c->state_cur];
            NK(c, m->nccl, m->nccl->Broadcast(plane, plane, c->g.npx * sizeof(uint16_t), ncclUint8, 0, m->comm, c->stream));
        } else if (what == 4) {                      // scatter + forward + wait over peer memory (the plane itself, no prime)
            m->epoch += 1;
            m->reduce_path = DIPSB_REDUCE_P2P;
            const int mode = c->cfg.mode;
            c->cfg.mode = DIPSB_MODE_OVERALL;
            rc = pass_exchange(c, p, 7);
            c->cfg.mode = mode;
            m->reduce_path = keep_path;
        } else {
            c->acc_sharded = true;
            rc = gather_enqueue(c, 3);
        }
        if (rc) return rc;
    }
    CK(c, cudaEventRecord(e1, c->stream));
    CK(c, cudaEventSynchronize(e1));
    float total = 0.f;
    CK(c, cudaEventElapsedTime(&total, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    c->acc_sharded = false;
    *ms = total / (float)reps;
    return DIPSB_OK;
}

// collective: after it every rank holds the complete accumulator planes
static int32_t gather_enqueue(dipsb_ctx* c, int phase) {
    Comm* m = c->comm;
    if (!m || m->nranks == 1 || !c->acc_sharded) { c->acc_sharded = false; return DIPSB_OK; }
    { int32_t zrc = ensure_acc_zero(c); if (zrc) return zrc; }
    const Geometry& g = c->g;
    GatherParams G{};
    G.acc = c->acc; G.n_elems = g.n_elems; G.chunk = m->chunk; G.nranks = (uint32_t)m->nranks; G.rank = (uint32_t)m->rank;
    for (int r = 0; r < m->nranks; ++r) {
        G.acc_peer[r] = m->acc_peer[r];
        G.stamp_peer[r] = r == m->rank ? nullptr : ctrl_of(m->win_peer[r])->gather_arrived + m->rank;
    }
    G.arrived_local = ctrl_of(m->win)->gather_arrived;
    G.target = m->gathers;
    G.timeout_ns = m->timeout_ns; G.status = &ctrl_of(m->win)->status;
    G.blocks_done = &ctrl_of(m->win)->blocks_done[1];
    if (phase & 1) {
        m->gathers += 1;
        G.target = m->gathers;
        gather_push_kernel<<<kXchgBlocks, kXchgThreads, 0, c->stream>>>(G);
        count_launch();
        CK(c, cudaGetLastError());
    }
    if (phase & 2) {
        wait_sources_kernel<<<1, 32, 0, c->stream>>>(G.arrived_local, ((1u << m->nranks) - 1u) & ~(1u << m->rank), G.target, G.timeout_ns, G.status);
        count_launch();
        CK(c, cudaGetLastError());
        c->acc_sharded = false;
    }
    return DIPSB_OK;
}

extern "C" int32_t dipsb_gather_accumulators(dipsb_ctx* c) {
    if (!c) return DIPSB_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    if (c->comm && c->comm->single_process && c->comm->nranks > 1)
        return fail(c, DIPSB_ERR_STATE, "gather_accumulators: this context belongs to a single-process group; use dipsb_group_gather_accumulators");
    return gather_enqueue(c, 3);
}

// ---- single-process group: one handle, all GPUs (SURVEY.md 8(b) dipsb_create_group, 8(e) ncclCommInitAll) -------------
static thread_local std::string g_group_err;

static int32_t gfail(dipsb_group* grp, int32_t code, const std::string& msg) {
    if (grp) grp->err = msg;
    g_group_err = msg;
    g_create_err = msg;
    return code;
}

extern "C" void dipsb_destroy_group(dipsb_group* grp) {
    if (!grp) return;
    for (dipsb_ctx* c : grp->ctx) {
        if (!c) continue;
        if (c->comm) c->comm->group = nullptr;
        dipsb_destroy(c);
    }
    delete grp;
}

extern "C" int32_t dipsb_create_group(const dipsb_config* cfg, uint32_t ndev, const int32_t* devices, dipsb_group** out) {
    if (!cfg || !out || ndev == 0 || ndev > (uint32_t)kMaxRanks) return gfail(nullptr, DIPSB_ERR_INVALID, "create_group: bad arguments");
    *out = nullptr;
    if (cfg->flavor != DIPSB_FLAVOR_FRAME0 || cfg->spatial_window > 1)
        return gfail(nullptr, DIPSB_ERR_INVALID, "create_group: sharded passes need the FRAME0 flavour and spatial_window 1");
    dipsb_group* grp = new (std::nothrow) dipsb_group();
    if (!grp) return gfail(nullptr, DIPSB_ERR_NOMEM, "create_group: out of host memory");
    std::vector<int> devs(ndev);
    for (uint32_t i = 0; i < ndev; ++i) devs[i] = devices ? devices[i] : (int)i;
    bool distinct = true;
    for (uint32_t i = 0; i < ndev; ++i)
        for (uint32_t j = i + 1; j < ndev; ++j) distinct = distinct && devs[i] != devs[j];
    grp->loopback = !distinct;
    for (uint32_t i = 0; i < ndev; ++i) {
        dipsb_config one = *cfg;
        one.device = devs[i];
        dipsb_ctx* c = nullptr;
        int32_t rc = dipsb_create(&one, &c);
        if (rc) { const std::string why = dipsb_last_error(nullptr); dipsb_destroy_group(grp); return gfail(nullptr, rc, why); }
        grp->ctx.push_back(c);
    }
    if (ndev == 1) { *out = grp; return DIPSB_OK; }
    // peer access between every pair of distinct devices
    for (uint32_t i = 0; i < ndev; ++i)
        for (uint32_t j = 0; j < ndev; ++j) {
            if (devs[i] == devs[j]) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devs[i], devs[j]);
            if (!can) { dipsb_destroy_group(grp); return gfail(nullptr, DIPSB_ERR_CUDA, "create_group: no peer access between the devices"); }
            cudaSetDevice(devs[i]);
            cudaError_t e = cudaDeviceEnablePeerAccess(devs[j], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                dipsb_destroy_group(grp);
                return gfail(nullptr, DIPSB_ERR_CUDA, std::string("create_group: cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
            }
            cudaGetLastError();
        }
    std::vector<ncclComm_t> comms(ndev, nullptr);
    if (distinct) {
        std::string why;
        grp->nccl = nccl_api(&why);
        if (!grp->nccl) { dipsb_destroy_group(grp); return gfail(nullptr, DIPSB_ERR_CUDA, "create_group: " + why); }
        ncclResult_t nr = grp->nccl->CommInitAll(comms.data(), (int)ndev, devs.data());
        if (nr != ncclSuccess) {
            const std::string why2 = grp->nccl->GetErrorString(nr);
            dipsb_destroy_group(grp);
            return gfail(nullptr, DIPSB_ERR_CUDA, "create_group: ncclCommInitAll: " + why2);
        }
    }
    for (uint32_t i = 0; i < ndev; ++i) {
        dipsb_ctx* c = grp->ctx[i];
        Comm* m = new (std::nothrow) Comm();
        if (!m) { dipsb_destroy_group(grp); return gfail(nullptr, DIPSB_ERR_NOMEM, "create_group: out of host memory"); }
        m->nranks = (int)ndev; m->rank = (int)i; m->single_process = true; m->loopback = grp->loopback;
        m->nccl = grp->nccl; m->comm = comms[i]; m->group = grp; m->p2p = true;
        c->comm = m;
        cudaSetDevice(c->device);
        int32_t rc = alloc_window(c, m);
        if (rc) { const std::string why = c->err; dipsb_destroy_group(grp); return gfail(nullptr, rc, why); }
    }
    for (uint32_t i = 0; i < ndev; ++i)
        for (uint32_t j = 0; j < ndev; ++j) {
            Comm* m = grp->ctx[i]->comm;
            m->win_peer[j] = grp->ctx[j]->comm->win;
            m->acc_peer[j] = grp->ctx[j]->acc;
            m->state_peer[j][0] = grp->ctx[j]->state[0];
            m->state_peer[j][1] = grp->ctx[j]->state[1];
        }
    *out = grp;
    return DIPSB_OK;
}

extern "C" uint32_t dipsb_group_size(const dipsb_group* grp) { return grp ? (uint32_t)grp->ctx.size() : 0; }
extern "C" dipsb_ctx* dipsb_group_ctx(dipsb_group* grp, uint32_t i) { return (grp && i < grp->ctx.size()) ? grp->ctx[i] : nullptr; }
extern "C" const char* dipsb_group_last_error(const dipsb_group* grp) { return grp ? grp->err.c_str() : g_group_err.c_str(); }

#define GRP(grp, i, call)                                                                       \
    do {                                                                                        \
        int32_t rc__ = (call);                                                                  \
        if (rc__) return gfail((grp), rc__, std::string("rank ") + std::to_string(i) + ": " + (grp)->ctx[i]->err); \
    } while (0)

extern "C" int32_t dipsb_group_reset(dipsb_group* grp) {
    if (!grp) return DIPSB_ERR_INVALID;
    for (size_t i = 0; i < grp->ctx.size(); ++i) GRP(grp, i, dipsb_reset(grp->ctx[i]));
    return DIPSB_OK;
}

extern "C" int32_t dipsb_group_synchronize(dipsb_group* grp) {
    if (!grp) return DIPSB_ERR_INVALID;
    for (size_t i = 0; i < grp->ctx.size(); ++i) GRP(grp, i, dipsb_synchronize(grp->ctx[i]));
    for (size_t i = 0; i < grp->ctx.size(); ++i) GRP(grp, i, dipsb_comm_check(grp->ctx[i]));
    return DIPSB_OK;
}

// in loopback mode (several ranks on one device) a kernel that waits for a peer must not start before that peer's
// preceding kernels have finished -- they could not become resident next to it
static int32_t loopback_fence(dipsb_group* grp) {
    if (!grp->loopback) return DIPSB_OK;
    for (size_t i = 0; i < grp->ctx.size(); ++i) {
        dipsb_ctx* c = grp->ctx[i];
        cudaSetDevice(c->device);
        if (cudaStreamSynchronize(c->stream) != cudaSuccess || cudaStreamSynchronize(c->copy_stream) != cudaSuccess)
            return gfail(grp, DIPSB_ERR_CUDA, "group: synchronisation failed on rank " + std::to_string(i));
    }
    return DIPSB_OK;
}

extern "C" int32_t dipsb_group_run_clip_device(dipsb_group* grp, const void* const* d_frames, const uint64_t* n_frames,
                                               uint64_t stride) {
    if (!grp || !d_frames || !n_frames) return DIPSB_ERR_INVALID;
    const size_t R = grp->ctx.size();
    std::vector<Pass> pass(R);
    uint64_t total = 0;
    for (size_t i = 0; i < R; ++i) total += n_frames[i];
    uint64_t first = 0;
    for (size_t i = 0; i < R; ++i) {
        pass[i].d_frames = (const uint8_t*)d_frames[i]; pass[i].n = n_frames[i]; pass[i].stride = stride;
        pass[i].first = first; pass[i].total = total;
        first += n_frames[i];
        if (!pass[i].d_frames) return gfail(grp, DIPSB_ERR_INVALID, "group_run_clip: null shard");
        if (R > 1) GRP(grp, i, validate_pass(grp->ctx[i], pass[i]));
    }
    if (R == 1) {
        cudaSetDevice(grp->ctx[0]->device);
        GRP(grp, 0, run_clip_on_stream(grp->ctx[0], pass[0].d_frames, pass[0].n, stride, 0, false, nullptr));
        return DIPSB_OK;
    }
    for (size_t i = 0; i < R; ++i) { cudaSetDevice(grp->ctx[i]->device); GRP(grp, i, pass_begin(grp->ctx[i], pass[i])); }
    int32_t rc = DIPSB_OK;
    for (int bit : {1, 2, 4}) {         // scatter (or the NCCL calls, grouped) | forward | wait; loopback: one phase at a time
        if (grp->loopback && (rc = loopback_fence(grp))) return rc;
        if (grp->nccl && bit == 1) grp->nccl->GroupStart();
        for (size_t i = 0; i < R; ++i) { cudaSetDevice(grp->ctx[i]->device); GRP(grp, i, pass_exchange(grp->ctx[i], pass[i], bit)); }
        if (grp->nccl && bit == 1) grp->nccl->GroupEnd();
    }
    if ((rc = loopback_fence(grp))) return rc;   // per-frame: the halo pushes; overall: the plane
    for (size_t i = 0; i < R; ++i) { cudaSetDevice(grp->ctx[i]->device); GRP(grp, i, pass_run(grp->ctx[i], pass[i])); }
    if (!use_p2p_reduce(grp->ctx[0]->comm)) {   // NCCL path (the same on every rank): its collective bits grouped
        for (int bit : {1, 2, 4, 8}) {
            const bool nccl_bit = bit == 2 || bit == 8;
            if (nccl_bit) grp->nccl->GroupStart();
            for (size_t i = 0; i < R; ++i) { cudaSetDevice(grp->ctx[i]->device); GRP(grp, i, pass_reduce(grp->ctx[i], pass[i], bit)); }
            if (nccl_bit) grp->nccl->GroupEnd();
        }
        return DIPSB_OK;
    }
    if (!grp->loopback) {
        for (size_t i = 0; i < R; ++i) { cudaSetDevice(grp->ctx[i]->device); GRP(grp, i, pass_reduce(grp->ctx[i], pass[i], 3)); }
        return DIPSB_OK;
    }
    if ((rc = loopback_fence(grp))) return rc;
    for (size_t i = 0; i < R; ++i) { cudaSetDevice(grp->ctx[i]->device); GRP(grp, i, pass_reduce(grp->ctx[i], pass[i], 1)); }
    if ((rc = loopback_fence(grp))) return rc;
    for (size_t i = 0; i < R; ++i) { cudaSetDevice(grp->ctx[i]->device); GRP(grp, i, pass_reduce(grp->ctx[i], pass[i], 2)); }
    return DIPSB_OK;
}

extern "C" int32_t dipsb_group_gather_accumulators(dipsb_group* grp) {
    if (!grp) return DIPSB_ERR_INVALID;
    int32_t rc = loopback_fence(grp);
    if (rc) return rc;
    for (size_t i = 0; i < grp->ctx.size(); ++i) { cudaSetDevice(grp->ctx[i]->device); GRP(grp, i, gather_enqueue(grp->ctx[i], 1)); }
    if ((rc = loopback_fence(grp))) return rc;
    for (size_t i = 0; i < grp->ctx.size(); ++i) { cudaSetDevice(grp->ctx[i]->device); GRP(grp, i, gather_enqueue(grp->ctx[i], 2)); }
    return DIPSB_OK;
}

// the combined maps of the clip (gathers first if the totals are still sharded) and the scalars of all its frames
extern "C" int32_t dipsb_group_get_accumulators(dipsb_group* grp, uint32_t* acc_sum, uint32_t* acc_cnt) {
    if (!grp) return DIPSB_ERR_INVALID;
    int32_t rc = dipsb_group_gather_accumulators(grp);
    if (rc) return rc;
    if ((rc = dipsb_group_synchronize(grp))) return rc;
    GRP(grp, 0, dipsb_get_accumulators(grp->ctx[0], acc_sum, acc_cnt));
    return DIPSB_OK;
}

extern "C" int32_t dipsb_group_get_scalars(dipsb_group* grp, uint64_t first, uint64_t n, uint64_t* sad, uint64_t* cnt) {
    if (!grp) return DIPSB_ERR_INVALID;
    int32_t rc = dipsb_group_synchronize(grp);
    if (rc) return rc;
    for (size_t i = 0; i < grp->ctx.size() && n; ++i) {   // every rank answers for the frames it owns
        dipsb_ctx* c = grp->ctx[i];
        const uint64_t lo = grp->ctx.size() > 1 ? c->shard_first : 0, hi = grp->ctx.size() > 1 ? lo + c->shard_n : c->frames_processed;
        const uint64_t a = std::max(first, lo), b = std::min(first + n, hi);
        if (a < b) GRP(grp, i, dipsb_get_scalars(c, a, b - a, sad ? sad + (a - first) : nullptr, cnt ? cnt + (a - first) : nullptr));
    }
    return DIPSB_OK;
}
