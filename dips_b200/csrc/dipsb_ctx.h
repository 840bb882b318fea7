// dipsb_ctx.h -- the context object behind the opaque dipsb_ctx handle, shared by the host-side translation units
// (api.cu: single-GPU runtime; comm.cu: frame-range shards of a clip over several GPUs).  Not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include <string>
#include <vector>

#include "../../include/dips_b200.h"
#include "dipsb_internal.h"

namespace dipsb { struct Comm; }
using namespace dipsb;   // internal header of the library's own translation units

struct dipsb_ctx {
    dipsb_config cfg;
    Geometry g;
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr, out_stream = nullptr;
    cudaEvent_t ev_copy[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_switch = nullptr;
    uint16_t* state[2] = {nullptr, nullptr};   // u16[n_elems] each, zero padded past npx
    int state_cur = 0;
    bool state_valid = false;
    bool snapshot_pending = false;
    uint32_t* acc = nullptr;                   // u32[2*n_elems]: sum plane then count plane, internal order
    struct { bool pending = false; uint32_t n = 0, words = 0; uint64_t first = 0; } fin;   // deferred finalize_scalars launch
    bool acc_zero_pending = false;             // dipsb_reset deferred the zeroing of `acc`: the next clip kernel does it in its
                                               // prologue, anything else that touches the planes calls ensure_acc_zero first
    uint32_t* planar = nullptr;                // u32[2*npx] scratch for get/set in pixel order
    uint64_t* d_sad = nullptr;                 // per logical frame index
    uint64_t* d_cnt = nullptr;
    uint64_t scal_cap = 0, scal_hi = 0;
    uint32_t* partials = nullptr;
    uint64_t partial_cap = 0;                  // in u32 words
    uint64_t frames_processed = 0;
    uint64_t stream_index = 0;                 // logical index of the next pushed frame
    uint32_t* xchg = nullptr;                  // packed accumulators for the cross-GPU sum (dipsb_pack_accumulators_device)
    int xchg_layout = 0, xchg_sum_bits = 0;
    uint16_t* i2_scratch = nullptr;            // spatial window > 1: 5 planes of npx u16 (raw + up to 4 filtered)
    uint16_t* ring = nullptr;                  // ring flavours: 4 (dips) or 2 (dips_alt) u16 I2 planes of npx
    uint32_t ring_seen = 0, ring_index = 0;    // frames pushed since the last (re)start, next slot to overwrite
    // streaming staging
    uint8_t* d_frame = nullptr; size_t d_frame_bytes = 0;
    uint8_t* d_rgba = nullptr;
    uint8_t* h_pin = nullptr; size_t h_pin_bytes = 0;   // pinned bounce buffer (frame in / rgba out)
    uint64_t* h_stat = nullptr;                          // pinned [2]
    // frame slots of the streaming path (slot 0 only for the synchronous call, both for the pipelined one)
    struct FrameSlot {
        uint8_t* h_in = nullptr; uint8_t* d_in = nullptr; size_t in_bytes = 0;
        uint8_t* d_out = nullptr; uint8_t* h_out = nullptr; uint64_t* h_stat = nullptr;
        cudaEvent_t ev_done = nullptr;
        cudaEvent_t ev_in[8] = {};    // per row band: uploaded
        cudaEvent_t ev_k[8] = {};     //               kernels done
        cudaEvent_t ev_out[8] = {};   //               read back into the staging buffer
        uint64_t out_off[9] = {};                          // byte offsets of the bands in the RGBA frame
        int out_pieces = 0;
        bool pending = false, want_rgba = false; int32_t status = 0; uint64_t idx = 0;
        bool out_direct = false;               // the read-back already targets the caller's (pinned) buffer
        bool out_deferred = false;             // the read-back is enqueued at collection time (pipelined, pinned caller)
    } slot[2];
    int next_slot = 0;
    bool staged = false; int32_t staged_format = 0; uint32_t staged_bands = 1;   // dipsb_stage_frame put a frame into slot 0 and started its upload (in row bands)
    bool out_pinned_hint = false;              // the pipelined caller's output buffers are page-locked
    // host clip staging
    uint8_t* h_chunk[2] = {nullptr, nullptr};
    uint8_t* d_chunk[2] = {nullptr, nullptr};
    size_t chunk_bytes = 0;
    bool chunk_used[2] = {false, false};       // ev_copy / ev_done of the slot were recorded (possibly by an earlier call)
    int chunk_slot = 0;                        // slot the next chunk goes to
    uint8_t* d_repack = nullptr; size_t repack_bytes = 0;   // aligned, zero-padded copy of an unaligned device clip
    uint32_t tune_stages = 0, tune_tile_px = 0, tune_segments = 0, tune_regs = 0;
    int tune_kernel = -1;                      // -1: automatic (clip_kernel_ws whenever the tuning allows it)
    uint32_t last_plan[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool timing = false;
    std::vector<cudaEvent_t> tev;              // start/stop pairs around clip kernel launches
    size_t tev_used = 0;
    std::vector<cudaEvent_t> pev;              // 4 per sharded pass: begin | exchanged | pass done | accumulators combined
    size_t pev_used = 0;
    // multi-GPU (comm.cu): communicator this context is a rank of, and what the last sharded pass left behind
    dipsb::Comm* comm = nullptr;
    bool acc_sharded = false;                  // the accumulators hold totals only inside this rank's owned range (reduce-scatter)
    uint64_t shard_total_frames = 0;           // frames of the whole clip of the last sharded pass (bounds the packed formats)
    uint64_t shard_first = 0, shard_n = 0;     // the frames of that clip this rank owns
    std::string err;
};

inline thread_local std::string g_create_err;   // last error of a call without a context (dipsb_create, utilities)

inline int32_t fail(dipsb_ctx* c, int32_t code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_err = buf;
    return code;
}

#define CK(c, call)                                                                                     \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return fail((c), DIPSB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

namespace dipsb {

// Extra trailing frame of a per-frame shard (the next shard's first frame, delivered into this rank's window by its owner):
// differenced after the call's last frame, scalars under index first + n.  flag == nullptr: the frame is already there.
struct ShardExtra {
    const uint8_t* frame = nullptr;
    const unsigned long long* flag = nullptr;
    unsigned long long epoch = 0, timeout_ns = 0;
    uint32_t* status = nullptr;
    // accumulator exchange fused into the last flush of the call's clip kernel (nranks == 0: none); *pushed tells the
    // caller whether the kernel took it (it cannot on re-packed, segmented or per-frame-kernel paths)
    ShardPush push;
    bool* pushed = nullptr;
    // leave the per-frame scalar finalisation of the call's (last) clip kernel to the caller (finalize_pending): the sharded
    // pass tells its peers that the accumulators are on their way before it spends 8-20 us on its own scalars
    bool defer_finalize = false;
};

// api.cu internals used by comm.cu
int32_t run_clip_on_stream(dipsb_ctx* c, const uint8_t* d_frames, uint64_t n, uint64_t stride, uint64_t first,
                           bool zero_padded, const ShardExtra* extra);
struct HostClipHooks {
    const ShardExtra* extra = nullptr;      // handed to the call's last chunk
    // called once, after the upload of the first chunk was enqueued on the copy stream (ordered behind it): the first
    // frame of the call then sits at d_first_frame
    int32_t (*after_first_upload)(dipsb_ctx* c, const uint8_t* d_first_frame, void* user) = nullptr;
    void* user = nullptr;
};
int32_t run_clip_host_impl(dipsb_ctx* c, const uint8_t* frames, uint64_t n, uint64_t stride, uint64_t first,
                           const HostClipHooks* hooks);
int32_t ensure_scalars(dipsb_ctx* c, uint64_t upto);
int32_t ensure_acc_zero(dipsb_ctx* c);
int32_t finalize_pending(dipsb_ctx* c);     // run a deferred finalize_scalars launch (no-op when none is pending)      // materialise a deferred zeroing of the accumulator planes
int bpp_of(int format);
int bit_length(uint64_t v);
// comm.cu: called by dipsb_destroy / geometry changes
void comm_detach(dipsb_ctx* c);

}  // namespace dipsb

