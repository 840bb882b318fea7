// clip_kernel.cu -- the hot kernel of libdips_b200: a whole clip (or frame-range shard) in one launch.
//
// Replaces the per-frame dispatch of compute_main (reference dips/src/gpu/shaders/dips_shader.wgsl:172-240 driven by
// dips/src/gpu/mod.rs:306-397 once per decoded frame) by one pass over the clip that reads every input byte exactly once.
//
// Design (B200 / sm_100a, HBM-bound, no tensor cores -- there is no contraction); measurements in DESIGN.md section 4:
//   * grid = (pixel tiles, frame segments).  A block owns a contiguous slice of tile_px <= 16*blockDim.x pixels of every
//     frame of its segment -- normally ONE block of 896 threads per SM, tile_px = npx / (148 * waves).  A thread owns 16
//     pixels; their reference/previous I2 (8 packed u16x2 registers) and the packed u16 accumulators (8 + 8 registers)
//     stay in registers across the frame loop, so the accumulators cost no HBM traffic per frame.
//   * per frame the block's slice (tile_px*bpp bytes, 40-56 KB for HD and above) arrives by ONE TMA bulk copy
//     (cp.async.bulk global->shared, completion on an mbarrier, L2 evict-first) into a ring of `stages` buffers; thread 0
//     re-arms the buffer released one iteration earlier, so stages-1 frames are in flight per block and no registers are
//     tied up by loads.  No __syncthreads in the loop: full/empty mbarriers only.
//   * threads read their 48 B (RGB8, stride 48 B) / 4x16 B (RGBx8, block-strided) from shared memory with conflict-free
//     128-bit loads, de-interleave with PRMT into u16x2 lanes, and use the packed DPX instructions (VIMNMX3.U16x2,
//     VIMNMX.U16x2, VIADDMNMX.S16x2.RELU) for max/min/threshold.  The integer ALU pipe is the second limiter after HBM,
//     so every add runs on the FMA pipe (IMAD with an opaque multiplier, IDP.2A for the per-frame sums).
//   * per-frame scalars: IDP.2A sums -> REDUX.SUM -> one 4-byte store per warp per frame (no atomics, no block barrier);
//     a small finalize kernel adds the per-warp words.
//   * every 128 frames (510*128 < 2^16) and at the end the packed accumulators are added to the u32 planes with
//     coalesced RED.ADD (the planes are kept in a tile order that makes thread-adjacent = address-adjacent).
//   * consecutive frames are serialised by a fake data dependency (mbar_wait_after): without it the compiler overlaps the
//     tail of one frame with the loads of the next and spills; with it the kernel needs 64-72 registers.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <algorithm>

#include "dipsb_internal.h"
#include "intensity.cuh"

namespace dipsb {
namespace {

// ---- PTX helpers -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// same, but ordered after the computation of `token` (a fake data dependency): keeps the compiler from overlapping the
// register-hungry tail of one frame with the loads of the next, which would cost ~16 more live registers per thread
__device__ __forceinline__ void mbar_wait_after(uint32_t bar, uint32_t parity, uint32_t token, uint32_t hint_ns) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        ".reg .b32 T;\n"
        "and.b32 T, %2, 0;\n"
        "add.u32 T, T, %0;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [T], %1, %3;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar),
        "r"(parity), "r"(token), "r"(hint_ns)
        : "memory");
}
// the same without a suspend-time hint (one register less in the warp-specialised kernel's loop)
__device__ __forceinline__ void mbar_wait_after(uint32_t bar, uint32_t parity, uint32_t token) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        ".reg .b32 T;\n"
        "and.b32 T, %2, 0;\n"
        "add.u32 T, T, %0;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [T], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar),
        "r"(parity), "r"(token)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier; frames are streamed once: evict-first in L2.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s_nohint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// Wait until a peer GPU (its copy engine, over NVLink) has raised *flag to `want`; the data it guards was written before
// the flag.  Bounded: a rank that never arrives must not hang this GPU (status := 1 and the pass carries on with garbage).
__device__ __forceinline__ void wait_peer_flag(const unsigned long long* flag, unsigned long long want, unsigned long long timeout_ns,
                                            uint32_t* status) {
    unsigned long long v, t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (uint32_t n = 1; flag != nullptr; ++n) {   // no flag: the frame is already there
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
        if (v >= want) break;
        __nanosleep(200);
        if ((n & 255u) == 0u) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > timeout_ns) { if (status) atomicExch(status, 1u); break; }
        }
    }
    asm volatile("fence.proxy.async.global;" ::: "memory");   // the frame is read through the async proxy (TMA)
}
// one lane of the (converged) warp, chosen by the hardware: no lane-id register has to stay live for "if (lane == 0)"
__device__ __forceinline__ bool elect_one() {
    uint32_t p;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(p));
    return p != 0u;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// ---- the kernel ----------------------------------------------------------------------------------------------------
struct KParams {
    const uint8_t* frames;
    uint64_t stride;
    uint64_t npx;
    const uint16_t* state_in;
    uint16_t* state_out;
    uint32_t* acc_sum;
    uint32_t* acc_cnt;
    uint32_t* partials;
    uint32_t n_frames;
    uint32_t n_segments;
    uint32_t tile_px;        // pixels per tile (multiple of 16, <= 16*blockDim.x)
    uint32_t stages;
    uint32_t stage_bytes;    // bytes reserved per stage: 16*blockDim.x*bpp rounded up to 128
    uint32_t words_per_frame;   // row pitch of the per-frame scratch: tiles*active_warps rounded up to 4 words
    uint32_t active_warps;   // warps per block that own pixels; the others leave after the setup
    uint32_t tau;
    uint32_t l2_evict_first; // 1: frames are fetched with an L2 evict-first policy (they are read exactly once)
    uint32_t wait_hint_ns;   // suspend-time hint of the consumers' mbarrier.try_wait
    uint32_t one;            // the constant 1 (an IMAD multiplier the compiler cannot fold away, see add_fma)
    // per-frame mode across GPUs: one extra trailing frame (the next shard's first frame, pushed into this rank's window by
    // its owner while this kernel runs) is differenced after frame n_frames-1 under scalar row n_frames; the producer waits
    // until *halo_flag >= halo_epoch before it fetches it
    const uint8_t* extra_frame;
    const unsigned long long* halo_flag;
    unsigned long long halo_epoch;
    unsigned long long wait_timeout_ns;
    uint32_t* status;        // set to 1 when that wait timed out (the results of this pass are then invalid)
    // frame-range shards: the LAST flush of the pass hands the elements this rank does not own straight to their owners --
    // total so far (local plane + registers), packed, stored into the owner's receive slot over NVLink -- so the accumulator
    // exchange needs no separate pass over the planes and overlaps the streaming of the tiles that are still running
    uint32_t first_store;    // clip_kernel_ws, one segment: dipsb_reset left the zeroing of the planes to this launch -- every thread
                             // clears its own 32 words while the TMA ring fills (no separate pass over 8 bytes per pixel)
    uint32_t xchg_nranks;    // 0: ordinary flush
    uint32_t xchg_own_lo, xchg_own_len, xchg_chunk;   // owned element range of this rank, elements per owner
    uint32_t xchg_fmt, xchg_sum_bits;                 // 1: sum | cnt << bits in one u32; 2: two u32 (cnt at +chunk)
    uint32_t* xchg_recv[kMaxRanks];                   // owner's receive slot for this rank, this pass
};

// a + b on the FMA pipe (IMAD with a multiplier the compiler cannot fold): the integer ALU pipe is the busier one here
__device__ __forceinline__ uint32_t add_fma(uint32_t a, uint32_t b, uint32_t one) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(b));
    return d;
}

// a * b + c that is executed where it stands (never hoisted out of a loop, never kept live across one)
__device__ __forceinline__ uint32_t mad_volatile(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t mad_fma(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// One frame for my pixels (N packed registers = 2N pixels): D = |cur-ref| per packed half, threshold, accumulate; returns
// sad | count<<20 of the thread.
//   hi = max(cur, ref)           VIMNMX.U16x2            (ALU pipe)
//   d  = 2*hi - (cur + ref)      2 x IMAD                (FMA pipe; cur+ref <= 1020 and d >= 0 per half: no carry/borrow)
//   m  = min(max(d - tau, 0), 1) VIADDMNMX.S16x2.RELU    (ALU pipe)
//   per-frame sums with IDP.2A (adds both halves into a scalar in one FMA-pipe instruction)
template <int N>
__device__ __forceinline__ uint32_t diff_px(const uint32_t* cur, const uint32_t* ref, uint32_t* accD, uint32_t* accM,
                                            uint32_t negtau2, uint32_t one) {
    uint32_t sD = 0u, sM = 0u;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const uint32_t hi = __vmaxu2(cur[j], ref[j]);
        const uint32_t sum = add_fma(cur[j], ref[j], one);
        const uint32_t d = mad_fma(hi, one + one, 0u - sum);
        const uint32_t m = __viaddmin_s16x2_relu(d, negtau2, 0x00010001u);
        accD[j] = add_fma(accD[j], d, one);
        accM[j] = add_fma(accM[j], m, one);
        sD = __dp2a_lo(d, 0x0101u, sD);
        sM = __dp2a_lo(m, 0x0101u, sM);
    }
    // sad <= 32*510 and cnt <= 32 per thread; summed over 32 lanes they still fit the 20-bit / 12-bit fields
    return sD + (sM << 20);
}

// G = groups of 16 pixels per thread (1 or 2).  G = 2 halves the per-warp, per-frame overhead (barrier wait, loop control,
// warp reduction) per pixel at the price of ~50 more registers per thread (fewer, fatter warps).
template <int BPP, int CH, int MODE, int G, int MAXREG>
__global__ void __maxnreg__(MAXREG) clip_kernel(const __grid_constant__ KParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int kWords = BPP * 4;  // 32-bit words of raw pixels per thread, frame and group
    constexpr int R = 8 * G;         // packed u16x2 registers per thread and plane
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t nthr = blockDim.x;
    const uint32_t tile = blockIdx.x, seg = blockIdx.y;
    const uint32_t S = P.stages;
    const uint32_t stage_bytes = P.stage_bytes;
    const uint32_t slots = nthr * (kPxPerThread * G);   // accumulator slots per tile (>= tile_px)

    // frame range of this segment, and the extra leading "prime" frame of per-frame mode
    const uint32_t t0 = (uint32_t)(((uint64_t)P.n_frames * seg) / P.n_segments);
    const uint32_t t1 = (uint32_t)(((uint64_t)P.n_frames * (seg + 1)) / P.n_segments);
    const bool prime_first = (MODE == 1) && (seg > 0);
    const uint32_t first = t0 - (prime_first ? 1u : 0u);
    const bool has_extra = (MODE == 1) && P.extra_frame != nullptr && seg == P.n_segments - 1;
    const uint32_t count = t1 - first + (has_extra ? 1u : 0u);

    const uint64_t tile_first_px = (uint64_t)tile * P.tile_px;
    const uint64_t remain_px = P.npx - tile_first_px;
    const uint32_t valid_px = remain_px < P.tile_px ? (uint32_t)remain_px : P.tile_px;
    // rounded up to the 16-byte granule of a bulk copy: only a frame whose size is not a multiple of 16 is affected, and
    // the host takes this kernel for such frames only from its own re-packed buffer, whose row padding is zero
    const uint32_t valid_bytes = (valid_px * BPP + 15u) & ~15u;

    const uint32_t smem_base = smem_u32(smem);
    const uint32_t full_bar = smem_base + S * stage_bytes;   // full[S] then empty[S], 8 bytes each
    const uint32_t empty_bar = full_bar + 8u * S;

    // Bytes past the tile's valid range must read as zero in every stage (TMA never writes them): the pixels they stand
    // for then have I2 = 0 against a zero reference and contribute nothing.
    if (valid_px < slots) {
        for (uint32_t o = tid * 16u; o < S * stage_bytes; o += nthr * 16u)
            *reinterpret_cast<uint4*>(smem + o) = make_uint4(0, 0, 0, 0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (tid == 0) {
        for (uint32_t s = 0; s < S; ++s) {
            mbar_init(full_bar + 8u * s, 1);
            mbar_init(empty_bar + 8u * s, P.active_warps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp >= P.active_warps) return;   // warps without pixels (tile_px small against the block); they never touch a barrier

    // ---- producer (thread 0): one TMA bulk copy per frame, re-arming the buffer freed one iteration ago ----------------
    // (the host only takes this kernel when base and stride are multiples of 16)
    const uint64_t tile_byte0 = tile_first_px * BPP;
    auto issue = [&](uint32_t it, uint32_t stage) {   // frame `first + it` of the call into buffer `stage`
        const uint8_t* src = P.frames + (uint64_t)(first + it) * P.stride + tile_byte0;
        if (has_extra && it + 1 == count) {
            wait_peer_flag(P.halo_flag, P.halo_epoch, P.wait_timeout_ns, P.status);
            src = P.extra_frame + tile_byte0;
        }
        mbar_arrive_expect_tx(full_bar + 8u * stage, valid_bytes);
        if (P.l2_evict_first) bulk_g2s(smem_base + stage * stage_bytes, src, valid_bytes, full_bar + 8u * stage, policy_evict_first());
        else bulk_g2s_nohint(smem_base + stage * stage_bytes, src, valid_bytes, full_bar + 8u * stage);
    };
    if (tid == 0) {
        const uint32_t pre = (S - 1 < count) ? S - 1 : count;
        for (uint32_t j = 0; j < pre; ++j) issue(j, j);
    }

    // ---- my pixels: reference / previous-frame I2 (zero for slots beyond the tile's pixels) ------------------------------
    // 3 B/px: group g = pixels 16*(g*nthr + tid) .. +15 of the tile (48 contiguous bytes, 128-bit shared loads at stride 48 B);
    // 4 B/px: quad q = pixels 4*(q*nthr + tid) .. +3, q = 0 .. 4G-1 (128-bit shared loads at stride 16 B).  Both conflict-free.
    uint32_t ra[R], rb[R];
    if constexpr (BPP == 3) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0;
            const uint32_t off = (g * nthr + tid) * kPxPerThread;
            if (off < valid_px) {
                const uint4* sp = reinterpret_cast<const uint4*>(P.state_in + tile_first_px + off);
                r0 = __ldg(sp); r1 = __ldg(sp + 1);
            }
            ra[8 * g + 0] = r0.x; ra[8 * g + 1] = r0.y; ra[8 * g + 2] = r0.z; ra[8 * g + 3] = r0.w;
            ra[8 * g + 4] = r1.x; ra[8 * g + 5] = r1.y; ra[8 * g + 6] = r1.z; ra[8 * g + 7] = r1.w;
        }
    } else {
#pragma unroll
        for (int q = 0; q < 4 * G; ++q) {
            uint2 r = make_uint2(0, 0);
            const uint32_t off = 4u * (q * nthr + tid);
            if (off < valid_px) r = __ldg(reinterpret_cast<const uint2*>(P.state_in + tile_first_px + off));
            ra[2 * q] = r.x; ra[2 * q + 1] = r.y;
        }
    }
    uint32_t accD[R], accM[R];
#pragma unroll
    for (int j = 0; j < R; ++j) accD[j] = accM[j] = 0u;

    const uint32_t tau = P.tau > 511u ? 511u : P.tau;
    const uint32_t negtau2 = ((0u - tau) & 0xFFFFu) * 0x00010001u;  // (-tau, -tau) as s16x2
    uint32_t part = first * P.words_per_frame + tile * P.active_warps + warp;   // index into P.partials (host checks < 2^32)
    const uint32_t my_smem = smem_base + (BPP == 3 ? tid * 48u : tid * 16u);
    const uint32_t grp_step = (BPP == 3 ? 48u : 64u) * nthr;   // bytes between my consecutive groups of 16 pixels
    const uint32_t my_step = (BPP == 3) ? 16u : nthr * 16u;    // bytes between the 128-bit pieces of one group

    auto flush = [&]() {
        uint32_t* const acc_sum = P.acc_sum + (uint64_t)tile * slots + tid;
        uint32_t* const acc_cnt = P.acc_cnt + (uint64_t)tile * slots + tid;
#pragma unroll
        for (int j = 0; j < R; ++j) {
            atomicAdd(acc_sum + (2 * j) * nthr, accD[j] & 0xFFFFu);
            atomicAdd(acc_sum + (2 * j + 1) * nthr, accD[j] >> 16);
            atomicAdd(acc_cnt + (2 * j) * nthr, accM[j] & 0xFFFFu);
            atomicAdd(acc_cnt + (2 * j + 1) * nthr, accM[j] >> 16);
            accD[j] = accM[j] = 0u;
        }
    };

    uint32_t stage = 0, parity = 0, iter = 0;
    // consume one frame: wait for its bytes, pull my pixels into registers, release the buffer, (thread 0) refill the
    // buffer released in the previous iteration, and turn the bytes into packed intensities
    auto fetch = [&](uint32_t* cur, uint32_t token) {
        mbar_wait_after(full_bar + 8u * stage, parity, token, P.wait_hint_ns);
        uint32_t w[G][kWords];
#pragma unroll
        for (int g = 0; g < G; ++g) {
#pragma unroll
            for (int v = 0; v < BPP; ++v) {
                const uint4 x = lds128(my_smem + stage * stage_bytes + grp_step * g + my_step * v);
                w[g][4 * v] = x.x; w[g][4 * v + 1] = x.y; w[g][4 * v + 2] = x.z; w[g][4 * v + 3] = x.w;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar + 8u * stage);
        if (tid == 0 && iter + S - 1 < count) {
            const uint32_t ps = stage == 0 ? S - 1 : stage - 1;   // == (iter + S - 1) % S
            if (iter > 0) mbar_wait(empty_bar + 8u * ps, stage == 0 ? parity ^ 1u : parity);
            issue(iter + S - 1, ps);
        }
        ++iter;
        if (++stage == S) { stage = 0; parity ^= 1u; }
#pragma unroll
        for (int g = 0; g < G; ++g) intensity16<BPP, CH>(w[g], cur + 8 * g);
    };
    auto emit = [&](uint32_t packed) {   // per-frame scalars: warp-reduce, one 4-byte store per warp per frame
        const uint32_t wsum = __reduce_add_sync(0xFFFFFFFFu, packed);
        if (lane == 0) P.partials[part] = wsum;
        part += P.words_per_frame;
        return wsum;
    };
    const uint32_t one = P.one;

    uint32_t token = 0;
    if (prime_first) {   // halo frame t0-1: only establishes the previous-frame plane
        fetch(ra, token);
        part += P.words_per_frame;
    }
    const uint32_t iter0 = iter;   // 1 after a halo frame, else 0
    // two frames per trip so that per-frame mode ping-pongs ra/rb without register moves
    while (iter + 2 <= count) {
        fetch(rb, token);
        token = emit(diff_px<R>(rb, ra, accD, accM, negtau2, one));
        if constexpr (MODE == 0) {
            fetch(rb, token);
            token = emit(diff_px<R>(rb, ra, accD, accM, negtau2, one));
        } else {
            fetch(ra, token);
            token = emit(diff_px<R>(ra, rb, accD, accM, negtau2, one));
        }
        if (((iter - iter0) & (uint32_t)(kFlushFrames - 1)) == 0u) flush();   // every 128 accumulated frames
    }
    if (iter < count) {
        fetch(rb, token);
        emit(diff_px<R>(rb, ra, accD, accM, negtau2, one));
        if constexpr (MODE == 1) {
#pragma unroll
            for (int j = 0; j < R; ++j) ra[j] = rb[j];
        }
    }
    flush();

    if (MODE == 1 && seg == P.n_segments - 1) {  // chain: I2 of the last frame becomes the next call's previous frame
        if constexpr (BPP == 3) {
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const uint32_t off = (g * nthr + tid) * kPxPerThread;
                if (off < valid_px) {
                    uint4* sp = reinterpret_cast<uint4*>(P.state_out + tile_first_px + off);
                    sp[0] = make_uint4(ra[8 * g + 0], ra[8 * g + 1], ra[8 * g + 2], ra[8 * g + 3]);
                    sp[1] = make_uint4(ra[8 * g + 4], ra[8 * g + 5], ra[8 * g + 6], ra[8 * g + 7]);
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4 * G; ++q) {
                const uint32_t off = 4u * (q * nthr + tid);
                if (off < valid_px)
                    *reinterpret_cast<uint2*>(P.state_out + tile_first_px + off) = make_uint2(ra[2 * q], ra[2 * q + 1]);
            }
        }
    }
}

// ---- warp-specialised variant ----------------------------------------------------------------------------------------
// Same arithmetic, tiles, accumulator order and barriers as clip_kernel, restructured to spend fewer instructions per frame
// (under the 1 kW power cap the SMs run at ~1.6 GHz and the RGB8 kernel becomes issue-bound, profiles/r01_sweeps.md):
//   * a dedicated producer warp (the last warp of the block) issues the TMA copies, so the consumer warps carry no
//     divergent "am I thread 0" branch and all S buffers are in flight;
//   * the frame loop is unrolled over the S pipeline stages (x2 for odd S, for the prev/cur ping-pong): stage index,
//     barrier addresses and mbarrier parities are compile-time, no per-frame stage bookkeeping;
//   * per-frame sums use 3-input adds on the packed lanes and one IDP.2A fold instead of an IDP.2A per register.
// The threshold counts of a flush interval (<= 128 frames) fit a byte: four pixels share one accumulator register, which
// leaves the warp-specialised kernel 4 registers more than one u16 per count would (it runs at the 64-register limit).
// accM4[i] holds pixels 4i .. 4i+3 of the thread as bytes 0, 2, 1, 3 (m[2i] + (m[2i+1] << 8)).
template <int N>
__device__ __forceinline__ uint32_t diff_px_ws(const uint32_t* cur, const uint32_t* ref, uint32_t* accD, uint32_t* accM4,
                                               uint32_t negtau2, uint32_t one) {
    uint32_t d[N], m[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const uint32_t hi = __vmaxu2(cur[j], ref[j]);
        const uint32_t sum = add_fma(cur[j], ref[j], one);
        d[j] = mad_fma(hi, one + one, 0u - sum);
        m[j] = __viaddmin_s16x2_relu(d[j], negtau2, 0x00010001u);
        accD[j] = add_fma(accD[j], d[j], one);
    }
#pragma unroll
    for (int i = 0; i < N / 2; ++i) accM4[i] = add_fma(accM4[i], m[2 * i + 1] * 256u + m[2 * i], one);
    uint32_t sD = 0u, sM = 0u;   // packed lane sums: <= 8*510 and <= 8 per lane
#pragma unroll
    for (int j = 0; j + 1 < N; j += 2) { sD = sD + d[j] + d[j + 1]; sM = sM + m[j] + m[j + 1]; }
    return __dp2a_lo(sD, 0x0101u, __dp2a_lo(sM, 0x0101u, 0u) << 20);
}
// count of the thread's pixel `px` (0..15) out of the byte-packed accumulators
__device__ __forceinline__ uint32_t count_of(const uint32_t* accM4, int px) {
    const int e = px & 3, byte = (e == 1) ? 2 : (e == 2) ? 1 : e;
    return (accM4[px >> 2] >> (8 * byte)) & 0xFFu;
}

template <int BPP, int CH, int MODE, int S>
__global__ void __maxnreg__(64) clip_kernel_ws(const __grid_constant__ KParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int kWords = BPP * 4;
    constexpr int R = 8;
    constexpr int U = (S % 2 == 0) ? S : 2 * S;                  // frames per trip
    constexpr uint32_t kFlushEvery = (uint32_t)(kFlushFrames / U) * U;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t nthr = blockDim.x - 32u;                      // consumer threads; the last warp is the producer
    const uint32_t tile = blockIdx.x, seg = blockIdx.y;
    const uint32_t stage_bytes = P.stage_bytes;
    const uint32_t slots = nthr * kPxPerThread;

    const uint32_t t0 = (uint32_t)(((uint64_t)P.n_frames * seg) / P.n_segments);
    const uint32_t t1 = (uint32_t)(((uint64_t)P.n_frames * (seg + 1)) / P.n_segments);
    const bool prime_first = (MODE == 1) && (seg > 0);
    const uint32_t first = t0 - (prime_first ? 1u : 0u);
    const bool has_extra = (MODE == 1) && P.extra_frame != nullptr && seg == P.n_segments - 1;   // see KParams
    const uint32_t count = t1 - first + (has_extra ? 1u : 0u);

    const uint64_t tile_first_px = (uint64_t)tile * P.tile_px;
    const uint64_t remain_px = P.npx - tile_first_px;
    const uint32_t valid_px = remain_px < P.tile_px ? (uint32_t)remain_px : P.tile_px;
    const uint32_t valid_bytes = (valid_px * BPP + 15u) & ~15u;   // see clip_kernel

    const uint32_t smem_base = smem_u32(smem);
    const uint32_t full_bar = smem_base + S * stage_bytes;
    const uint32_t empty_bar = full_bar + 8u * S;

    if (valid_px < slots) {
        for (uint32_t o = tid * 16u; o < S * stage_bytes; o += blockDim.x * 16u)
            *reinterpret_cast<uint4*>(smem + o) = make_uint4(0, 0, 0, 0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (tid == 0) {
        for (uint32_t s = 0; s < (uint32_t)S; ++s) {
            mbar_init(full_bar + 8u * s, 1);
            mbar_init(empty_bar + 8u * s, P.active_warps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == nthr / 32u) {   // ---- producer warp: one TMA bulk copy per frame, S frames in flight ----
        if (lane == 0) {
            const uint8_t* src = P.frames + (uint64_t)first * P.stride + tile_first_px * BPP;
            uint32_t stage = 0, par = 1;   // the wait on a never-used buffer falls through (parity of the preceding phase)
            for (uint32_t it = 0; it < count; ++it) {
                mbar_wait(empty_bar + 8u * stage, par);
                if (has_extra && it + 1 == count) {   // the next shard's first frame, delivered into this rank's window
                    wait_peer_flag(P.halo_flag, P.halo_epoch, P.wait_timeout_ns, P.status);
                    src = P.extra_frame + tile_first_px * BPP;
                }
                mbar_arrive_expect_tx(full_bar + 8u * stage, valid_bytes);
                if (P.l2_evict_first) bulk_g2s(smem_base + stage * stage_bytes, src, valid_bytes, full_bar + 8u * stage, policy_evict_first());
                else bulk_g2s_nohint(smem_base + stage * stage_bytes, src, valid_bytes, full_bar + 8u * stage);
                src += P.stride;
                if (++stage == (uint32_t)S) { stage = 0; par ^= 1u; }
            }
        }
        return;
    }
    if (P.first_store) {   // my accumulator words start at zero (same thread, same addresses as the flushes that follow)
        uint32_t* const acc_sum = P.acc_sum + (uint64_t)tile * slots + tid;
        uint32_t* const acc_cnt = P.acc_cnt + (uint64_t)tile * slots + tid;
#pragma unroll
        for (int k = 0; k < 2 * R; ++k) { acc_sum[k * nthr] = 0u; acc_cnt[k * nthr] = 0u; }
    }
    if (warp >= P.active_warps) return;

    uint32_t ra[R], rb[R];
    if constexpr (BPP == 3) {
        uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0;
        const uint32_t off = tid * kPxPerThread;
        if (off < valid_px) {
            const uint4* sp = reinterpret_cast<const uint4*>(P.state_in + tile_first_px + off);
            r0 = __ldg(sp); r1 = __ldg(sp + 1);
        }
        ra[0] = r0.x; ra[1] = r0.y; ra[2] = r0.z; ra[3] = r0.w;
        ra[4] = r1.x; ra[5] = r1.y; ra[6] = r1.z; ra[7] = r1.w;
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint2 r = make_uint2(0, 0);
            const uint32_t off = 4u * (q * nthr + tid);
            if (off < valid_px) r = __ldg(reinterpret_cast<const uint2*>(P.state_in + tile_first_px + off));
            ra[2 * q] = r.x; ra[2 * q + 1] = r.y;
        }
    }
    uint32_t accD[R], accM[R / 2];   // sums as u16 pairs, counts as bytes (see diff_px_ws)
#pragma unroll
    for (int j = 0; j < R; ++j) accD[j] = 0u;
#pragma unroll
    for (int j = 0; j < R / 2; ++j) accM[j] = 0u;

    const uint32_t tau = P.tau > 511u ? 511u : P.tau;
    const uint32_t negtau2 = ((0u - tau) & 0xFFFFu) * 0x00010001u;
    uint32_t part = first * P.words_per_frame + tile * P.active_warps + warp;
    const uint32_t my_smem = smem_base + (BPP == 3 ? tid * 48u : tid * 16u);
    const uint32_t my_step = (BPP == 3) ? 16u : nthr * 16u;
    const uint32_t one = P.one;

    auto flush = [&]() {
        uint32_t* const acc_sum = P.acc_sum + (uint64_t)tile * slots + tid;
        uint32_t* const acc_cnt = P.acc_cnt + (uint64_t)tile * slots + tid;
#pragma unroll
        for (int j = 0; j < R; ++j) {
            atomicAdd(acc_sum + (2 * j) * nthr, accD[j] & 0xFFFFu);
            atomicAdd(acc_sum + (2 * j + 1) * nthr, accD[j] >> 16);
            atomicAdd(acc_cnt + (2 * j) * nthr, count_of(accM, 2 * j));
            atomicAdd(acc_cnt + (2 * j + 1) * nthr, count_of(accM, 2 * j + 1));
            accD[j] = 0u;
        }
#pragma unroll
        for (int j = 0; j < R / 2; ++j) accM[j] = 0u;
    };
    // last flush of a sharded pass (KParams::xchg_*): owned elements as above; every other element's total goes to its owner
    auto flush_to_owners = [&]() {
        const uint32_t base = tile * slots + tid;          // the host takes this path only for planes of < 2^32 elements
#pragma unroll
        for (int k = 0; k < 2 * R; ++k) {
            const uint32_t idx = base + (uint32_t)k * nthr;
            const uint32_t d = (k & 1) ? accD[k / 2] >> 16 : accD[k / 2] & 0xFFFFu;
            const uint32_t m = count_of(accM, k);
            if (idx - P.xchg_own_lo < P.xchg_own_len) {
                atomicAdd(P.acc_sum + idx, d);
                atomicAdd(P.acc_cnt + idx, m);
            } else {
                // atomics with return: ordered behind this thread's earlier RED.ADDs to the same words
                const uint32_t sum = atomicAdd(P.acc_sum + idx, d) + d, cnt = atomicAdd(P.acc_cnt + idx, m) + m;
                const uint32_t owner = idx / P.xchg_chunk, off = idx - owner * P.xchg_chunk;
                uint32_t* slot = P.xchg_recv[owner];
                if (P.xchg_fmt == 1u) slot[off] = sum | (cnt << P.xchg_sum_bits);
                else { slot[off] = sum; slot[P.xchg_chunk + off] = cnt; }
            }
        }
    };
    // bytes -> packed intensities of one frame sitting in buffer `stage` (compile-time or run-time index)
    auto fetch = [&](uint32_t stage, uint32_t par, uint32_t* cur, uint32_t token) {
        mbar_wait_after(full_bar + 8u * stage, par, token);
        if constexpr (BPP == 4) {
            // convert piece by piece: only 4 raw words are live next to the state / accumulator registers
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const uint4 x = lds128(my_smem + stage * stage_bytes + my_step * v);
                intensity4_x<CH>(x, cur[2 * v], cur[2 * v + 1]);
            }
            __syncwarp();
            if (elect_one()) mbar_arrive(empty_bar + 8u * stage);
        } else {
            uint32_t w[kWords];
#pragma unroll
            for (int v = 0; v < BPP; ++v) {
                const uint4 x = lds128(my_smem + stage * stage_bytes + my_step * v);
                w[4 * v] = x.x; w[4 * v + 1] = x.y; w[4 * v + 2] = x.z; w[4 * v + 3] = x.w;
            }
            __syncwarp();
            if (elect_one()) mbar_arrive(empty_bar + 8u * stage);
            intensity16<BPP, CH>(w, cur);
        }
    };
    auto emit = [&](uint32_t packed) {
        const uint32_t wsum = __reduce_add_sync(0xFFFFFFFFu, packed);
        if (elect_one()) P.partials[part] = wsum;
        part += P.words_per_frame;
        return wsum;
    };

    uint32_t iter = 0, token = 0, since_flush = 0;
    uint32_t rs = 0, par = 0;   // run-time stage index / parity of the lead-in and tail (the main loop starts every trip at stage 0)
    // ---- lead-in with run-time stage index: the halo frame of a per-frame segment, then up to the next trip boundary
    if (prime_first) {
        fetch(0u, 0u, ra, token);
        part += P.words_per_frame;
        iter = 1; rs = 1;
        if (rs == (uint32_t)S) { rs = 0; par ^= 1u; }
        while (iter % U != 0 && iter < count) {
            fetch(rs, par, rb, token);
            token = emit(diff_px_ws<R>(rb, ra, accD, accM, negtau2, one));
#pragma unroll
            for (int j = 0; j < R; ++j) ra[j] = rb[j];
            ++iter; ++since_flush;
            if (++rs == (uint32_t)S) { rs = 0; par ^= 1u; }
        }
    }
    // ---- main loop: U frames per trip, compile-time stages; iter is a multiple of U (hence rs == 0) here
    while (iter + U <= count) {
#pragma unroll
        for (int u = 0; u < U; u += 2) {
            fetch((uint32_t)(u % S), par ^ (uint32_t)((u / S) & 1), rb, token);
            token = emit(diff_px_ws<R>(rb, ra, accD, accM, negtau2, one));
            if constexpr (MODE == 0) {
                fetch((uint32_t)((u + 1) % S), par ^ (uint32_t)(((u + 1) / S) & 1), rb, token);
                token = emit(diff_px_ws<R>(rb, ra, accD, accM, negtau2, one));
            } else {
                fetch((uint32_t)((u + 1) % S), par ^ (uint32_t)(((u + 1) / S) & 1), ra, token);
                token = emit(diff_px_ws<R>(ra, rb, accD, accM, negtau2, one));
            }
        }
        iter += U;
        par ^= (uint32_t)((U / S) & 1);
        since_flush += U;
        if (since_flush + U > (uint32_t)kFlushFrames) { flush(); since_flush = 0; }
    }
    // ---- tail with run-time stage index (starts at stage 0)
    while (iter < count) {
        fetch(rs, par, rb, token);
        token = emit(diff_px_ws<R>(rb, ra, accD, accM, negtau2, one));
        if constexpr (MODE == 1) {
#pragma unroll
            for (int j = 0; j < R; ++j) ra[j] = rb[j];
        }
        ++iter;
        if (++rs == (uint32_t)S) { rs = 0; par ^= 1u; }
        if (++since_flush >= (uint32_t)kFlushFrames) { flush(); since_flush = 0; }
    }
    if (P.xchg_nranks) flush_to_owners(); else flush();
    (void)kFlushEvery;

    if (MODE == 1 && seg == P.n_segments - 1) {
        if constexpr (BPP == 3) {
            const uint32_t off = tid * kPxPerThread;
            if (off < valid_px) {
                uint4* sp = reinterpret_cast<uint4*>(P.state_out + tile_first_px + off);
                sp[0] = make_uint4(ra[0], ra[1], ra[2], ra[3]);
                sp[1] = make_uint4(ra[4], ra[5], ra[6], ra[7]);
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t off = 4u * (q * nthr + tid);
                if (off < valid_px)
                    *reinterpret_cast<uint2*>(P.state_out + tile_first_px + off) = make_uint2(ra[2 * q], ra[2 * q + 1]);
            }
        }
    }
}

// Roofline probe: the clip kernel's memory side only -- same tiles, same TMA ring, same barriers, but the consumers just
// release each buffer without reading it.  Its bandwidth is the ceiling of this access pattern on this GPU.
__global__ void stream_probe_kernel(const __grid_constant__ KParams P, int bpp) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const uint32_t S = P.stages, stage_bytes = P.stage_bytes, nwarps = blockDim.x >> 5;
    const uint64_t tile_first_px = (uint64_t)blockIdx.x * P.tile_px;
    const uint64_t remain_px = P.npx - tile_first_px;
    const uint32_t valid_bytes = ((remain_px < P.tile_px ? (uint32_t)remain_px : P.tile_px) * (uint32_t)bpp + 15u) & ~15u;   // as the clip kernels
    const uint32_t smem_base = smem_u32(smem), full_bar = smem_base + S * stage_bytes, empty_bar = full_bar + 8u * S;
    if (tid == 0) {
        for (uint32_t s = 0; s < S; ++s) { mbar_init(full_bar + 8u * s, 1); mbar_init(empty_bar + 8u * s, nwarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t count = P.n_frames;
    auto issue = [&](uint32_t it, uint32_t stage) {
        const uint8_t* src = P.frames + (uint64_t)it * P.stride + tile_first_px * bpp;
        mbar_arrive_expect_tx(full_bar + 8u * stage, valid_bytes);
        bulk_g2s(smem_base + stage * stage_bytes, src, valid_bytes, full_bar + 8u * stage, policy_evict_first());
    };
    if (tid == 0) for (uint32_t j = 0; j < (S - 1 < count ? S - 1 : count); ++j) issue(j, j);
    uint32_t stage = 0, parity = 0;
    for (uint32_t iter = 0; iter < count; ++iter) {
        mbar_wait(full_bar + 8u * stage, parity);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar + 8u * stage);
        if (tid == 0 && iter + S - 1 < count) {
            const uint32_t ps = stage == 0 ? S - 1 : stage - 1;
            if (iter > 0) mbar_wait(empty_bar + 8u * ps, stage == 0 ? parity ^ 1u : parity);
            issue(iter + S - 1, ps);
        }
        if (++stage == S) { stage = 0; parity ^= 1u; }
    }
}

template <int BPP, int CH, int MODE, int G, int MAXREG>
cudaError_t launch_r(const Geometry& g, const ClipArgs& a, const KParams& kp, size_t smem, cudaStream_t s) {
    auto kfn = clip_kernel<BPP, CH, MODE, G, MAXREG>;
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // the kernel uses no L1-cached global loads in its loop: give the whole unified array to shared memory, so that the
    // residency the planner assumed (blocks_per_sm) is what the hardware grants
    e = cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    dim3 grid(g.n_tiles, a.n_segments, 1), block(g.threads, 1, 1);
    kfn<<<grid, block, smem, s>>>(kp);
    count_launch();
    return cudaGetLastError();
}

template <int BPP, int CH, int MODE>
cudaError_t launch_t(const Geometry& g, const ClipArgs& a, const KParams& kp, size_t smem, cudaStream_t s) {
    switch (g.regs) {   // kernel variant chosen by the planner: more registers <-> fewer resident warps
        case 128: return launch_r<BPP, CH, MODE, 2, 128>(g, a, kp, smem, s);   // 32 pixels per thread
        case 96: return launch_r<BPP, CH, MODE, 1, 96>(g, a, kp, smem, s);
        case 80: return launch_r<BPP, CH, MODE, 1, 80>(g, a, kp, smem, s);
        case 72: return launch_r<BPP, CH, MODE, 1, 72>(g, a, kp, smem, s);
        default: return launch_r<BPP, CH, MODE, 1, 64>(g, a, kp, smem, s);
    }
}

template <int BPP, int CH, int MODE, int S>
cudaError_t launch_ws(const Geometry& g, const ClipArgs& a, const KParams& kp, size_t smem, cudaStream_t s) {
    auto kfn = clip_kernel_ws<BPP, CH, MODE, S>;
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    dim3 grid(g.n_tiles, a.n_segments, 1), block(g.threads + 32, 1, 1);   // + the producer warp
    kfn<<<grid, block, smem, s>>>(kp);
    count_launch();
    return cudaGetLastError();
}

template <int BPP, int CH>
cudaError_t launch_m(const Geometry& g, const ClipArgs& a, const KParams& kp, size_t smem, cudaStream_t s) {
    if (g.kernel == 1) {   // warp-specialised variant: 16 px/thread, 64 registers, 3 or 4 compile-time stages
        if (g.stages != 3 && g.stages != 4) return cudaErrorInvalidValue;
        if (a.mode == 0) return g.stages == 4 ? launch_ws<BPP, CH, 0, 4>(g, a, kp, smem, s) : launch_ws<BPP, CH, 0, 3>(g, a, kp, smem, s);
        // 4 B/px + chroma filter + per-frame mode does not fit 64 registers without spilling: not instantiated, the planner
        // gives such contexts clip_kernel (72 registers) -- clip_ws_available()
        if constexpr (BPP == 4 && CH >= 0) return cudaErrorInvalidValue;
        else return g.stages == 4 ? launch_ws<BPP, CH, 1, 4>(g, a, kp, smem, s) : launch_ws<BPP, CH, 1, 3>(g, a, kp, smem, s);
    }
    return a.mode == 0 ? launch_t<BPP, CH, 0>(g, a, kp, smem, s) : launch_t<BPP, CH, 1>(g, a, kp, smem, s);
}

template <int BPP>
cudaError_t launch_c(const Geometry& g, const ClipArgs& a, const KParams& kp, size_t smem, cudaStream_t s) {
    switch (g.chan_byte) {
        case 0: return launch_m<BPP, 0>(g, a, kp, smem, s);
        case 1: return launch_m<BPP, 1>(g, a, kp, smem, s);
        case 2: return launch_m<BPP, 2>(g, a, kp, smem, s);
        default: return launch_m<BPP, -1>(g, a, kp, smem, s);
    }
}

inline uint32_t stage_bytes_of(uint32_t threads, int bpp, int groups) { return (threads * kPxPerThread * (uint32_t)groups * (uint32_t)bpp + 127u) & ~127u; }

}  // namespace

bool clip_ws_available(int bpp, int chan_byte, int mode) { return !(bpp == 4 && chan_byte >= 0 && mode == 1); }
bool clip_can_store_first(const Geometry& g, uint32_t n_segments) { return g.kernel == 1 && n_segments == 1; }
bool clip_can_push(const Geometry& g, uint32_t n_segments) { return g.kernel == 1 && n_segments == 1 && g.n_elems < (1ull << 32); }

int clip_groups(int regs) { return regs >= 128 ? 2 : 1; }

size_t clip_smem_bytes(uint32_t threads, int bpp, uint32_t stages, int regs) {
    return (size_t)stages * stage_bytes_of(threads, bpp, clip_groups(regs)) + 16u * stages;  // buffers + full/empty mbarriers
}

// registers are allocated per warp in units of 256: warps/SM = floor(65536 / (regs*32)), e.g. 72 -> 28, 80 -> 25, 96 -> 21
int clip_max_threads_per_sm(int regs) { return (65536 / (regs * 32)) * 32; }

int clip_occupancy(uint32_t threads, int bpp, uint32_t stages, int regs) {
    const size_t smem = clip_smem_bytes(threads, bpp, stages, regs) + 1024;  // + per-block reservation
    const size_t smem_sm = 227 * 1024;
    if (smem > smem_sm) return 0;
    const int by_smem = (int)(smem_sm / smem);
    const int by_thr = clip_max_threads_per_sm(regs) / (int)threads;
    int occ = by_smem < by_thr ? by_smem : by_thr;
    return occ > 32 ? 32 : occ;
}

uint32_t clip_active_warps(const Geometry& g) {
    // 3 B/px: warp w owns pixels from 512*w upwards in its first group; 4 B/px: every warp owns pixels of every quad row
    const uint32_t warps = g.threads / 32u;
    return g.bpp == 3 ? std::min(warps, (g.tile_px + 511u) / 512u) : warps;
}

cudaError_t launch_stream_probe(const Geometry& g, const uint8_t* frames, uint64_t stride, uint32_t n_frames, cudaStream_t s) {
    KParams kp{};
    kp.frames = frames; kp.stride = stride; kp.npx = g.npx; kp.n_frames = n_frames;
    kp.tile_px = g.tile_px; kp.stages = g.stages; kp.stage_bytes = stage_bytes_of(g.threads, g.bpp, clip_groups(g.regs));
    const size_t smem = clip_smem_bytes(g.threads, g.bpp, g.stages, g.regs);
    cudaError_t e = cudaFuncSetAttribute(stream_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    stream_probe_kernel<<<g.n_tiles, 128, smem, s>>>(kp, g.bpp);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_clip(const Geometry& g, const ClipArgs& a, cudaStream_t s) {
    KParams kp;
    kp.frames = a.frames; kp.stride = a.stride; kp.npx = g.npx;
    kp.state_in = a.state_in; kp.state_out = a.state_out;
    kp.acc_sum = a.acc_sum; kp.acc_cnt = a.acc_cnt; kp.partials = a.partials;
    kp.n_frames = a.n_frames; kp.n_segments = a.n_segments;
    kp.tile_px = g.tile_px; kp.stages = g.stages; kp.stage_bytes = stage_bytes_of(g.threads, g.bpp, clip_groups(g.regs));
    kp.active_warps = clip_active_warps(g);
    kp.words_per_frame = (g.n_tiles * kp.active_warps + 3u) & ~3u;
    kp.tau = a.tau;
    static const int l2_hint = [] { const char* e = getenv("DIPSB_L2_EVICT_FIRST"); return e ? atoi(e) : 1; }();
    kp.l2_evict_first = (uint32_t)l2_hint;
    static const int wait_hint = [] { const char* e = getenv("DIPSB_WAIT_HINT_NS"); return e ? atoi(e) : 0; }();
    kp.wait_hint_ns = (uint32_t)wait_hint;
    kp.one = 1u;
    kp.extra_frame = a.extra_frame; kp.halo_flag = a.halo_flag; kp.halo_epoch = a.halo_epoch;
    kp.wait_timeout_ns = a.wait_timeout_ns; kp.status = a.status;
    kp.xchg_nranks = 0;
    kp.first_store = (a.first_store && clip_can_store_first(g, a.n_segments)) ? 1u : 0u;
    if (a.push && a.push->nranks && clip_can_push(g, a.n_segments)) {
        const ShardPush& x = *a.push;
        kp.xchg_nranks = x.nranks; kp.xchg_chunk = (uint32_t)x.chunk;
        kp.xchg_own_lo = (uint32_t)((uint64_t)x.rank * x.chunk);
        kp.xchg_own_len = (uint32_t)(std::min<uint64_t>((uint64_t)(x.rank + 1) * x.chunk, g.n_elems) - kp.xchg_own_lo);
        kp.xchg_fmt = (uint32_t)x.fmt; kp.xchg_sum_bits = (uint32_t)x.sum_bits;
        for (uint32_t r = 0; r < x.nranks; ++r) kp.xchg_recv[r] = x.recv[r];
    }
    const size_t smem = clip_smem_bytes(g.threads, g.bpp, g.stages, g.regs);
    return g.bpp == 3 ? launch_c<3>(g, a, kp, smem, s) : launch_c<4>(g, a, kp, smem, s);
}

}  // namespace dipsb
