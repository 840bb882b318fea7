// clip_kernel.cu -- the hot kernel of libdips_b200: a whole clip (or frame-range shard) in one launch.
//
// Replaces the per-frame dispatch of compute_main (reference dips/src/gpu/shaders/dips_shader.wgsl:172-240 driven by
// dips/src/gpu/mod.rs:306-397 once per decoded frame) by one pass over the clip that reads every input byte exactly once.
//
// Design (B200 / sm_100a, HBM-bound, no tensor cores -- there is no contraction):
//   * grid = (pixel tiles, frame segments).  A block owns tile_px = 16*blockDim.x pixels for its whole frame segment:
//     thread i owns 16 consecutive pixels; their reference/previous I2 (8 packed u16x2 registers) and the packed u16
//     accumulators (8 + 8 registers) live in registers across the frame loop, so the accumulators cost no HBM traffic
//     per frame.
//   * per frame the block's contiguous byte range of that frame (tile_px*bpp bytes, 6-24 KB) is brought in by ONE
//     TMA bulk copy (cp.async.bulk global->shared, completion on an mbarrier) into a ring of `stages` buffers;
//     thread 0 re-arms the buffer released in the previous iteration, so stages-1 frames are always in flight per
//     block with no registers tied up by loads.
//   * threads read their 48 B (RGB8) / 64 B (RGBx8) from shared memory with conflict-free 128-bit loads, de-interleave
//     with PRMT into u16x2 lanes, and use the packed DPX instructions (VIMNMX3.U16x2, VIADDMNMX.S16x2.RELU) for
//     max/min/threshold: ~4.5 ALU-pipe + ~2.5 FMA-pipe instructions per pixel.
//   * per-frame scalars: packed per-thread sums -> IDP.2A fold -> REDUX.SUM -> one 4-byte store per warp per frame
//     (no atomics, no block barrier in the frame loop); a small finalize kernel adds the per-warp words.
//   * every 128 frames (510*128 < 2^16) and at the end the packed accumulators are added to the u32 planes with
//     coalesced RED.ADD (the planes are kept in a tile order that makes thread-adjacent = address-adjacent).
#include <cuda_runtime.h>
#include <stdint.h>

#include "dipsb_internal.h"

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
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// ---- de-interleave + intensity ---------------------------------------------------------------------------------------
constexpr uint32_t kLoMask = 0x00FF00FFu;

// selector picking byte q1 of x into byte 0 and byte q2 of x into byte 2, zeros (from y == 0) elsewhere
__host__ __device__ constexpr uint32_t sel_same(int q1, int q2) { return (uint32_t)(q1 | (4 << 4) | (q2 << 8) | (4 << 12)); }
// selector picking byte q1 of x into byte 0 and byte q2 of y into byte 2 (other bytes arbitrary, masked later)
__host__ __device__ constexpr uint32_t sel_cross(int q1, int q2) { return (uint32_t)(q1 | ((4 + q2) << 8)); }

// packed u16x2 (lo = byte at position Q1, hi = byte at position Q2) out of a run of words w[]
template <int Q1, int Q2>
__device__ __forceinline__ uint32_t pair_bytes(const uint32_t* w) {
    if constexpr (Q1 / 4 == Q2 / 4) return __byte_perm(w[Q1 / 4], 0u, sel_same(Q1 % 4, Q2 % 4));
    else return __byte_perm(w[Q1 / 4], w[Q2 / 4], sel_cross(Q1 % 4, Q2 % 4)) & kLoMask;
}

// I2 of 16 pixels as 8 packed u16x2 registers (pixel 2j in the low half of I[j], pixel 2j+1 in the high half).
// CH < 0: max+min over the three colour bytes (get_intensity, dips_shader.wgsl:64-82, x510); CH >= 0: 2 * byte CH.
template <int BPP, int CH>
__device__ __forceinline__ void intensity16(const uint32_t* w, uint32_t* I) {
    if constexpr (BPP == 3) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {  // 4 pixels = 12 bytes = words a,b,c:  a: r0 g0 b0 r1 | b: g1 b1 r2 g2 | c: b2 r3 g3 b3
            const uint32_t* q = w + 3 * g;
            if constexpr (CH < 0) {
                const uint32_t a = q[0], b = q[1], c = q[2];
                const uint32_t x01 = __byte_perm(a, 0u, 0x4340);  // (a0, a3)
                const uint32_t t = __byte_perm(a, b, 0x5421);     // a1 a2 b0 b1
                const uint32_t y01 = __byte_perm(t, 0u, 0x4240);  // (a1, b0)
                const uint32_t z01 = __byte_perm(t, 0u, 0x4341);  // (a2, b1)
                const uint32_t u = __byte_perm(b, c, 0x6532);     // b2 b3 c1 c2
                const uint32_t x23 = __byte_perm(u, 0u, 0x4240);  // (b2, c1)
                const uint32_t y23 = __byte_perm(u, 0u, 0x4341);  // (b3, c2)
                const uint32_t z23 = __byte_perm(c, 0u, 0x4340);  // (c0, c3)
                I[2 * g] = __vimax3_u16x2(x01, y01, z01) + __vimin3_u16x2(x01, y01, z01);
                I[2 * g + 1] = __vimax3_u16x2(x23, y23, z23) + __vimin3_u16x2(x23, y23, z23);
            } else {
                const uint32_t p01 = pair_bytes<CH, CH + 3>(q);
                const uint32_t p23 = pair_bytes<CH + 6, CH + 9>(q);
                I[2 * g] = p01 + p01;
                I[2 * g + 1] = p23 + p23;
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {  // 2 pixels = words a (px 2j), b (px 2j+1)
            const uint32_t a = w[2 * j], b = w[2 * j + 1];
            if constexpr (CH < 0) {
                const uint32_t t = __byte_perm(a, b, 0x5410);  // a0 a1 b0 b1
                const uint32_t x = __byte_perm(t, 0u, 0x4240);
                const uint32_t y = __byte_perm(t, 0u, 0x4341);
                const uint32_t z = __byte_perm(a, b, 0x0602) & kLoMask;  // (a2, b2)
                I[j] = __vimax3_u16x2(x, y, z) + __vimin3_u16x2(x, y, z);
            } else {
                const uint32_t p = __byte_perm(a, b, sel_cross(CH, CH)) & kLoMask;
                I[j] = p + p;
            }
        }
    }
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
    uint32_t tile_px;
    uint32_t stages;
    uint32_t stage_bytes;   // bytes reserved per stage (tile_px*bpp rounded up to 128)
    uint32_t words_per_frame;
    uint32_t tau;
};

template <int BPP, int CH, int MODE>
__global__ void __maxnreg__(64) clip_kernel(const KParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int kWords = BPP * 4;  // 32-bit words of raw pixels per thread per frame
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t nthr = blockDim.x, nwarps = nthr >> 5;
    const uint32_t tile = blockIdx.x, seg = blockIdx.y;
    const uint32_t S = P.stages;

    // frame range of this segment, and the extra leading "prime" frame of per-frame mode
    const uint32_t t0 = (uint32_t)(((uint64_t)P.n_frames * seg) / P.n_segments);
    const uint32_t t1 = (uint32_t)(((uint64_t)P.n_frames * (seg + 1)) / P.n_segments);
    const bool prime_first = (MODE == 1) && (seg > 0);
    const uint32_t first = t0 - (prime_first ? 1u : 0u);
    const uint32_t count = t1 - first;

    const uint64_t tile_first_px = (uint64_t)tile * P.tile_px;
    const uint64_t remain_px = P.npx - tile_first_px;
    const uint32_t valid_px = remain_px < P.tile_px ? (uint32_t)remain_px : P.tile_px;
    const uint32_t valid_bytes = valid_px * BPP;
    const uint32_t bulk_bytes = valid_bytes & ~15u, tail_bytes = valid_bytes & 15u;
    const uint32_t tile_bytes = P.tile_px * BPP;

    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_base + S * P.stage_bytes;  // full[S] then empty[S], 8 bytes each
    auto full_bar = [&](uint32_t s) { return bar_base + 8u * s; };
    auto empty_bar = [&](uint32_t s) { return bar_base + 8u * (S + s); };

    if (valid_px < P.tile_px) {  // partial last tile: bytes past the valid range must read as zero
        for (uint32_t o = tid * 16u; o < S * P.stage_bytes; o += nthr * 16u)
            *reinterpret_cast<uint4*>(smem + o) = make_uint4(0, 0, 0, 0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (tid == 0) {
        for (uint32_t s = 0; s < S; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), nwarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint8_t* src0 = P.frames + (uint64_t)first * P.stride + tile_first_px * BPP;
    uint64_t policy = 0;
    if (tid == 0) policy = policy_evict_first();

    auto issue = [&](uint32_t iter, uint32_t stage) {  // thread 0 only
        const uint8_t* src = src0 + (uint64_t)iter * P.stride;
        const uint32_t dst = smem_base + stage * P.stage_bytes;
        if (tail_bytes) {
            for (uint32_t k = 0; k < tail_bytes; ++k) smem[stage * P.stage_bytes + bulk_bytes + k] = src[bulk_bytes + k];
        }
        mbar_arrive_expect_tx(full_bar(stage), bulk_bytes);
        if (bulk_bytes) bulk_g2s(dst, src, bulk_bytes, full_bar(stage), policy);
    };

    if (tid == 0) {
        const uint32_t pre = (S - 1 < count) ? S - 1 : count;
        for (uint32_t j = 0; j < pre; ++j) issue(j, j);
    }

    // reference / previous-frame I2 of my 16 pixels (state planes are padded with zeros up to n_tiles*tile_px).
    // 3 B/px: pixels tid*16 .. +15 of the tile; 4 B/px: groups of 4 pixels at 4*(v*nthr + tid), v = 0..3.
    uint32_t ref[8];
    if constexpr (BPP == 3) {
        const uint4* sp = reinterpret_cast<const uint4*>(P.state_in + tile_first_px + (uint64_t)tid * kPxPerThread);
        const uint4 r0 = __ldg(sp), r1 = __ldg(sp + 1);
        ref[0] = r0.x; ref[1] = r0.y; ref[2] = r0.z; ref[3] = r0.w;
        ref[4] = r1.x; ref[5] = r1.y; ref[6] = r1.z; ref[7] = r1.w;
    } else {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const uint2 r = __ldg(reinterpret_cast<const uint2*>(P.state_in + tile_first_px + 4u * (v * nthr + tid)));
            ref[2 * v] = r.x; ref[2 * v + 1] = r.y;
        }
    }
    uint32_t accD[8], accM[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) accD[j] = accM[j] = 0u;

    const uint32_t tau = P.tau > 511u ? 511u : P.tau;
    const uint32_t negtau2 = ((0u - tau) & 0xFFFFu) * 0x00010001u;  // (-tau, -tau) as s16x2
    const uint32_t one2 = 0x00010001u;
    uint32_t* const acc_sum = P.acc_sum + tile_first_px + tid;
    uint32_t* const acc_cnt = P.acc_cnt + tile_first_px + tid;
    uint32_t* part = P.partials + (uint64_t)first * P.words_per_frame + tile * nwarps + warp;
    // 3 B/px: 48 contiguous bytes per thread; 4 B/px: four 16-byte groups strided by the block (both conflict-free)
    const uint32_t my_smem = smem_base + (BPP == 3 ? tid * 48u : tid * 16u);
    const uint32_t my_step = (BPP == 3) ? 16u : nthr * 16u;

    auto flush = [&]() {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            atomicAdd(acc_sum + (2 * j) * nthr, accD[j] & 0xFFFFu);
            atomicAdd(acc_sum + (2 * j + 1) * nthr, accD[j] >> 16);
            atomicAdd(acc_cnt + (2 * j) * nthr, accM[j] & 0xFFFFu);
            atomicAdd(acc_cnt + (2 * j + 1) * nthr, accM[j] >> 16);
            accD[j] = accM[j] = 0u;
        }
    };

    uint32_t stage = 0, parity = 0, prev_stage = 0, prev_parity = 0, since_flush = 0;
#pragma unroll 1
    for (uint32_t i = 0; i < count; ++i) {
        // ---- consume stage: raw bytes -> registers, release the buffer
        mbar_wait(full_bar(stage), parity);
        uint32_t w[kWords];
#pragma unroll
        for (int v = 0; v < BPP; ++v) {
            const uint4 x = lds128(my_smem + stage * P.stage_bytes + my_step * v);
            w[4 * v] = x.x; w[4 * v + 1] = x.y; w[4 * v + 2] = x.z; w[4 * v + 3] = x.w;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar(stage));
        // ---- producer duty (thread 0): refill the buffer released during the previous iteration
        if (tid == 0) {
            const uint32_t j = i + S - 1;
            if (j < count) {
                if (i > 0) mbar_wait(empty_bar(prev_stage), prev_parity);
                issue(j, i > 0 ? prev_stage : S - 1);
            }
        }
        prev_stage = stage; prev_parity = parity;
        if (++stage == S) { stage = 0; parity ^= 1u; }

        // ---- intensity, difference, threshold, accumulate
        uint32_t cur[8];
        intensity16<BPP, CH>(w, cur);
        if (prime_first && i == 0) {  // halo frame t0-1: only establishes the previous-frame plane
#pragma unroll
            for (int j = 0; j < 8; ++j) ref[j] = cur[j];
            part += P.words_per_frame;
            continue;
        }
        uint32_t sD = 0u, sM = 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t d = __vmaxu2(cur[j], ref[j]) - __vminu2(cur[j], ref[j]);  // |cur-ref| per half, no borrow
            const uint32_t m = __viaddmin_s16x2_relu(d, negtau2, one2);              // (d - tau > 0) ? 1 : 0 per half
            accD[j] += d; accM[j] += m;
            sD += d; sM += m;
            if (MODE == 1) ref[j] = cur[j];
        }
        // per-frame scalars: fold halves (IDP.2A), pack sad | cnt<<20, warp-reduce, one store per warp
        const uint32_t packed = __dp2a_lo(sD, 0x0101u, __dp2a_lo(sM, 0x0101u, 0u) << 20);
        const uint32_t wsum = __reduce_add_sync(0xFFFFFFFFu, packed);
        if (lane == 0) *part = wsum;
        part += P.words_per_frame;
        if (++since_flush == (uint32_t)kFlushFrames) { flush(); since_flush = 0; }
    }
    flush();

    if (MODE == 1 && seg == P.n_segments - 1) {  // chain: I2 of the last frame becomes the next call's previous frame
        if constexpr (BPP == 3) {
            uint4* sp = reinterpret_cast<uint4*>(P.state_out + tile_first_px + (uint64_t)tid * kPxPerThread);
            sp[0] = make_uint4(ref[0], ref[1], ref[2], ref[3]);
            sp[1] = make_uint4(ref[4], ref[5], ref[6], ref[7]);
        } else {
#pragma unroll
            for (int v = 0; v < 4; ++v)
                *reinterpret_cast<uint2*>(P.state_out + tile_first_px + 4u * (v * nthr + tid)) =
                    make_uint2(ref[2 * v], ref[2 * v + 1]);
        }
    }
}

template <int BPP, int CH, int MODE>
cudaError_t launch_t(const Geometry& g, const ClipArgs& a, const KParams& kp, size_t smem, cudaStream_t s) {
    auto kfn = clip_kernel<BPP, CH, MODE>;
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid(g.n_tiles, a.n_segments, 1), block(g.threads, 1, 1);
    kfn<<<grid, block, smem, s>>>(kp);
    count_launch();
    return cudaGetLastError();
}

template <int BPP, int CH>
cudaError_t launch_m(const Geometry& g, const ClipArgs& a, const KParams& kp, size_t smem, cudaStream_t s) {
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

inline uint32_t stage_bytes_of(const Geometry& g) { return (g.tile_px * (uint32_t)g.bpp + 127u) & ~127u; }

}  // namespace

size_t clip_smem_bytes(const Geometry& g, uint32_t stages) {
    return (size_t)stages * stage_bytes_of(g) + 16u * stages;  // buffers + full/empty mbarriers
}

int clip_occupancy(const Geometry& g, uint32_t threads, uint32_t stages) {
    Geometry t = g;
    t.threads = threads;
    t.tile_px = threads * kPxPerThread;
    const size_t smem = clip_smem_bytes(t, stages) + 1024;  // + per-block reservation
    const size_t smem_sm = 227 * 1024;
    if (smem > smem_sm) return 0;
    int by_smem = (int)(smem_sm / smem);
    int by_regs = (int)(65536 / (64 * threads));
    int by_thr = (int)(2048 / threads);
    int occ = by_smem < by_regs ? by_smem : by_regs;
    occ = occ < by_thr ? occ : by_thr;
    return occ > 32 ? 32 : occ;
}

cudaError_t launch_clip(const Geometry& g, const ClipArgs& a, cudaStream_t s) {
    KParams kp;
    kp.frames = a.frames; kp.stride = a.stride; kp.npx = g.npx;
    kp.state_in = a.state_in; kp.state_out = a.state_out;
    kp.acc_sum = a.acc_sum; kp.acc_cnt = a.acc_cnt; kp.partials = a.partials;
    kp.n_frames = a.n_frames; kp.n_segments = a.n_segments;
    kp.tile_px = g.tile_px; kp.stages = g.stages; kp.stage_bytes = stage_bytes_of(g);
    kp.words_per_frame = g.n_tiles * (g.threads / 32);
    kp.tau = a.tau;
    const size_t smem = clip_smem_bytes(g, g.stages);
    return g.bpp == 3 ? launch_c<3>(g, a, kp, smem, s) : launch_c<4>(g, a, kp, smem, s);
}

}  // namespace dipsb
