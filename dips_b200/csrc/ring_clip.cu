// ring_clip.cu -- the reference-exact temporal rings over a whole device-resident clip in ONE launch.
//
// What the per-frame ring kernel (aux_kernels.cu, ring4_kernel) does frame by frame -- re-reading the ring planes, the start
// plane and the accumulators for every frame, ~45 MB of traffic for a 6 MB frame -- this kernel does with the ring in
// registers: a thread owns 8 consecutive pixels for a run of frames, keeps the 4 (dips) or 2 (dips_alt) ring slots, the start /
// snapshot plane, the packed u16 sums and the byte counts of its pixels in registers and reads every frame byte once.
//
//   `dips` steady state (frame 5 onwards; dips/src/gpu/bind_groups.rs:407-427, dips_shader.wgsl:187-214):
//       slot[idx] = grey(I2(frame)); idx = (idx + 1) % 4; D = | start - sorted4(slots)[2] |
//   `dips_alt` (dips_alt/src/dips_compute/shaders/pre_compute_shader.wgsl:212-262, NUM_TEXTURES = 2):
//       slot[idx] = I2(frame); idx ^= 1; D = | snapshot - min(slots) |   (max with the in-bounds median)
//
// The frames that do something else -- the three pass-through frames and the start-plane frame of the `dips` warm-up, a
// frame that takes a snapshot -- stay with the per-frame kernel (api.cu, run_clip_ring): at most four per call.
//
// Work split: grid.x = blocks of 256 threads over the pixels, grid.y = frame segments whose starts are multiples of the
// ring length.  In steady state the ring is a function of the last NS frames alone, so a segment that does not start the
// call rebuilds it from the NS frames before its first (re-read, < 4 % with the segment lengths the host picks); segment 0
// loads it from the context's ring planes and the last segment stores it back, so that per-frame calls and batch calls mix.
// Sums/counts: packed registers flushed to the u32 accumulator planes (internal tile order) with RED.ADD every 128 frames.
// Per-frame scalars: sad | cnt << 20 per warp per frame into the scratch rows that finalize_scalars_kernel sums.
#include <algorithm>
#include <type_traits>

#include "dipsb_internal.h"
#include "intensity.cuh"

namespace dipsb {

namespace {

constexpr int kRcThreads = 256;
constexpr int kRcPx = 8;          // pixels per thread: 4 packed u16x2 registers per plane
constexpr int kRcAhead = 2;       // frames between the register load and the L2 prefetch

struct RingClipK {
    const uint8_t* frames; uint64_t stride; uint32_t n_frames, seg_frames;
    uint32_t n_units;             // npx / 8
    uint64_t npx;
    uint16_t* ring; uint32_t rot; // ring planes; slot the first frame of the launch overwrites
    const uint16_t* start;
    uint32_t* acc_sum; uint32_t* acc_cnt;
    uint32_t* partials; uint32_t row_pitch;
    uint32_t tau, tile_px, threads; int geo_bpp, geo_groups;
};

template <int FBPP>
__device__ __forceinline__ void load_px8(const uint8_t* p, bool active, uint32_t (&w)[2 * FBPP]) {
    if (!active) {
#pragma unroll
        for (int i = 0; i < 2 * FBPP; ++i) w[i] = 0u;
        return;
    }
    if constexpr (FBPP == 4) {
        const uint4 a = __ldcs(reinterpret_cast<const uint4*>(p)), b = __ldcs(reinterpret_cast<const uint4*>(p) + 1);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    } else {
        const uint2 a = __ldcs(reinterpret_cast<const uint2*>(p)), b = __ldcs(reinterpret_cast<const uint2*>(p) + 1),
                    c = __ldcs(reinterpret_cast<const uint2*>(p) + 2);
        w[0] = a.x; w[1] = a.y; w[2] = b.x; w[3] = b.y; w[4] = c.x; w[5] = c.y;
    }
}

// pull the 128-byte lines of a later frame into L2 (one lane in four: consecutive lanes are 24 / 32 bytes apart): a thread
// has registers for one frame in flight only, and HBM latency is longer than one frame's worth of math
__device__ __forceinline__ void prefetch_l2(const uint8_t* p, uint32_t lane) {
    if ((lane & 3u) == 0u) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// what a frame leaves in its ring slot: I2 of the 8 pixels, quantised to grey for `dips` (the rgba8unorm store of
// dips_shader.wgsl:187: 2 * ((I2 + 1) >> 1) == (I2 + 1) & ~1; no carry between the halves, I2 <= 510)
template <int FBPP, int CH, int NS>
__device__ __forceinline__ void slot_value(const uint32_t (&w)[2 * FBPP], uint32_t (&v)[4]) {
    if constexpr (FBPP == 4) {
        intensity4_x<CH>(make_uint4(w[0], w[1], w[2], w[3]), v[0], v[1]);
        intensity4_x<CH>(make_uint4(w[4], w[5], w[6], w[7]), v[2], v[3]);
    } else {
        intensity4_3<CH>(w[0], w[1], w[2], v[0], v[1]);
        intensity4_3<CH>(w[3], w[4], w[5], v[2], v[3]);
    }
    if constexpr (NS == 4) {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = (v[i] + 0x00010001u) & 0xFFFEFFFEu;
    }
}

template <int FBPP, int CH, int NS, int MEDMAX>
__global__ void __launch_bounds__(kRcThreads, (NS == 4 || FBPP == 4) ? 4 : 5) ring_clip_kernel(const RingClipK K) {
    const uint32_t unit = blockIdx.x * (uint32_t)kRcThreads + threadIdx.x;
    const bool active = unit < K.n_units;
    const uint32_t lane = threadIdx.x & 31u, gwarp = unit >> 5;
    const uint32_t seg = blockIdx.y;
    const uint32_t f0 = seg * K.seg_frames, f1 = min(K.n_frames, f0 + K.seg_frames);
    const uint64_t p = (uint64_t)kRcPx * unit;
    const uint8_t* src = K.frames + p * FBPP;

    uint32_t r[NS][4], st[4], w[2 * FBPP];
    if (active) {
        const uint4 s4 = __ldg(reinterpret_cast<const uint4*>(K.start + p));
        st[0] = s4.x; st[1] = s4.y; st[2] = s4.z; st[3] = s4.w;
    } else {
        st[0] = st[1] = st[2] = st[3] = 0u;
    }
    if (seg == 0) {                                   // register i <-> ring slot (rot + i) % NS: the order the frames overwrite them
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (active) v = *reinterpret_cast<const uint4*>(K.ring + (uint64_t)((K.rot + i) % NS) * K.npx + p);
            r[i][0] = v.x; r[i][1] = v.y; r[i][2] = v.z; r[i][3] = v.w;
        }
    } else {                                          // steady state: the ring is what the NS frames before f0 left
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            load_px8<FBPP>(src + (uint64_t)(f0 - NS + i) * K.stride, active, w);
            slot_value<FBPP, CH, NS>(w, r[i]);
        }
    }
    uint32_t base[2] = {0u, 0u};                      // accumulator index of my two 4-pixel runs (element e at + e * threads)
    if (active) {
        base[0] = (uint32_t)tile_order_index(p, K.tile_px, K.threads, K.geo_bpp, K.geo_groups);
        base[1] = (uint32_t)tile_order_index(p + 4, K.tile_px, K.threads, K.geo_bpp, K.geo_groups);
    }
    uint32_t as[4] = {0u, 0u, 0u, 0u}, ac[2] = {0u, 0u};   // sums as u16 pairs; counts as bytes (pixels 4b .. 4b+3 in bytes 0, 2, 1, 3 of ac[b])
    const uint32_t tau = K.tau > 511u ? 511u : K.tau;
    const uint32_t negtau2 = ((0u - tau) & 0xFFFFu) * 0x00010001u;

    auto flush = [&]() {
        if (active) {
#pragma unroll
            for (int b = 0; b < 2; ++b) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint32_t d = (e & 1) ? as[2 * b + e / 2] >> 16 : as[2 * b + e / 2] & 0xFFFFu;
                    const int byte = (e == 1) ? 2 : (e == 2) ? 1 : e;
                    const uint32_t m = (ac[b] >> (8 * byte)) & 0xFFu;
                    const uint32_t idx = base[b] + (uint32_t)e * K.threads;
                    atomicAdd(K.acc_sum + idx, d);
                    atomicAdd(K.acc_cnt + idx, m);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) as[i] = 0u;
        ac[0] = ac[1] = 0u;
    };
    // frame j arrives in `w` and lands in register slot S
    auto step = [&](auto slot_tag, uint32_t j) {
        constexpr int S = decltype(slot_tag)::value;
        slot_value<FBPP, CH, NS>(w, r[S]);
        if (j + 1 < f1) load_px8<FBPP>(src + (uint64_t)(j + 1) * K.stride, active, w);   // next frame in flight under the math
        if (j + 1 + kRcAhead < f1 && active) prefetch_l2(src + (uint64_t)(j + 1 + kRcAhead) * K.stride, lane);
        uint32_t d[4], m[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t med;
            if constexpr (NS == 4) {                  // sorted[2] of the four (dips_shader.wgsl:191-214, in-bounds reading)
                const uint32_t lo01 = __vminu2(r[0][i], r[1][i]), hi01 = __vmaxu2(r[0][i], r[1][i]);
                const uint32_t lo23 = __vminu2(r[2][i], r[3][i]), hi23 = __vmaxu2(r[2][i], r[3][i]);
                med = __vimax3_u16x2(lo01, lo23, __vminu2(hi01, hi23));
            } else {
                med = MEDMAX ? __vmaxu2(r[0][i], r[1][i]) : __vminu2(r[0][i], r[1][i]);
            }
            d[i] = 2u * __vmaxu2(st[i], med) - (st[i] + med);            // |start - median| per half
            m[i] = __viaddmin_s16x2_relu(d[i], negtau2, 0x00010001u);   // D > tau
            as[i] += d[i];
        }
        ac[0] += m[0] + (m[1] << 8);
        ac[1] += m[2] + (m[3] << 8);
        const uint32_t sD = d[0] + d[1] + d[2] + d[3], sM = m[0] + m[1] + m[2] + m[3];
        const uint32_t word = __dp2a_lo(sD, 0x0101u, __dp2a_lo(sM, 0x0101u, 0u) << 20);
        const uint32_t tot = __reduce_add_sync(0xFFFFFFFFu, word);       // <= 32*8*510 < 2^20, <= 256 < 2^12
        if (lane == 0) K.partials[(uint64_t)j * K.row_pitch + gwarp] = tot;
    };

    if (f0 < f1) load_px8<FBPP>(src + (uint64_t)f0 * K.stride, active, w);
#pragma unroll
    for (int a = 1; a <= kRcAhead; ++a)
        if (f0 + a < f1 && active) prefetch_l2(src + (uint64_t)(f0 + a) * K.stride, lane);
    uint32_t j = f0, since = 0;
    for (; j + NS <= f1; j += NS) {
        step(std::integral_constant<int, 0>{}, j);
        step(std::integral_constant<int, 1>{}, j + 1);
        if constexpr (NS == 4) {
            step(std::integral_constant<int, 2>{}, j + 2);
            step(std::integral_constant<int, 3>{}, j + 3);
        }
        since += NS;
        if (since >= (uint32_t)kFlushFrames) { flush(); since = 0; }
    }
    // ragged tail (only the last segment has one: the others are whole multiples of NS)
    if (j < f1) step(std::integral_constant<int, 0>{}, j);
    if constexpr (NS == 4) {
        if (j + 1 < f1) step(std::integral_constant<int, 1>{}, j + 1);
        if (j + 2 < f1) step(std::integral_constant<int, 2>{}, j + 2);
    }
    flush();
    if (seg == gridDim.y - 1 && active) {
#pragma unroll
        for (int i = 0; i < NS; ++i)
            *reinterpret_cast<uint4*>(K.ring + (uint64_t)((K.rot + i) % NS) * K.npx + p) = make_uint4(r[i][0], r[i][1], r[i][2], r[i][3]);
    }
}

template <int FBPP, int CH>
cudaError_t launch_variant(const RingClipK& K, int n_slots, int median_is_max, dim3 grid, cudaStream_t s, int* occupancy) {
    auto go = [&](auto kernel) -> cudaError_t {
        if (occupancy) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occupancy, kernel, kRcThreads, 0);
        kernel<<<grid, kRcThreads, 0, s>>>(K);
        return cudaGetLastError();
    };
    if (n_slots == 4) return go(ring_clip_kernel<FBPP, CH, 4, 0>);
    return median_is_max ? go(ring_clip_kernel<FBPP, CH, 2, 1>) : go(ring_clip_kernel<FBPP, CH, 2, 0>);
}
cudaError_t dispatch(const Geometry& g, const RingClipK& K, int n_slots, int median_is_max, dim3 grid, cudaStream_t s, int* occupancy) {
#define DIPSB_RC(FB, C) return launch_variant<FB, C>(K, n_slots, median_is_max, grid, s, occupancy)
    if (g.bpp == 4) {
        switch (g.chan_byte) { case 0: DIPSB_RC(4, 0); case 1: DIPSB_RC(4, 1); case 2: DIPSB_RC(4, 2); default: DIPSB_RC(4, -1); }
    } else {
        switch (g.chan_byte) { case 0: DIPSB_RC(3, 0); case 1: DIPSB_RC(3, 1); case 2: DIPSB_RC(3, 2); default: DIPSB_RC(3, -1); }
    }
#undef DIPSB_RC
}

}  // namespace

bool ring_clip_available(const Geometry& g, const uint8_t* frames, uint64_t stride) {
    const uint64_t align = g.bpp == 4 ? 15u : 7u;     // 128-bit / 64-bit loads of 8 pixels
    return (g.npx % kRcPx) == 0 && g.n_elems < (1ull << 32) && g.npx / kRcPx < (1ull << 31) && (((uintptr_t)frames | stride) & align) == 0;
}

uint32_t ring_clip_words_per_frame(const Geometry& g) {
    const uint32_t blocks = (uint32_t)((g.npx / kRcPx + kRcThreads - 1) / kRcThreads);
    return blocks * (kRcThreads / 32);
}

// segment length for n frames: the fewest segments that fill whole waves of resident blocks (the blocks of a partial last
// wave run the full frame loop next to an idle machine), each a multiple of the ring length and long enough that the
// re-read of n_slots frames per segment stays small
static uint32_t plan_seg_frames(const Geometry& g, uint32_t n, int n_slots, int occupancy) {
    const uint64_t blocks = (g.npx / kRcPx + kRcThreads - 1) / kRcThreads;
    const uint64_t cap = (uint64_t)std::max(1, occupancy) * (g.num_sms ? g.num_sms : 148);
    const uint32_t min_len = 32u * (uint32_t)n_slots;
    const uint32_t max_segs = std::max<uint32_t>(1u, std::min<uint32_t>(n / min_len, 64u));
    uint32_t best = 1;
    double best_cost = 1e300;
    for (uint32_t segs = 1; segs <= max_segs; ++segs) {
        const uint32_t len = ((n + segs - 1) / segs + n_slots - 1) / n_slots * n_slots;
        const uint32_t real = (n + len - 1) / len;
        const uint64_t waves = (blocks * real + cap - 1) / cap;
        const double cost = (double)waves * (len + (real > 1 ? n_slots : 0));   // time ~ waves x frames each block walks
        if (cost < best_cost * 0.98) { best_cost = cost; best = segs; }
    }
    return ((n + best - 1) / best + n_slots - 1) / n_slots * n_slots;
}

cudaError_t launch_ring_clip(const Geometry& g, const RingClipArgs& a, cudaStream_t s) {
    if (a.n_frames == 0) return cudaSuccess;
    RingClipK K;
    K.frames = a.frames; K.stride = a.stride; K.n_frames = a.n_frames;
    K.n_units = (uint32_t)(g.npx / kRcPx); K.npx = g.npx;
    K.ring = a.ring; K.rot = (uint32_t)a.first_slot; K.start = a.start;
    K.acc_sum = a.acc_sum; K.acc_cnt = a.acc_cnt; K.partials = a.partials;
    K.row_pitch = (ring_clip_words_per_frame(g) + 3u) & ~3u;
    K.tau = a.tau; K.tile_px = g.tile_px; K.threads = g.threads; K.geo_bpp = g.bpp; K.geo_groups = g.groups;
    int occ = 0;
    cudaError_t e = dispatch(g, K, a.n_slots, a.median_is_max, dim3(), s, &occ);
    if (e != cudaSuccess) return e;
    K.seg_frames = a.seg_frames ? (a.seg_frames + a.n_slots - 1) / a.n_slots * a.n_slots : plan_seg_frames(g, a.n_frames, a.n_slots, occ);
    const dim3 grid((K.n_units + kRcThreads - 1) / kRcThreads, (a.n_frames + K.seg_frames - 1) / K.seg_frames, 1);
    e = dispatch(g, K, a.n_slots, a.median_is_max, grid, s, nullptr);
    count_launch();
    return e;
}

}  // namespace dipsb
