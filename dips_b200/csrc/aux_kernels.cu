// aux_kernels.cu -- everything around the clip kernel: reference-plane builders (K1), the per-frame streaming kernel with
// the reference's visual colour mapping (K4), scalar finalisation, accumulator re-ordering and the synthetic clip generator.
// All of these are once-per-clip or PCIe-bound, so they are written for clarity; the roofline kernel is clip_kernel.cu.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <atomic>

#include "dipsb_internal.h"
#include "intensity.cuh"
#include "peer_sync.cuh"

namespace dipsb {

static std::atomic<uint64_t> g_launches{0};
void count_launch(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
uint64_t launch_count_value() { return g_launches.load(std::memory_order_relaxed); }

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ void chan_offsets(int format, int& r, int& g, int& b) {
    if (format == 2 || format == 3) { r = 2; g = 1; b = 0; }
    else { r = 0; g = 1; b = 2; }
}

// get_intensity (dips_shader.wgsl:64-82) as the integer 510*luminance; chan_byte >= 0 selects one byte (x2)
__device__ __forceinline__ uint32_t intensity2(const uint8_t* px, int chan_byte) {
    if (chan_byte >= 0) return 2u * px[chan_byte];
    const uint32_t a = px[0], b = px[1], c = px[2];
    return max(max(a, b), c) + min(min(a, b), c);
}

// K1: state[p] = I2(frame[p])          (pre_compute_main analogue with a single start frame)
__global__ void prime_kernel(const uint8_t* __restrict__ frame, uint64_t npx, int bpp, int chan_byte,
                             uint16_t* __restrict__ state) {
    for (uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < npx; p += (uint64_t)gridDim.x * blockDim.x)
        state[p] = (uint16_t)intensity2(frame + p * bpp, chan_byte);
}

// K1, vectorised: 16 pixels per thread -- 3 or 4 128-bit loads, packed intensity (base 16-byte aligned, npx a multiple of
// 16).  The 32 bytes a thread produces are exchanged through shared memory so that every store instruction of a warp covers
// 512 contiguous bytes: half-filled 32-byte sectors cost nothing in the local L2 but double the packets of a remote store.
// With a PlaneScatter (rank 0 of a sharded overall-mode pass) slice k of the plane is stored into rank k+1's state plane as
// well, and the last block stamps the peers: the scatter half of the reference-plane broadcast.
template <int BPP, int CH>
__global__ void __launch_bounds__(256) prime16_kernel(const uint8_t* __restrict__ frame, uint64_t npx, uint16_t* __restrict__ state,
                                                      const PlaneScatter sc) {
    __shared__ uint4 xbuf[8][64];
    const uint64_t groups = npx / 16;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    // whole warps iterate together (g0 = the warp's first group): the exchange below needs all 32 lanes
    for (uint64_t g0 = blockIdx.x * (uint64_t)blockDim.x + warp * 32u; g0 < groups; g0 += stride) {
        const uint64_t g = g0 + lane;
        uint32_t I[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (g < groups) {
            const uint4* src = reinterpret_cast<const uint4*>(frame + g * (16 * BPP));
            uint32_t w[BPP * 4];
#pragma unroll
            for (int v = 0; v < BPP; ++v) {
                const uint4 x = __ldg(src + v);
                w[4 * v] = x.x; w[4 * v + 1] = x.y; w[4 * v + 2] = x.z; w[4 * v + 3] = x.w;
            }
            intensity16<BPP, CH>(w, I);
        }
        xbuf[warp][2 * lane] = make_uint4(I[0], I[1], I[2], I[3]);
        xbuf[warp][2 * lane + 1] = make_uint4(I[4], I[5], I[6], I[7]);
        __syncwarp();
        const uint4 a = xbuf[warp][lane], b = xbuf[warp][32 + lane];
        __syncwarp();
        // 16-byte unit u of the warp's 1 KB holds pixels 8u .. 8u+7 of the 512 starting at 16*g0
        const uint64_t valid_units = min((uint64_t)64, 2 * (groups - g0));
        uint4* dst = reinterpret_cast<uint4*>(state + g0 * 16);
        if (lane < valid_units) dst[lane] = a;
        if (32 + lane < valid_units) dst[32 + lane] = b;
        if (sc.nranks) {   // slice k of the plane belongs to rank k + 1; a warp's 512 pixels may straddle two slices
            const uint64_t pa = g0 * 16 + 8ull * lane, pb = pa + 256;
            if (lane < valid_units) reinterpret_cast<uint4*>(sc.plane_peer[1u + (uint32_t)(pa / sc.slice_px)] + g0 * 16)[lane] = a;
            if (32 + lane < valid_units) reinterpret_cast<uint4*>(sc.plane_peer[1u + (uint32_t)(pb / sc.slice_px)] + g0 * 16)[32 + lane] = b;
        }
    }
    if (sc.nranks) stamp_when_last(sc.blocks_done, sc.stamp_peer, sc.nranks, sc.epoch);
}

// K1': upper median of 4 start frames, pre_compute_shader.wgsl:103-131 (element [2] of the ascending sort)
__global__ void prime_median4_kernel(const uint8_t* __restrict__ frames, uint64_t stride, uint64_t npx, int bpp,
                                     int chan_byte, uint16_t* __restrict__ state) {
    for (uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < npx; p += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = intensity2(frames + k * stride + p * bpp, chan_byte);
        // third smallest of four = min of the two pair-maxima and ... written as a 5-comparator network
        uint32_t lo01 = min(v[0], v[1]), hi01 = max(v[0], v[1]), lo23 = min(v[2], v[3]), hi23 = max(v[2], v[3]);
        uint32_t mid_hi = max(lo01, lo23);          // second or third smallest candidates
        uint32_t mid_lo = min(hi01, hi23);
        state[p] = (uint16_t)max(mid_hi, mid_lo);   // sorted[2]
    }
}

// per-frame scalars: sum the per-warp words (sad | cnt<<20) of one frame; one 128-thread block per frame, 128-bit loads
// (rows are pitched to a multiple of 4 words so that every row starts 16-byte aligned)
__global__ void finalize_scalars_kernel(const uint32_t* __restrict__ partials, uint32_t words_per_frame, uint32_t row_pitch,
                                        uint64_t* __restrict__ sad, uint64_t* __restrict__ cnt) {
    const uint32_t t = blockIdx.x;
    const uint32_t* row = partials + (uint64_t)t * row_pitch;
    unsigned long long s = 0, c = 0;
    const uint32_t nvec = words_per_frame / 4;
    const uint4* row4 = reinterpret_cast<const uint4*>(row);
    for (uint32_t i = threadIdx.x; i < nvec; i += blockDim.x) {
        const uint4 w = __ldg(row4 + i);
        s += (w.x & 0xFFFFFu) + (w.y & 0xFFFFFu) + (unsigned long long)(w.z & 0xFFFFFu) + (w.w & 0xFFFFFu);
        c += (w.x >> 20) + (w.y >> 20) + (w.z >> 20) + (w.w >> 20);
    }
    for (uint32_t i = nvec * 4 + threadIdx.x; i < words_per_frame; i += blockDim.x) {
        const uint32_t w = __ldg(row + i);
        s += w & 0xFFFFFu;
        c += w >> 20;
    }
    __shared__ unsigned long long sh_s[4], sh_c[4];
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_down_sync(0xFFFFFFFFu, s, o);
        c += __shfl_down_sync(0xFFFFFFFFu, c, o);
    }
    if ((threadIdx.x & 31) == 0) { sh_s[threadIdx.x >> 5] = s; sh_c[threadIdx.x >> 5] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 4; ++k) { s += sh_s[k]; c += sh_c[k]; }
        sad[t] = s;
        cnt[t] = c;
    }
}

__global__ void unpermute_kernel(const uint32_t* __restrict__ internal, uint32_t* __restrict__ planar, uint64_t npx,
                                 uint32_t tile_px, uint32_t threads, int bpp, int groups) {
    for (uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < npx; p += (uint64_t)gridDim.x * blockDim.x)
        planar[p] = internal[tile_order_index(p, tile_px, threads, bpp, groups)];
}
__global__ void permute_kernel(const uint32_t* __restrict__ planar, uint32_t* __restrict__ internal, uint64_t npx,
                               uint32_t tile_px, uint32_t threads, int bpp, int groups) {
    for (uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < npx; p += (uint64_t)gridDim.x * blockDim.x)
        internal[tile_order_index(p, tile_px, threads, bpp, groups)] = planar[p];
}
// Accumulator exchange format for the cross-GPU sum: fewer bytes over NVLink than the two u32 planes.
//   layout 1: one u32 per element, sum | count << sum_bits (valid while both totals fit their fields: no carry can cross);
//   layout 2: sum plane as is + counts as u16 pairs packed two per u32 (valid while the total count < 65536).
// Element-wise int32 addition of packed buffers == addition of the fields.
__global__ void pack_acc_kernel(const uint32_t* __restrict__ acc, uint64_t n_elems, uint32_t* __restrict__ out, int layout,
                                int sum_bits) {
    const uint32_t* sum = acc;
    const uint32_t* cnt = acc + n_elems;
    if (layout == 1) {
        for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_elems; i += (uint64_t)gridDim.x * blockDim.x)
            out[i] = sum[i] | (cnt[i] << sum_bits);
    } else {
        const uint64_t half = n_elems / 2;   // n_elems is a multiple of 512
        for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_elems + half; i += (uint64_t)gridDim.x * blockDim.x)
            out[i] = i < n_elems ? sum[i] : (cnt[2 * (i - n_elems)] | (cnt[2 * (i - n_elems) + 1] << 16));
    }
}
__global__ void unpack_acc_kernel(const uint32_t* __restrict__ in, uint64_t n_elems, uint32_t* __restrict__ acc, int layout,
                                  int sum_bits) {
    uint32_t* sum = acc;
    uint32_t* cnt = acc + n_elems;
    if (layout == 1) {
        const uint32_t mask = (sum_bits >= 32) ? 0xFFFFFFFFu : ((1u << sum_bits) - 1u);
        for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_elems; i += (uint64_t)gridDim.x * blockDim.x) {
            const uint32_t w = in[i];
            sum[i] = w & mask;
            cnt[i] = w >> sum_bits;
        }
    } else {
        const uint64_t half = n_elems / 2;
        for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_elems + half; i += (uint64_t)gridDim.x * blockDim.x) {
            const uint32_t w = in[i];
            if (i < n_elems) sum[i] = w;
            else { cnt[2 * (i - n_elems)] = w & 0xFFFFu; cnt[2 * (i - n_elems) + 1] = w >> 16; }
        }
    }
}

__global__ void intensity_map_kernel(const uint32_t* __restrict__ internal, float* __restrict__ out, uint64_t npx,
                                     uint32_t tile_px, uint32_t threads, int bpp, int groups, double inv_den) {
    for (uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < npx; p += (uint64_t)gridDim.x * blockDim.x)
        out[p] = (float)((double)internal[tile_order_index(p, tile_px, threads, bpp, groups)] * inv_den);
}

// ---- visual chain: compute_main colour mapping, dips_shader.wgsl:30-62, :97-118, :213-239 -----------------------------
__device__ __forceinline__ uint8_t unorm8(float v) {
    if (!(v > 0.0f)) return 0;
    if (v >= 1.0f) return 255;
    return (uint8_t)floorf(v * 255.0f + 0.5f);
}
__device__ void hsl_to_rgb(float h, float s, float l, float& r, float& g, float& b) {
    const float chroma = s * (1.0f - fabsf(2.0f * l - 1.0f));
    const float hp = h / 60.0f;
    const float x = chroma * (1.0f - fabsf(fmodf(hp, 2.0f) - 1.0f));
    const float m = l - chroma / 2.0f;
    r = g = b = 0.0f;
    if (hp >= 0 && hp < 1) { r = chroma; g = x; }
    else if (hp >= 1 && hp < 2) { r = x; g = chroma; }
    else if (hp >= 2 && hp < 3) { g = chroma; b = x; }
    else if (hp >= 3 && hp < 4) { g = x; b = chroma; }
    else if (hp >= 4 && hp < 5) { r = x; b = chroma; }
    else if (hp >= 5 && hp <= 6) { r = chroma; b = x; }
    r += m; g += m; b += m;
}
__device__ __forceinline__ uchar4 visual_pixel(int s_i2, int colorize, int filter, float sig) {
    float diff = (float)s_i2 / 510.0f;                                  // start.r - median, :213-214
    diff = diff * 0.5f;                                                 // map(-1,1 -> -0.5,0.5), :217
    if (filter == 0) diff = 1.0f / (1.0f + expf(-sig * diff)) - 0.5f;   // sigmoid, :108-112
    else if (filter == 1) diff = (-logf((1.0f / (diff + 0.5f)) - 1.0f)) / sig;  // inv_sigmoid, :114-118
    diff *= 5.0f;                                                       // SENSITIVITY, :229
    float r, g, b;
    if (colorize) {
        if (diff < 0) hsl_to_rgb(0.0f, fabsf(diff), 0.5f, r, g, b);
        else hsl_to_rgb(120.0f, diff, 0.5f, r, g, b);
    } else {
        r = g = b = 0.5f - diff;
    }
    return make_uchar4(unorm8(r), unorm8(g), unorm8(b), 255);
}

// Streaming kernel: one frame against the state plane.  Used by dipsb_push_frame (PCIe-bound path) and as the fallback
// of run_clip for frames whose base/stride are not 16-byte aligned (the TMA bulk copy needs that).
struct FrameK {
    const uint8_t* frame; uint64_t pitch; uint32_t width, height; int bpp, chan_byte;
    const uint16_t* i2src;       // spatially filtered intensity plane of this frame (window > 1), else nullptr
    const uint16_t* state_in; uint16_t* state_out; uint32_t* acc_sum; uint32_t* acc_cnt;
    unsigned long long* sad; unsigned long long* cnt; uint8_t* out_rgba;
    uint32_t tau, tile_px, threads; int geo_bpp, geo_groups; int accumulate, colorize, filter; float sig;
    uint64_t p_begin, p_end;     // pixel range of this launch (a row band of the frame, or all of it)
};
__global__ void frame_kernel(const FrameK K) {
    unsigned long long s = 0, c = 0;
    for (uint64_t p = K.p_begin + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < K.p_end; p += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t y = (uint32_t)(p / K.width), x = (uint32_t)(p - (uint64_t)y * K.width);
        const uint32_t cur = K.i2src ? K.i2src[p] : intensity2(K.frame + (uint64_t)y * K.pitch + (uint64_t)x * K.bpp, K.chan_byte);
        const uint32_t ref = K.state_in[p];
        const int sdiff = (int)ref - (int)cur;                          // sign convention start - current
        const uint32_t d = (uint32_t)(sdiff < 0 ? -sdiff : sdiff);
        const uint32_t m = d > K.tau ? 1u : 0u;
        if (K.accumulate) {
            const uint64_t q = tile_order_index(p, K.tile_px, K.threads, K.geo_bpp, K.geo_groups);
            K.acc_sum[q] += d;                                          // each pixel is owned by exactly one thread
            K.acc_cnt[q] += m;
            s += d; c += m;
        }
        if (K.state_out) K.state_out[p] = (uint16_t)cur;
        if (K.out_rgba) reinterpret_cast<uchar4*>(K.out_rgba)[p] = visual_pixel(sdiff, K.colorize, K.filter, K.sig);
    }
    if (K.accumulate) {
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_down_sync(0xFFFFFFFFu, s, o);
            c += __shfl_down_sync(0xFFFFFFFFu, c, o);
        }
        __shared__ unsigned long long sh_s[kThreads / 32], sh_c[kThreads / 32];
        if ((threadIdx.x & 31) == 0) { sh_s[threadIdx.x >> 5] = s; sh_c[threadIdx.x >> 5] = c; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int k = 1; k < kThreads / 32; ++k) { s += sh_s[k]; c += sh_c[k]; }
            atomicAdd(K.sad, s);
            atomicAdd(K.cnt, c);
        }
    }
}

// Reference-flavour rings (N1).  Everything is kept in I2 units (x510): a raw slot holds max+min, a grey slot holds
// 2*q with q = unorm8 store of the intensity = floor(I2/2 + 0.5) = (I2+1)>>1 (dips_shader.wgsl:187 writes the filtered
// newest frame back as rgba8unorm grey), so the float expressions of the shader become exact integers until the colour map.
struct RingK {
    const uint8_t* frame; uint64_t pitch; uint32_t width, height; int bpp, chan_byte;
    const uint16_t* i2src;
    uint16_t* ring; int n_slots, write_slot, grey_slot, compute_start, snapshot, median_is_max, do_diff;
    uint16_t* start; uint32_t* acc_sum; uint32_t* acc_cnt; unsigned long long* sad; unsigned long long* cnt; uint8_t* out_rgba;
    uint32_t tau, tile_px, threads; int geo_bpp, geo_groups, colorize, filter; float sig;
    uint64_t p_begin, p_end;     // pixel range of this launch
};
__device__ __forceinline__ uint32_t upper_median4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {   // sorted[2]
    const uint32_t lo01 = min(a, b), hi01 = max(a, b), lo23 = min(c, d), hi23 = max(c, d);
    return max(max(lo01, lo23), min(hi01, hi23));
}
__global__ void ring_kernel(const RingK K) {
    const uint64_t npx = (uint64_t)K.width * K.height;
    unsigned long long s = 0, c = 0;
    for (uint64_t p = K.p_begin + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < K.p_end; p += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t y = (uint32_t)(p / K.width), x = (uint32_t)(p - (uint64_t)y * K.width);
        const uint32_t raw = K.i2src ? K.i2src[p] : intensity2(K.frame + (uint64_t)y * K.pitch + (uint64_t)x * K.bpp, K.chan_byte);
        uint32_t v[4];                                              // compile-time indices only: stays in registers
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = k == K.write_slot ? raw : (k < K.n_slots ? K.ring[(uint64_t)k * npx + p] : 0u);
        uint32_t start = K.start[p], med;
        bool grey_out = false;
        if (K.n_slots == 4) {                                       // `dips`
            if (K.compute_start) {                                   // pre_compute_shader.wgsl:103-131, stored rgba8unorm
                start = 2u * ((upper_median4(v[0], v[1], v[2], v[3]) + 1u) >> 1);
                K.start[p] = (uint16_t)start;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t grey = 2u * ((v[k] + 1u) >> 1);          // dips_shader.wgsl:187
                v[k] = k == K.grey_slot ? grey : v[k];
                if (k == K.grey_slot || k == K.write_slot) K.ring[(uint64_t)k * npx + p] = (uint16_t)v[k];
            }
            med = upper_median4(v[0], v[1], v[2], v[3]);             // :191-214
        } else {                                                     // `dips_alt`, NUM_TEXTURES = 2
            K.ring[(uint64_t)K.write_slot * npx + p] = (uint16_t)raw;
            med = K.median_is_max ? max(v[0], v[1]) : min(v[0], v[1]);
            if (K.snapshot) {                                        // pre_compute_shader.wgsl:231-235
                start = 2u * ((med + 1u) >> 1);
                K.start[p] = (uint16_t)start;
                grey_out = true;
            }
        }
        if (!K.do_diff) continue;
        if (grey_out) {
            const uint8_t q = (uint8_t)(start >> 1);
            if (K.out_rgba) reinterpret_cast<uchar4*>(K.out_rgba)[p] = make_uchar4(q, q, q, 255);
            continue;
        }
        const int sdiff = (int)start - (int)med;
        const uint32_t d = (uint32_t)(sdiff < 0 ? -sdiff : sdiff);
        const uint32_t m = d > K.tau ? 1u : 0u;
        const uint64_t q = tile_order_index(p, K.tile_px, K.threads, K.geo_bpp, K.geo_groups);
        K.acc_sum[q] += d;
        K.acc_cnt[q] += m;
        s += d; c += m;
        if (K.out_rgba) reinterpret_cast<uchar4*>(K.out_rgba)[p] = visual_pixel(sdiff, K.colorize, K.filter, K.sig);
    }
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_down_sync(0xFFFFFFFFu, s, o);
        c += __shfl_down_sync(0xFFFFFFFFu, c, o);
    }
    __shared__ unsigned long long sh_s[kThreads / 32], sh_c[kThreads / 32];
    if ((threadIdx.x & 31) == 0) { sh_s[threadIdx.x >> 5] = s; sh_c[threadIdx.x >> 5] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < kThreads / 32; ++k) { s += sh_s[k]; c += sh_c[k]; }
        if (s) atomicAdd(K.sad, s);
        if (c) atomicAdd(K.cnt, c);
    }
}

// ---- 4 pixels per thread ("quad") versions of the per-frame kernels -----------------------------------------------------
// One thread = 4 consecutive pixels of a row: one 128-bit load of RGBx (three 32-bit loads of RGB), packed intensity, one
// 64-bit load / store per u16 plane, one 128-bit store of the RGBA output.  Row and column come from one 32-bit division per
// quad; the accumulator slots of a quad are idx0 + e*threads (e = 0..3) in the clip kernel's tile order, so a warp touches
// whole 32-byte sectors of the planes.  Taken when the width is a multiple of 4 and base / pitch are aligned; every other
// frame goes to the one-pixel-per-thread kernels above.
__device__ __forceinline__ uint32_t rgba_word(uchar4 v) { return (uint32_t)v.x | ((uint32_t)v.y << 8) | ((uint32_t)v.z << 16) | ((uint32_t)v.w << 24); }

template <int FBPP, int CH>
__device__ __forceinline__ void quad_intensity(const uint8_t* __restrict__ row, uint32_t x, uint32_t& i01, uint32_t& i23) {
    if constexpr (FBPP == 4) {
        intensity4_x<CH>(__ldg(reinterpret_cast<const uint4*>(row + 4ull * x)), i01, i23);
    } else {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(row + 3ull * x);
        intensity4_3<CH>(__ldg(w), __ldg(w + 1), __ldg(w + 2), i01, i23);
    }
}
__device__ __forceinline__ void block_add_scalars(unsigned long long s, unsigned long long c, unsigned long long* sad, unsigned long long* cnt) {
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_down_sync(0xFFFFFFFFu, s, o);
        c += __shfl_down_sync(0xFFFFFFFFu, c, o);
    }
    __shared__ unsigned long long sh_s[kThreads / 32], sh_c[kThreads / 32];
    if ((threadIdx.x & 31) == 0) { sh_s[threadIdx.x >> 5] = s; sh_c[threadIdx.x >> 5] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < kThreads / 32; ++k) { s += sh_s[k]; c += sh_c[k]; }
        if (s) atomicAdd(sad, s);
        if (c) atomicAdd(cnt, c);
    }
}

template <int FBPP, int CH>
__global__ void __launch_bounds__(kThreads) frame4_kernel(const FrameK K) {
    unsigned long long s = 0, c = 0;
    const uint32_t q_begin = (uint32_t)(K.p_begin / 4), q_end = (uint32_t)(K.p_end / 4);
    for (uint32_t q = q_begin + blockIdx.x * blockDim.x + threadIdx.x; q < q_end; q += gridDim.x * blockDim.x) {
        const uint32_t p = 4u * q;
        uint32_t cur01, cur23;
        if (K.i2src) {
            const uint2 v = __ldg(reinterpret_cast<const uint2*>(K.i2src + p));
            cur01 = v.x; cur23 = v.y;
        } else {
            const uint32_t y = p / K.width, x = p - y * K.width;
            quad_intensity<FBPP, CH>(K.frame + (uint64_t)y * K.pitch, x, cur01, cur23);
        }
        const uint2 ref = __ldg(reinterpret_cast<const uint2*>(K.state_in + p));
        const int cur[4] = {(int)(cur01 & 0xFFFFu), (int)(cur01 >> 16), (int)(cur23 & 0xFFFFu), (int)(cur23 >> 16)};
        const int rf[4] = {(int)(ref.x & 0xFFFFu), (int)(ref.x >> 16), (int)(ref.y & 0xFFFFu), (int)(ref.y >> 16)};
        int sd[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) sd[e] = rf[e] - cur[e];                 // sign convention start - current
        if (K.accumulate) {
            const uint64_t i0 = tile_order_index(p, K.tile_px, K.threads, K.geo_bpp, K.geo_groups);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint32_t d = (uint32_t)(sd[e] < 0 ? -sd[e] : sd[e]), m = d > K.tau ? 1u : 0u;
                K.acc_sum[i0 + (uint64_t)e * K.threads] += d;               // each slot is owned by exactly one thread
                K.acc_cnt[i0 + (uint64_t)e * K.threads] += m;
                s += d; c += m;
            }
        }
        if (K.state_out) *reinterpret_cast<uint2*>(K.state_out + p) = make_uint2(cur01, cur23);
        if (K.out_rgba)
            *reinterpret_cast<uint4*>(K.out_rgba + 4ull * p) =
                make_uint4(rgba_word(visual_pixel(sd[0], K.colorize, K.filter, K.sig)), rgba_word(visual_pixel(sd[1], K.colorize, K.filter, K.sig)),
                           rgba_word(visual_pixel(sd[2], K.colorize, K.filter, K.sig)), rgba_word(visual_pixel(sd[3], K.colorize, K.filter, K.sig)));
    }
    if (K.accumulate) block_add_scalars(s, c, K.sad, K.cnt);
}

__device__ __forceinline__ uint32_t quad_elem(uint2 v, int e) { return e == 0 ? (v.x & 0xFFFFu) : e == 1 ? (v.x >> 16) : e == 2 ? (v.y & 0xFFFFu) : (v.y >> 16); }
__device__ __forceinline__ uint32_t grey_i2(uint32_t i2) { return 2u * ((i2 + 1u) >> 1); }   // rgba8unorm store of an intensity, in I2 units

template <int FBPP, int CH>
__global__ void __launch_bounds__(kThreads) ring4_kernel(const RingK K) {
    const uint64_t npx = (uint64_t)K.width * K.height;
    unsigned long long s = 0, c = 0;
    const uint32_t q_begin = (uint32_t)(K.p_begin / 4), q_end = (uint32_t)(K.p_end / 4);
    for (uint32_t q = q_begin + blockIdx.x * blockDim.x + threadIdx.x; q < q_end; q += gridDim.x * blockDim.x) {
        const uint32_t p = 4u * q;
        uint2 raw;
        if (K.i2src) raw = __ldg(reinterpret_cast<const uint2*>(K.i2src + p));
        else {
            const uint32_t y = p / K.width, x = p - y * K.width;
            quad_intensity<FBPP, CH>(K.frame + (uint64_t)y * K.pitch, x, raw.x, raw.y);
        }
        // the ring slots of my 4 pixels (compile-time slot indices only: everything stays in registers)
        const uint2 s0 = 0 == K.write_slot ? raw : *reinterpret_cast<const uint2*>(K.ring + p);
        const uint2 s1 = 1 == K.write_slot ? raw : *reinterpret_cast<const uint2*>(K.ring + npx + p);
        const uint2 s2 = K.n_slots < 4 ? make_uint2(0, 0) : (2 == K.write_slot ? raw : *reinterpret_cast<const uint2*>(K.ring + 2 * npx + p));
        const uint2 s3 = K.n_slots < 4 ? make_uint2(0, 0) : (3 == K.write_slot ? raw : *reinterpret_cast<const uint2*>(K.ring + 3 * npx + p));
        const uint2 st = *reinterpret_cast<const uint2*>(K.start + p);
        uint32_t start[4], med[4], g0[4], g1[4], g2[4], g3[4];       // g*: the slots after the in-place grey quantisation
        bool grey_out = false;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const uint32_t a = quad_elem(s0, e), b = quad_elem(s1, e), cc = quad_elem(s2, e), d = quad_elem(s3, e);
            start[e] = quad_elem(st, e);
            if (K.n_slots == 4) {                                    // `dips`
                if (K.compute_start) start[e] = grey_i2(upper_median4(a, b, cc, d));      // pre_compute_shader.wgsl:103-131
                g0[e] = K.grey_slot == 0 ? grey_i2(a) : a; g1[e] = K.grey_slot == 1 ? grey_i2(b) : b;             // dips_shader.wgsl:187
                g2[e] = K.grey_slot == 2 ? grey_i2(cc) : cc; g3[e] = K.grey_slot == 3 ? grey_i2(d) : d;
                med[e] = upper_median4(g0[e], g1[e], g2[e], g3[e]);                       // :191-214
            } else {                                                 // `dips_alt`, NUM_TEXTURES = 2
                g0[e] = a; g1[e] = b; g2[e] = cc; g3[e] = d;
                med[e] = K.median_is_max ? max(a, b) : min(a, b);
                if (K.snapshot) start[e] = grey_i2(med[e]);          // pre_compute_shader.wgsl:231-235
            }
        }
        if (K.n_slots == 4) {
            if (K.compute_start) *reinterpret_cast<uint2*>(K.start + p) = make_uint2(start[0] | (start[1] << 16), start[2] | (start[3] << 16));
            if (K.grey_slot == 0 || K.write_slot == 0) *reinterpret_cast<uint2*>(K.ring + p) = make_uint2(g0[0] | (g0[1] << 16), g0[2] | (g0[3] << 16));
            if (K.grey_slot == 1 || K.write_slot == 1) *reinterpret_cast<uint2*>(K.ring + npx + p) = make_uint2(g1[0] | (g1[1] << 16), g1[2] | (g1[3] << 16));
            if (K.grey_slot == 2 || K.write_slot == 2) *reinterpret_cast<uint2*>(K.ring + 2 * npx + p) = make_uint2(g2[0] | (g2[1] << 16), g2[2] | (g2[3] << 16));
            if (K.grey_slot == 3 || K.write_slot == 3) *reinterpret_cast<uint2*>(K.ring + 3 * npx + p) = make_uint2(g3[0] | (g3[1] << 16), g3[2] | (g3[3] << 16));
        } else {
            *reinterpret_cast<uint2*>(K.ring + (uint64_t)K.write_slot * npx + p) = raw;
            if (K.snapshot) {
                *reinterpret_cast<uint2*>(K.start + p) = make_uint2(start[0] | (start[1] << 16), start[2] | (start[3] << 16));
                grey_out = true;
            }
        }
        if (!K.do_diff) continue;
        if (grey_out) {
            if (K.out_rgba)
                *reinterpret_cast<uint4*>(K.out_rgba + 4ull * p) =
                    make_uint4((start[0] >> 1) * 0x010101u | 0xFF000000u, (start[1] >> 1) * 0x010101u | 0xFF000000u,
                               (start[2] >> 1) * 0x010101u | 0xFF000000u, (start[3] >> 1) * 0x010101u | 0xFF000000u);
            continue;
        }
        const uint64_t i0 = tile_order_index(p, K.tile_px, K.threads, K.geo_bpp, K.geo_groups);
        uint32_t o[4] = {0, 0, 0, 0};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int sdiff = (int)start[e] - (int)med[e];
            const uint32_t d = (uint32_t)(sdiff < 0 ? -sdiff : sdiff), m = d > K.tau ? 1u : 0u;
            K.acc_sum[i0 + (uint64_t)e * K.threads] += d;
            K.acc_cnt[i0 + (uint64_t)e * K.threads] += m;
            s += d; c += m;
            if (K.out_rgba) o[e] = rgba_word(visual_pixel(sdiff, K.colorize, K.filter, K.sig));
        }
        if (K.out_rgba) *reinterpret_cast<uint4*>(K.out_rgba + 4ull * p) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    block_add_scalars(s, c, K.sad, K.cnt);
}

template <int FBPP>
__global__ void __launch_bounds__(kThreads) passthrough4_kernel(const uint8_t* __restrict__ frame, uint64_t pitch, uint32_t width, uint32_t q_begin,
                                                                uint32_t q_end, int bgr, uint8_t* __restrict__ out) {
    for (uint32_t q = q_begin + blockIdx.x * blockDim.x + threadIdx.x; q < q_end; q += gridDim.x * blockDim.x) {
        const uint32_t p = 4u * q, y = p / width, x = p - y * width;
        const uint8_t* row = frame + (uint64_t)y * pitch;
        uint32_t px[4];                                               // r | g<<8 | b<<16 | a<<24 per pixel
        if constexpr (FBPP == 4) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(row + 4ull * x));
            px[0] = v.x; px[1] = v.y; px[2] = v.z; px[3] = v.w;
        } else {
            const uint32_t* w = reinterpret_cast<const uint32_t*>(row + 3ull * x);
            const uint32_t a = __ldg(w), b = __ldg(w + 1), c = __ldg(w + 2);
            px[0] = __byte_perm(a, 0xFF000000u, 0x7210);
            px[1] = __byte_perm(a, b, 0x0543) | 0xFF000000u;
            px[2] = __byte_perm(b, c, 0x0432) | 0xFF000000u;
            px[3] = __byte_perm(c, 0xFF000000u, 0x7321);
        }
        if (bgr) {
#pragma unroll
            for (int e = 0; e < 4; ++e) px[e] = __byte_perm(px[e], 0u, 0x3012);      // swap bytes 0 and 2
        }
        *reinterpret_cast<uint4*>(out + 4ull * p) = make_uint4(px[0], px[1], px[2], px[3]);
    }
}

// N4: spatial median of the intensity plane, window w in {3,5,7}.  A CORRECT median: the full symmetric window
// [-w/2, +w/2]^2, zero for taps outside the frame (as the reference pads, dips_shader.wgsl:135-139), element w*w/2 of the
// ascending order.  The reference's own loop covers only the half-open window [-w/2, w/2) and picks the wrong element
//  -- a defect that is deliberately not reproduced.  Selection by bitwise binary search on the 9-bit value.
template <int w>
__global__ void spatial_median_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, uint32_t W, uint32_t H) {
    const uint64_t npx = (uint64_t)W * H;
    constexpr int r = w / 2, k = (w * w) / 2, n = w * w;        // compile-time window: the taps live in registers
    for (uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < npx; p += (uint64_t)gridDim.x * blockDim.x) {
        const int y = (int)(p / W), x = (int)(p - (uint64_t)y * W);
        uint32_t v[n];
#pragma unroll
        for (int dy = -r; dy <= r; ++dy)
#pragma unroll
            for (int dx = -r; dx <= r; ++dx) {
                const int yy = y + dy, xx = x + dx;
                v[(dy + r) * w + dx + r] = (yy >= 0 && yy < (int)H && xx >= 0 && xx < (int)W) ? in[(uint64_t)yy * W + xx] : 0u;
            }
        uint32_t lo = 0;                                   // largest value with #(v < value) <= k  ==  sorted[k]
        for (int bit = 8; bit >= 0; --bit) {
            const uint32_t cand = lo | (1u << bit);
            int cnt = 0;
#pragma unroll
            for (int i = 0; i < n; ++i) cnt += v[i] < cand;
            if (cnt <= k) lo = cand;
        }
        out[p] = (uint16_t)lo;
    }
}

// The same median, tiled: a block of 32 x 8 threads filters a 64 x 8 pixel tile out of shared memory, every thread two
// horizontally adjacent pixels packed in one u16x2 register.  The w*w packed taps stay in registers; each of the 9 steps of
// the bitwise binary search counts "tap >= candidate" for both pixels, and the count is split over the two instruction
// pipes that issue side by side (measured, tools/native/pipe_rate.cu: VIADDMNMX 2 clk per warp and sub-partition, HFMA2 2 clk,
// the two interleaved 1 clk per instruction; HSET2 and LOP3 queue behind VIADDMNMX on the ALU pipe, IMAD half does):
//   * 2/3 of the taps: VIADDMNMX.S16x2.RELU gives 0/1 per half, summed two at a time by IADD3 (ALU pipe as well);
//   * 1/3 of the taps are kept as the half-precision numbers 1024 + tap (bit pattern tap + 0x6400, exact): HFMA2.SAT
//     computes clamp(tap - candidate + 1, 0, 1) exactly, HADD2 sums it onto 1024.0, whose bit pattern is 0x6400 + count;
//   * half of the warps run the half-precision part first, the other half the integer part (see the loop).
// ncu (profiles/r02_median7_ncu.txt): 1084 instructions per warp, ALU pipe 76 % busy, issue slots 73 % -- the kernel is bound
// by the ALU pipe (VIADDMNMX and IADD3 both live there); other splits of the taps (1/2, 3/7) measure the same +-3 %.
// Even width only (a pair never straddles the right edge); odd widths take spatial_median_kernel.
__device__ __forceinline__ uint32_t hfma2_sat(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("fma.rn.sat.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t hadd2(uint32_t a, uint32_t b) {
    uint32_t r;
    asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
template <int w>
__global__ void __launch_bounds__(256) spatial_median_tile_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, uint32_t W, uint32_t H) {
    constexpr int r = w / 2, n = w * w, k = n / 2;
    constexpr int nB = n - n / 3;                                 // taps counted on the ALU pipe; the rest in half precision on the FMA pipe
    constexpr uint32_t kHalf1024 = 0x64006400u;                   // 1024.0 in both halves
    constexpr int kRows = 8 + 2 * r, kWords = 36;                 // tile columns x0-4 .. x0+67 as 36 words of two pixels
    __shared__ uint32_t tile[kRows][kWords];
    const uint32_t x0 = blockIdx.x * 64u, y0 = blockIdx.y * 8u;
    for (uint32_t i = threadIdx.x; i < (uint32_t)(kRows * kWords); i += 256u) {
        const uint32_t ry = i / kWords, j = i - ry * kWords;
        const int y = (int)(y0 + ry) - r, x = (int)(x0 + 2u * j) - 4;      // x even: the pair is inside the frame or outside, never split
        tile[ry][j] = (y >= 0 && y < (int)H && x >= 0 && x < (int)W) ? __ldg(reinterpret_cast<const uint32_t*>(in + (uint64_t)y * W + x)) : 0u;
    }
    __syncthreads();
    const uint32_t t = threadIdx.x & 31u, ty = threadIdx.x >> 5;
    const uint32_t x = x0 + 2u * t, y = y0 + ty;
    uint32_t v[n];
#pragma unroll
    for (int dy = 0; dy < w; ++dy) {
        uint32_t q[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) q[i] = tile[ty + dy][t + i];
#pragma unroll
        for (int dx = -r; dx <= r; ++dx) {
            // pixels (x+dx, x+1+dx) sit at tile columns 2t+4+dx, 2t+5+dx
            const int c = 4 + dx, i = dy * w + dx + r;
            const uint32_t tap = (c & 1) ? __funnelshift_r(q[(c - 1) / 2], q[(c + 1) / 2], 16) : q[c / 2];
            v[i] = i < nB ? tap : tap + kHalf1024;
        }
    }
    uint32_t lo = 0u;                                             // per half: largest value with #(tap < value) <= k == sorted[k]
#pragma unroll 1
    for (int bit = 8; bit >= 0; --bit) {
        const uint32_t cand = lo + (0x00010001u << bit);
        const uint32_t one_minus_cand = __vsub2(0x00010001u, cand);
        const uint32_t neg_cand_h = (cand + 0x63FF63FFu) | 0x80008000u;      // -(1023 + candidate) per half
        // ptxas emits the two kinds as two runs, and the warps of a block advance in step: with every warp in the same run
        // one pipe would idle at a time.  The two warps a block has on each sub-partition (warp w and w + 4) therefore take
        // the runs in opposite order; each run starts from the other's count, which pins the order (1024 + count as a
        // half-precision number is the bit pattern 0x6400 + count).
        uint32_t ge;
        auto count_int = [&](uint32_t from) {
#pragma unroll
            for (int i = 0; i < nB; ++i) from += __viaddmin_s16x2_relu(v[i], one_minus_cand, 0x00010001u);
            return from;
        };
        auto count_half = [&](uint32_t from) {
            uint32_t acc = from + kHalf1024;
#pragma unroll
            for (int i = nB; i < n; ++i) acc = hadd2(acc, hfma2_sat(v[i], 0x3C003C00u, neg_cand_h));
            return acc - kHalf1024;
        };
        if (ty & 4u) ge = count_int(count_half(0u));
        else ge = count_half(count_int(0u));
        // #(tap < cand) <= k  <=>  ge >= n - k
        const uint32_t keep = __viaddmin_s16x2_relu(ge, (uint32_t)(((1 - (n - k)) & 0xFFFF) * 0x00010001u), 0x00010001u);
        lo += keep << bit;
    }
    if (x < W && y < H) *reinterpret_cast<uint32_t*>(out + (uint64_t)y * W + x) = lo;
}

__global__ void median4_planes_kernel(const uint16_t* __restrict__ planes, uint64_t npx, uint16_t* __restrict__ out) {
    for (uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < npx; p += (uint64_t)gridDim.x * blockDim.x)
        out[p] = (uint16_t)upper_median4(planes[p], planes[npx + p], planes[2 * npx + p], planes[3 * npx + p]);
}

// Re-pack of an unaligned device clip (odd base, pitch or frame size) into rows of a 16-byte pitch whose padding is zero:
// one thread per 16-byte output chunk, source read as aligned 32-bit words and realigned with funnel shifts.
__global__ void repack_kernel(const uint8_t* __restrict__ src, uint64_t stride, uint64_t fb, uint64_t n_frames,
                              uint8_t* __restrict__ dst, uint64_t dpitch) {
    const uint64_t chunks = dpitch / 16, total = chunks * n_frames;
    const uint8_t* end = src + (n_frames - 1) * stride + fb;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t k = i / chunks, o = (i - k * chunks) * 16;
        const uintptr_t a = (uintptr_t)(src + k * stride + o);
        const uint32_t* a4 = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
        const uint32_t sh = (uint32_t)(a & 3u) * 8u;
        uint32_t w[5];
#pragma unroll
        for (int t = 0; t < 5; ++t) w[t] = (reinterpret_cast<const uint8_t*>(a4 + t) < end) ? __ldg(a4 + t) : 0u;
        uint32_t v[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) v[t] = __funnelshift_r(w[t], w[t + 1], sh);
        const uint64_t remain = fb > o ? fb - o : 0;             // bytes of this chunk that belong to the frame
        if (remain < 16) {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const uint64_t vb = remain > 4u * t ? remain - 4u * t : 0;
                v[t] &= vb >= 4 ? 0xFFFFFFFFu : (vb == 0 ? 0u : (1u << (8u * (uint32_t)vb)) - 1u);
            }
        }
        *reinterpret_cast<uint4*>(dst + k * dpitch + o) = make_uint4(v[0], v[1], v[2], v[3]);
    }
}

// warm-up passthrough of frame_callback (dips/src/lib.rs:241-245): input converted to RGBA8, alpha 255
__global__ void passthrough_kernel(const uint8_t* __restrict__ frame, uint64_t pitch, uint32_t width, uint64_t p_begin,
                                   uint64_t p_end, int format, uint8_t* __restrict__ out) {
    const int bpp = (format == 0 || format == 2) ? 3 : 4;
    int ro, go, bo;
    chan_offsets(format, ro, go, bo);
    for (uint64_t p = p_begin + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < p_end; p += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t y = (uint32_t)(p / width), x = (uint32_t)(p - (uint64_t)y * width);
        const uint8_t* px = frame + (uint64_t)y * pitch + (uint64_t)x * bpp;
        reinterpret_cast<uchar4*>(out)[p] = make_uchar4(px[ro], px[go], px[bo], bpp == 4 ? px[3] : 255);
    }
}

// ---- synthetic clips (same bytes as oracle/dips_oracle.c: dipso_synth_fill) ------------------------------------------
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}
__device__ __forceinline__ uint32_t hash_byte(uint64_t seed, uint64_t index) {
    const uint64_t v = mix64(seed + ((index >> 3) + 1) * 0x9E3779B97F4A7C15ull);
    return (uint32_t)(v >> (8 * (index & 7))) & 0xFFu;
}
__global__ void synth_kernel(uint8_t* __restrict__ dst, uint64_t first_frame, uint64_t n_frames, uint32_t W, uint32_t H,
                             int bpp, uint64_t seed, int profile) {
    const uint64_t fb = (uint64_t)W * H * bpp;
    const uint64_t total = fb * n_frames;
    for (uint64_t o = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; o < total; o += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t k = o / fb, i = o - k * fb, t = first_frame + k;
        const uint64_t g = t * fb + i;
        uint32_t v;
        if (profile == 0) {
            v = hash_byte(seed, g);
        } else {
            const int bg = (int)hash_byte(seed ^ 0xB5AD4ECEDA1CE2A9ull, i);
            const int noise = (int)(hash_byte(seed, g) % 17u) - 8;
            const uint64_t p = i / (uint64_t)bpp;
            const uint32_t x = (uint32_t)(p % W), y = (uint32_t)(p / W);
            const uint32_t bx = (uint32_t)((t * 7u) % W), by = (uint32_t)((t * 3u) % H);
            const uint32_t bw = W * 5u / 16u, bh = H * 5u / 16u;
            const uint32_t dx = (x + W - bx) % W, dy = (y + H - by) % H;
            int q = bg + noise + ((dx < bw && dy < bh) ? 64 : 0);
            q = q < 0 ? 0 : (q > 255 ? 255 : q);
            v = (uint32_t)q;
        }
        dst[o] = (uint8_t)v;
    }
}

inline int grid_for(uint64_t n, const Geometry& g) {
    uint64_t b = (n + kThreads - 1) / kThreads;
    const uint64_t cap = (uint64_t)(g.num_sms ? g.num_sms : 148) * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

bool prime_fast_path(const Geometry& g, const uint8_t* frame) { return ((uintptr_t)frame & 15u) == 0 && (g.npx & 15u) == 0; }

template <int BPP>
static void launch_prime16(const Geometry& g, const uint8_t* frame, uint16_t* state, cudaStream_t s, const PlaneScatter& sc) {
    const uint64_t groups = g.npx / 16;
    const uint64_t cap = (uint64_t)(g.num_sms ? g.num_sms : 148) * 8;
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((groups + 255) / 256, cap));
    switch (g.chan_byte) {
        case 0: prime16_kernel<BPP, 0><<<grid, 256, 0, s>>>(frame, g.npx, state, sc); break;
        case 1: prime16_kernel<BPP, 1><<<grid, 256, 0, s>>>(frame, g.npx, state, sc); break;
        case 2: prime16_kernel<BPP, 2><<<grid, 256, 0, s>>>(frame, g.npx, state, sc); break;
        default: prime16_kernel<BPP, -1><<<grid, 256, 0, s>>>(frame, g.npx, state, sc); break;
    }
}

// scatter != nullptr requires prime_fast_path (the caller falls back to prime + a separate push otherwise)
cudaError_t launch_prime(const Geometry& g, const uint8_t* frame, uint16_t* state, cudaStream_t s, const PlaneScatter* scatter) {
    if (prime_fast_path(g, frame)) {
        const PlaneScatter none;
        if (g.bpp == 3) launch_prime16<3>(g, frame, state, s, scatter ? *scatter : none);
        else launch_prime16<4>(g, frame, state, s, scatter ? *scatter : none);
    } else {
        if (scatter) return cudaErrorInvalidValue;
        prime_kernel<<<grid_for(g.npx, g), kThreads, 0, s>>>(frame, g.npx, g.bpp, g.chan_byte, state);
    }
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_prime_median4(const Geometry& g, const uint8_t* frames, uint64_t stride, uint16_t* state,
                                 cudaStream_t s) {
    prime_median4_kernel<<<grid_for(g.npx, g), kThreads, 0, s>>>(frames, stride, g.npx, g.bpp, g.chan_byte, state);
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_finalize_scalars(const Geometry&, const uint32_t* partials, uint32_t n_frames,
                                    uint32_t words_per_frame, uint64_t* sad, uint64_t* cnt, cudaStream_t s) {
    if (n_frames == 0) return cudaSuccess;
    finalize_scalars_kernel<<<n_frames, 128, 0, s>>>(partials, words_per_frame, (words_per_frame + 3u) & ~3u, sad, cnt);
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_unpermute(const Geometry& g, const uint32_t* internal, uint32_t* planar, cudaStream_t s) {
    unpermute_kernel<<<grid_for(g.npx, g), kThreads, 0, s>>>(internal, planar, g.npx, g.tile_px, g.threads, g.bpp, g.groups);
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_permute(const Geometry& g, const uint32_t* planar, uint32_t* internal, cudaStream_t s) {
    permute_kernel<<<grid_for(g.npx, g), kThreads, 0, s>>>(planar, internal, g.npx, g.tile_px, g.threads, g.bpp, g.groups);
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_pack_acc(const Geometry& g, const uint32_t* acc, uint32_t* out, int layout, int sum_bits, cudaStream_t s) {
    pack_acc_kernel<<<grid_for(g.n_elems, g), kThreads, 0, s>>>(acc, g.n_elems, out, layout, sum_bits);
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_unpack_acc(const Geometry& g, const uint32_t* in, uint32_t* acc, int layout, int sum_bits, cudaStream_t s) {
    unpack_acc_kernel<<<grid_for(g.n_elems, g), kThreads, 0, s>>>(in, g.n_elems, acc, layout, sum_bits);
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_intensity_map(const Geometry& g, const uint32_t* internal, uint64_t n_eff, float* out, cudaStream_t s) {
    const double inv = 1.0 / (510.0 * (double)(n_eff ? n_eff : 1));
    intensity_map_kernel<<<grid_for(g.npx, g), kThreads, 0, s>>>(internal, out, g.npx, g.tile_px, g.threads, g.bpp, g.groups, inv);
    count_launch();
    return cudaGetLastError();
}
// the 4-pixels-per-thread kernels need whole quads inside a row, aligned loads and 32-bit pixel indices
static bool quad_path(const Geometry& g, const uint8_t* frame, uint64_t pitch, int frame_bpp, uint64_t p_begin, uint64_t p_end) {
    const uint64_t align = frame_bpp == 4 ? 15u : 3u;
    return (g.width & 3u) == 0 && g.npx < (1ull << 32) && ((p_begin | p_end) & 3u) == 0 &&
           (frame == nullptr || (((uintptr_t)frame | pitch) & align) == 0);
}
#define DIPSB_DISPATCH_QUAD(LAUNCH, fbpp, ch)                                         \
    do {                                                                              \
        if ((fbpp) == 4) {                                                            \
            switch (ch) { case 0: LAUNCH(4, 0); break; case 1: LAUNCH(4, 1); break; case 2: LAUNCH(4, 2); break; default: LAUNCH(4, -1); } \
        } else {                                                                      \
            switch (ch) { case 0: LAUNCH(3, 0); break; case 1: LAUNCH(3, 1); break; case 2: LAUNCH(3, 2); break; default: LAUNCH(3, -1); } \
        }                                                                             \
    } while (0)

cudaError_t launch_frame(const Geometry& g, const FrameArgs& a, cudaStream_t s) {
    FrameK K;
    K.frame = a.frame; K.pitch = a.pitch; K.width = g.width; K.height = g.height;
    K.bpp = (a.format == 0 || a.format == 2) ? 3 : 4; K.chan_byte = a.chan_byte; K.i2src = a.i2src;
    K.state_in = a.state_in; K.state_out = a.state_out; K.acc_sum = a.acc_sum; K.acc_cnt = a.acc_cnt;
    K.sad = reinterpret_cast<unsigned long long*>(a.sad); K.cnt = reinterpret_cast<unsigned long long*>(a.cnt);
    K.out_rgba = a.out_rgba; K.tau = a.tau; K.tile_px = g.tile_px; K.threads = g.threads; K.geo_bpp = g.bpp; K.geo_groups = g.groups;
    K.accumulate = a.accumulate; K.colorize = a.colorize; K.filter = a.filter; K.sig = a.sig_scalar;
    K.p_begin = a.p_begin; K.p_end = a.p_end ? a.p_end : g.npx;
    if (K.p_end <= K.p_begin) return cudaSuccess;
    if (quad_path(g, a.frame, a.pitch, K.bpp, K.p_begin, K.p_end)) {
        const int grid = grid_for((K.p_end - K.p_begin) / 4, g);
#define DIPSB_FRAME4(FB, C) frame4_kernel<FB, C><<<grid, kThreads, 0, s>>>(K)
        DIPSB_DISPATCH_QUAD(DIPSB_FRAME4, K.bpp, K.chan_byte);
#undef DIPSB_FRAME4
    } else {
        frame_kernel<<<grid_for(K.p_end - K.p_begin, g), kThreads, 0, s>>>(K);
    }
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_spatial_median(const Geometry& g, const uint16_t* in, uint16_t* out, int window, cudaStream_t s) {
    if ((g.width & 1u) == 0 && (((uintptr_t)in | (uintptr_t)out) & 3u) == 0) {   // tiled, two pixels per thread
        const dim3 grid((g.width + 63) / 64, (g.height + 7) / 8, 1);
        switch (window) {
            case 3: spatial_median_tile_kernel<3><<<grid, 256, 0, s>>>(in, out, g.width, g.height); break;
            case 5: spatial_median_tile_kernel<5><<<grid, 256, 0, s>>>(in, out, g.width, g.height); break;
            case 7: spatial_median_tile_kernel<7><<<grid, 256, 0, s>>>(in, out, g.width, g.height); break;
            default: return cudaErrorInvalidValue;
        }
        count_launch();
        return cudaGetLastError();
    }
    switch (window) {
        case 3: spatial_median_kernel<3><<<grid_for(g.npx, g), kThreads, 0, s>>>(in, out, g.width, g.height); break;
        case 5: spatial_median_kernel<5><<<grid_for(g.npx, g), kThreads, 0, s>>>(in, out, g.width, g.height); break;
        case 7: spatial_median_kernel<7><<<grid_for(g.npx, g), kThreads, 0, s>>>(in, out, g.width, g.height); break;
        default: return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_median4_planes(const Geometry& g, const uint16_t* planes, uint16_t* out, cudaStream_t s) {
    median4_planes_kernel<<<grid_for(g.npx, g), kThreads, 0, s>>>(planes, g.npx, out);
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_ring(const Geometry& g, const RingArgs& a, cudaStream_t s) {
    RingK K;
    K.frame = a.frame; K.pitch = a.pitch; K.width = g.width; K.height = g.height;
    K.bpp = (a.format == 0 || a.format == 2) ? 3 : 4; K.chan_byte = a.chan_byte; K.i2src = a.i2src;
    K.ring = a.ring; K.n_slots = a.n_slots; K.write_slot = a.write_slot; K.grey_slot = a.grey_slot;
    K.compute_start = a.compute_start; K.snapshot = a.snapshot; K.median_is_max = a.median_is_max; K.do_diff = a.do_diff;
    K.start = a.start; K.acc_sum = a.acc_sum; K.acc_cnt = a.acc_cnt;
    K.sad = reinterpret_cast<unsigned long long*>(a.sad); K.cnt = reinterpret_cast<unsigned long long*>(a.cnt);
    K.out_rgba = a.out_rgba; K.tau = a.tau; K.tile_px = g.tile_px; K.threads = g.threads; K.geo_bpp = g.bpp; K.geo_groups = g.groups;
    K.colorize = a.colorize; K.filter = a.filter; K.sig = a.sig_scalar;
    K.p_begin = a.p_begin; K.p_end = a.p_end ? a.p_end : g.npx;
    if (K.p_end <= K.p_begin) return cudaSuccess;
    if (quad_path(g, a.frame, a.pitch, K.bpp, K.p_begin, K.p_end)) {
        const int grid = grid_for((K.p_end - K.p_begin) / 4, g);
#define DIPSB_RING4(FB, C) ring4_kernel<FB, C><<<grid, kThreads, 0, s>>>(K)
        DIPSB_DISPATCH_QUAD(DIPSB_RING4, K.bpp, K.chan_byte);
#undef DIPSB_RING4
    } else {
        ring_kernel<<<grid_for(K.p_end - K.p_begin, g), kThreads, 0, s>>>(K);
    }
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_passthrough_rgba(const Geometry& g, const uint8_t* frame, uint64_t pitch, int format, uint8_t* out,
                                    cudaStream_t s, uint64_t p_begin, uint64_t p_end) {
    if (!p_end) p_end = g.npx;
    if (p_end <= p_begin) return cudaSuccess;
    const int fbpp = (format == 0 || format == 2) ? 3 : 4;
    if (quad_path(g, frame, pitch, fbpp, p_begin, p_end) && ((uintptr_t)out & 15u) == 0) {
        const int grid = grid_for((p_end - p_begin) / 4, g), bgr = (format == 2 || format == 3) ? 1 : 0;
        if (fbpp == 4) passthrough4_kernel<4><<<grid, kThreads, 0, s>>>(frame, pitch, g.width, (uint32_t)(p_begin / 4), (uint32_t)(p_end / 4), bgr, out);
        else passthrough4_kernel<3><<<grid, kThreads, 0, s>>>(frame, pitch, g.width, (uint32_t)(p_begin / 4), (uint32_t)(p_end / 4), bgr, out);
        count_launch();
        return cudaGetLastError();
    }
    passthrough_kernel<<<grid_for(p_end - p_begin, g), kThreads, 0, s>>>(frame, pitch, g.width, p_begin, p_end, format, out);
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_repack(const Geometry& g, const uint8_t* src, uint64_t stride, uint64_t fb, uint64_t n_frames, uint8_t* dst,
                          uint64_t dpitch, cudaStream_t s) {
    if (!n_frames) return cudaSuccess;
    uint64_t b = (dpitch / 16 * n_frames + kThreads - 1) / kThreads;
    const uint64_t cap = (uint64_t)(g.num_sms ? g.num_sms : 148) * 32;
    repack_kernel<<<(int)(b > cap ? cap : b), kThreads, 0, s>>>(src, stride, fb, n_frames, dst, dpitch);
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_synth(uint8_t* dst, uint64_t first_frame, uint64_t n_frames, uint32_t w, uint32_t h, int bpp,
                         uint64_t seed, int profile, cudaStream_t s) {
    const uint64_t total = (uint64_t)w * h * bpp * n_frames;
    if (total == 0) return cudaSuccess;
    uint64_t b = (total + kThreads - 1) / kThreads;
    if (b > 148ull * 64) b = 148ull * 64;
    synth_kernel<<<(int)b, kThreads, 0, s>>>(dst, first_frame, n_frames, w, h, bpp, seed, profile);
    count_launch();
    return cudaGetLastError();
}

}  // namespace dipsb
