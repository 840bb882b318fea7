// intensity.cuh -- packed-u16 intensity of raw pixels, shared by the device code of libdips_b200 (clip kernel, per-frame
// kernels, reference-plane builders).  get_intensity of the reference (dips/src/gpu/shaders/dips_shader.wgsl:64-82) as the
// exact integer I2 = max(r,g,b) + min(r,g,b) in [0,510] (2 * channel with a chroma filter), 16 pixels at a time in 8 packed
// u16x2 registers, with the packed DPX instructions (VIMNMX3.U16x2) and byte permutes (PRMT).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dipsb {

// ---- de-interleave + intensity ---------------------------------------------------------------------------------------
static constexpr uint32_t kLoMask = 0x00FF00FFu;

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

// 4 B/px: one 128-bit piece = 4 pixels -> two packed registers (pixels 0,1 and 2,3)
template <int CH>
__device__ __forceinline__ void intensity4_x(const uint4 x, uint32_t& i01, uint32_t& i23) {
    const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const uint32_t a = w[2 * j], b = w[2 * j + 1];
        uint32_t r;
        if constexpr (CH < 0) {
            const uint32_t t = __byte_perm(a, b, 0x5410);  // a0 a1 b0 b1
            const uint32_t p = __byte_perm(t, 0u, 0x4240);
            const uint32_t q = __byte_perm(t, 0u, 0x4341);
            const uint32_t z = __byte_perm(a, b, 0x0602) & kLoMask;  // (a2, b2)
            r = __vimax3_u16x2(p, q, z) + __vimin3_u16x2(p, q, z);
        } else {
            const uint32_t p = __byte_perm(a, b, sel_cross(CH, CH)) & kLoMask;
            r = p + p;
        }
        if (j == 0) i01 = r; else i23 = r;
    }
}

// 3 B/px: 4 pixels = 12 bytes = three words -> two packed registers (pixels 0,1 and 2,3)
template <int CH>
__device__ __forceinline__ void intensity4_3(const uint32_t a, const uint32_t b, const uint32_t c, uint32_t& i01, uint32_t& i23) {
    const uint32_t w[3] = {a, b, c};
    if constexpr (CH < 0) {
        const uint32_t x01 = __byte_perm(a, 0u, 0x4340);  // (a0, a3)
        const uint32_t t = __byte_perm(a, b, 0x5421);     // a1 a2 b0 b1
        const uint32_t y01 = __byte_perm(t, 0u, 0x4240);  // (a1, b0)
        const uint32_t z01 = __byte_perm(t, 0u, 0x4341);  // (a2, b1)
        const uint32_t u = __byte_perm(b, c, 0x6532);     // b2 b3 c1 c2
        const uint32_t x23 = __byte_perm(u, 0u, 0x4240);  // (b2, c1)
        const uint32_t y23 = __byte_perm(u, 0u, 0x4341);  // (b3, c2)
        const uint32_t z23 = __byte_perm(c, 0u, 0x4340);  // (c0, c3)
        i01 = __vimax3_u16x2(x01, y01, z01) + __vimin3_u16x2(x01, y01, z01);
        i23 = __vimax3_u16x2(x23, y23, z23) + __vimin3_u16x2(x23, y23, z23);
    } else {
        const uint32_t p01 = pair_bytes<CH, CH + 3>(w);
        const uint32_t p23 = pair_bytes<CH + 6, CH + 9>(w);
        i01 = p01 + p01;
        i23 = p23 + p23;
    }
}

}  // namespace dipsb
