/*
 * dips_oracle.h -- CPU ORACLE for the DiPs per-pixel frame-difference hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (libdips_b200.so) never links, loads or calls anything in oracle/.
 *
 * PARITY UNPINNED: the reference (RubenMovsesyan/DiPs) has no CPU implementation, no
 * tests, no golden vectors and cannot be built or run here (no Rust toolchain, no
 * wgpu adapter).  This file is a plain-C restatement of the arithmetic in the
 * reference's own WGSL shaders; every function cites the reference file:line it
 * follows.  The known-answer vectors in tests/golden/ are hand-derived from those
 * formulas, not emitted by the reference.
 *
 * Integer contract:
 *   I2(p)      = max(r,g,b) + min(r,g,b)            in [0,510]   (= 510 x WGSL get_intensity)
 *              = 2 * channel                         when a chroma filter is set
 *   overall    D_t(p) = |I2_t(p) - I2_ref(p)|        ref = frame 0 (or upper-median of 4)
 *   per-frame  D_t(p) = |I2_t(p) - I2_{t-1}(p)|,     D_0 = 0
 *   mask       M_t(p) = D_t(p) > tau
 *   acc_sum[p] = sum_t D_t(p)  (u32)   acc_cnt[p] = sum_t M_t(p)  (u32)
 *   sad[t]     = sum_p D_t(p)  (u64)   cnt[t]     = sum_p M_t(p)  (u64)
 */
#ifndef DIPS_ORACLE_H
#define DIPS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* pixel formats: bytes per pixel and which byte is R/G/B */
enum { DIPSO_FMT_RGB8 = 0, DIPSO_FMT_RGBX8 = 1, DIPSO_FMT_BGR8 = 2, DIPSO_FMT_BGRX8 = 3 };
/* chroma filter, numbering of dips/src/lib.rs:52-61 and dips_shader.wgsl:64-71 */
enum { DIPSO_CHROMA_NONE = 0, DIPSO_CHROMA_RED = 1, DIPSO_CHROMA_GREEN = 2, DIPSO_CHROMA_BLUE = 3 };
enum { DIPSO_MODE_OVERALL = 0, DIPSO_MODE_PERFRAME = 1 };
/* filter, numbering of dips/src/lib.rs:32-41 (Unfiltered -> 255 -> WGSL "default:" branch) */
enum { DIPSO_FILTER_SIGMOID = 0, DIPSO_FILTER_INV_SIGMOID = 1, DIPSO_FILTER_NONE = 255 };
enum { DIPSO_SYNTH_UNIFORM = 0, DIPSO_SYNTH_SCENE = 1 };

int dipso_bytes_per_pixel(int fmt);

/* A1: get_intensity, dips/src/gpu/shaders/dips_shader.wgsl:64-82, as the exact integer 2*luminance*255 */
uint16_t dipso_intensity2(const uint8_t *px, int fmt, int chroma);

/* K1: I2 plane of one frame (npx pixels) */
void dipso_i2_plane(const uint8_t *frame, size_t npx, int fmt, int chroma, uint16_t *out);

/* A2: upper median (element [2] of the ascending sort of 4) of the I2 planes of 4 frames,
 * dips/src/gpu/shaders/pre_compute_shader.wgsl:103-131 */
void dipso_median4_plane(const uint8_t *const frames[4], size_t npx, int fmt, int chroma, uint16_t *out);

/*
 * X1-X5: run n_frames frames (frame k at frames + k*stride) through the difference path.
 *   state      in/out u16[npx]: overall mode -> the reference plane (read only);
 *              per-frame mode -> I2 of the previous frame on entry, of the last frame on exit.
 *   acc_sum/acc_cnt  u32[npx], accumulated into (caller zeroes them).
 *   sad/cnt    u64[n_frames], overwritten.
 *   nthreads   OpenMP threads (<=0: all cores).
 */
void dipso_run_clip(const uint8_t *frames, size_t n_frames, size_t stride, size_t npx, int fmt,
                    int chroma, int mode, uint32_t tau, uint16_t *state, uint32_t *acc_sum,
                    uint32_t *acc_cnt, uint64_t *sad, uint64_t *cnt, int nthreads);

/*
 * N4: spatial median filter of an I2 plane, window w in {1,3,5,7} -- spatial_median_filter, dips_shader.wgsl:122-170,
 * restated as the CORRECT median: full symmetric window [-w/2, +w/2]^2, zero for taps outside the frame (the reference
 * pads with 0.0, :135-139), element w*w/2 of the ascending order.  The reference's loop only visits the half-open window
 * [-w/2, w/2) and reads element w*w/2+1 of a mostly-zero array; that defect is deliberately not restated.
 */
void dipso_spatial_median_plane(const uint16_t *in, uint32_t width, uint32_t height, int window, uint16_t *out);

/* X6: float outputs derived from the integers */
void dipso_intensity_map(const uint32_t *acc_sum, size_t npx, uint64_t n_eff, float *out);
void dipso_frame_means(const uint64_t *sad, size_t n_frames, size_t npx, float *out);

/* A3/X7: the float "visual" chain of compute_main, dips_shader.wgsl:213-239, applied to the signed
 * difference S = I2_ref - I2_cur (in I2 units).  Writes one RGBA8 pixel. */
void dipso_visual_pixel(int32_t s_i2, int colorize, int filter, float sig_scalar, uint8_t out[4]);
void dipso_visual_frame(const uint16_t *ref, const uint16_t *cur, size_t npx, int colorize, int filter,
                        float sig_scalar, uint8_t *out_rgba);
/* float value of the chain just before colour mapping (for KATs) */
float dipso_visual_diff(int32_t s_i2, int filter, float sig_scalar);

/* deterministic synthetic clips: counter-based hash, identical on CPU and GPU */
uint64_t dipso_mix64(uint64_t z);
void dipso_synth_fill(uint8_t *dst, uint64_t first_frame, uint64_t n_frames, uint32_t width,
                      uint32_t height, int fmt, uint64_t seed, int profile, int nthreads);

/*
 * Reference-flavour state machine of the `dips` crate (N1): 4-frame ring + start plane.
 * Follows dips/src/gpu/mod.rs:170-216 (add_texture), :306-397 (dispatch), bind_groups.rs:18,
 * :407-427 (ring slot update), dips/src/lib.rs:233-246 (frame_callback passthrough).
 * Input/output frames are tightly packed RGBA8.  Returns 1 when the frame was passed through
 * unchanged (warm-up, fewer than 4 frames seen), 0 when out_rgba holds the visual frame.
 */
typedef struct dipso_cs dipso_cs;
dipso_cs *dipso_cs_new(uint32_t width, uint32_t height, int colorize, int filter, float sig_scalar,
                       int chroma);
void dipso_cs_free(dipso_cs *cs);
int dipso_cs_frame(dipso_cs *cs, const uint8_t *rgba_in, uint8_t *rgba_out);

/*
 * Reference-flavour state machine of `dips_alt` (N1): DiPsCompute::send_frame, dips_alt/src/dips_compute/mod.rs:498-646,
 * kernel dips_alt/src/dips_compute/shaders/pre_compute_shader.wgsl:188-263 with NUM_TEXTURES = FRAME_COUNT = 2
 * (dips_alt/src/lib.rs:36).  `snapshot` != 0 == send_frame(.., Some(()), ..).  `intended_median` == 0 reproduces the
 * as-shipped sort (a zero sentinel is sorted in, so median_array[N/2] is the MIN of the two frames);
 * 1 gives the in-bounds reading (sorted[N/2] of the N real values).  Textures start zero-initialised (wgpu).
 */
typedef struct dipso_alt dipso_alt;
dipso_alt *dipso_alt_new(uint32_t width, uint32_t height, int colorize, int filter, float sig_scalar, int chroma,
                         int intended_median);
void dipso_alt_free(dipso_alt *a);
void dipso_alt_frame(dipso_alt *a, const uint8_t *rgba_in, int snapshot, uint8_t *rgba_out);

int dipso_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
