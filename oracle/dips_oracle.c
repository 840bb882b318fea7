/*
 * dips_oracle.c -- CPU ORACLE (test infrastructure only; see dips_oracle.h header comment).
 * PARITY UNPINNED: restatement of the reference's WGSL, no reference-shipped vectors exist.
 * File:line citations are relative to the reference repository root.
 */
#include "dips_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int dipso_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int dipso_bytes_per_pixel(int fmt) { return (fmt == DIPSO_FMT_RGB8 || fmt == DIPSO_FMT_BGR8) ? 3 : 4; }

/* byte offsets of R, G, B inside one pixel */
static inline void channel_offsets(int fmt, int *r, int *g, int *b) {
    if (fmt == DIPSO_FMT_BGR8 || fmt == DIPSO_FMT_BGRX8) { *r = 2; *g = 1; *b = 0; }
    else { *r = 0; *g = 1; *b = 2; }
}

/*
 * get_intensity: dips/src/gpu/shaders/dips_shader.wgsl:64-82 (duplicates:
 * pre_compute_shader.wgsl:20-38, dips_alt/.../pre_compute_shader.wgsl:67-85).
 * WGSL returns color.{r,g,b} for CHROMA_FILTER 1/2/3, else (max+min)/2 on unorm floats.
 * Integer restatement: I2 = 510 * that value = max+min, or 2*channel.
 */
static inline uint16_t intensity2_rgb(unsigned r, unsigned g, unsigned b, int chroma) {
    if (chroma == DIPSO_CHROMA_RED) return (uint16_t)(2u * r);
    if (chroma == DIPSO_CHROMA_GREEN) return (uint16_t)(2u * g);
    if (chroma == DIPSO_CHROMA_BLUE) return (uint16_t)(2u * b);
    unsigned mx = r > g ? r : g; mx = mx > b ? mx : b;
    unsigned mn = r < g ? r : g; mn = mn < b ? mn : b;
    return (uint16_t)(mx + mn);
}

uint16_t dipso_intensity2(const uint8_t *px, int fmt, int chroma) {
    int ro, go, bo;
    channel_offsets(fmt, &ro, &go, &bo);
    return intensity2_rgb(px[ro], px[go], px[bo], chroma);
}

void dipso_i2_plane(const uint8_t *frame, size_t npx, int fmt, int chroma, uint16_t *out) {
    const int bpp = dipso_bytes_per_pixel(fmt);
    int ro, go, bo;
    channel_offsets(fmt, &ro, &go, &bo);
    for (size_t p = 0; p < npx; ++p) {
        const uint8_t *px = frame + p * (size_t)bpp;
        out[p] = intensity2_rgb(px[ro], px[go], px[bo], chroma);
    }
}

/* ascending sort of 4 (the in-bounds reading of the bubble sort at dips_shader.wgsl:196-211) */
static inline void sort4_u16(uint16_t v[4]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3 - i; ++j)
            if (v[j] > v[j + 1]) { uint16_t t = v[j]; v[j] = v[j + 1]; v[j + 1] = t; }
}

/* pre_compute_main: dips/src/gpu/shaders/pre_compute_shader.wgsl:103-131, element [MEDIAN_ARRAY_SIZE/2] */
void dipso_median4_plane(const uint8_t *const frames[4], size_t npx, int fmt, int chroma, uint16_t *out) {
    const int bpp = dipso_bytes_per_pixel(fmt);
    for (size_t p = 0; p < npx; ++p) {
        uint16_t v[4];
        for (int k = 0; k < 4; ++k) v[k] = dipso_intensity2(frames[k] + p * (size_t)bpp, fmt, chroma);
        sort4_u16(v);
        out[p] = v[2];
    }
}

/*
 * The hot path.  The subtraction and its sign follow compute_main,
 * dips_shader.wgsl:213-214 (start - current); the absolute value, threshold, accumulation and
 * per-frame scalars are the north-star extensions defined in SURVEY.md.
 */
void dipso_run_clip(const uint8_t *frames, size_t n_frames, size_t stride, size_t npx, int fmt,
                    int chroma, int mode, uint32_t tau, uint16_t *state, uint32_t *acc_sum,
                    uint32_t *acc_cnt, uint64_t *sad, uint64_t *cnt, int nthreads) {
    const int bpp = dipso_bytes_per_pixel(fmt);
    int ro, go, bo;
    channel_offsets(fmt, &ro, &go, &bo);
    int nt = 1;
#ifdef _OPENMP
    nt = nthreads > 0 ? nthreads : omp_get_max_threads();
#else
    (void)nthreads;
#endif
    if ((size_t)nt > npx) nt = npx ? (int)npx : 1;
    uint64_t *part = (uint64_t *)calloc((size_t)nt * n_frames * 2 + 1, sizeof(uint64_t));

#ifdef _OPENMP
#pragma omp parallel num_threads(nt)
#endif
    {
#ifdef _OPENMP
        const int tid = omp_get_thread_num();
#else
        const int tid = 0;
#endif
        const size_t p0 = npx * (size_t)tid / (size_t)nt, p1 = npx * (size_t)(tid + 1) / (size_t)nt;
        uint64_t *my = part + (size_t)tid * n_frames * 2;
        for (size_t t = 0; t < n_frames; ++t) {
            const uint8_t *f = frames + t * stride;
            uint64_t s = 0, c = 0;
            for (size_t p = p0; p < p1; ++p) {
                const uint8_t *px = f + p * (size_t)bpp;
                const int cur = intensity2_rgb(px[ro], px[go], px[bo], chroma);
                const int ref = state[p];
                const int d = cur > ref ? cur - ref : ref - cur;
                const unsigned m = (uint32_t)d > tau;
                acc_sum[p] += (uint32_t)d;
                acc_cnt[p] += m;
                s += (uint64_t)d;
                c += m;
                if (mode == DIPSO_MODE_PERFRAME) state[p] = (uint16_t)cur;
            }
            my[2 * t] = s;
            my[2 * t + 1] = c;
        }
    }
    for (size_t t = 0; t < n_frames; ++t) {
        uint64_t s = 0, c = 0;
        for (int k = 0; k < nt; ++k) {
            s += part[((size_t)k * n_frames + t) * 2];
            c += part[((size_t)k * n_frames + t) * 2 + 1];
        }
        sad[t] = s;
        cnt[t] = c;
    }
    free(part);
}

/* N4: see the header; insertion sort of at most 49 taps */
void dipso_spatial_median_plane(const uint16_t *in, uint32_t width, uint32_t height, int window, uint16_t *out) {
    const int r = window / 2, k = (window * window) / 2;
    if (window <= 1) { memcpy(out, in, (size_t)width * height * sizeof(uint16_t)); return; }
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int64_t y = 0; y < (int64_t)height; ++y) {
        for (int64_t x = 0; x < (int64_t)width; ++x) {
            uint16_t v[49];
            int n = 0;
            for (int dy = -r; dy <= r; ++dy)
                for (int dx = -r; dx <= r; ++dx) {
                    const int64_t yy = y + dy, xx = x + dx;
                    const uint16_t t = (yy >= 0 && yy < (int64_t)height && xx >= 0 && xx < (int64_t)width)
                                           ? in[(size_t)yy * width + (size_t)xx] : (uint16_t)0;
                    int i = n++;
                    while (i > 0 && v[i - 1] > t) { v[i] = v[i - 1]; --i; }
                    v[i] = t;
                }
            out[(size_t)y * width + (size_t)x] = v[k];
        }
    }
}

/* X6 */
void dipso_intensity_map(const uint32_t *acc_sum, size_t npx, uint64_t n_eff, float *out) {
    const double den = 510.0 * (double)(n_eff ? n_eff : 1);
    for (size_t p = 0; p < npx; ++p) out[p] = (float)((double)acc_sum[p] / den);
}

void dipso_frame_means(const uint64_t *sad, size_t n_frames, size_t npx, float *out) {
    const double den = 510.0 * (double)(npx ? npx : 1);
    for (size_t t = 0; t < n_frames; ++t) out[t] = (float)((double)sad[t] / den);
}

/* ---- the float visual chain ------------------------------------------------------------ */

/* hsl_to_rgb: dips_shader.wgsl:40-62 (h in degrees) */
static void hsl_to_rgb(float h, float s, float l, float rgb[3]) {
    const float chroma = s * (1.0f - fabsf(2.0f * l - 1.0f));
    const float hp = h / 60.0f;
    /* WGSL % on f32 is truncated remainder == fmodf */
    const float x = chroma * (1.0f - fabsf(fmodf(hp, 2.0f) - 1.0f));
    const float m = l - chroma / 2.0f;
    float r = 0.f, g = 0.f, b = 0.f;
    if (hp >= 0 && hp < 1) { r = chroma; g = x; }
    else if (hp >= 1 && hp < 2) { r = x; g = chroma; }
    else if (hp >= 2 && hp < 3) { g = chroma; b = x; }
    else if (hp >= 3 && hp < 4) { g = x; b = chroma; }
    else if (hp >= 4 && hp < 5) { r = x; b = chroma; }
    else if (hp >= 5 && hp <= 6) { r = chroma; b = x; }
    rgb[0] = r + m; rgb[1] = g + m; rgb[2] = b + m;
}

/* rgba8unorm store: clamp to [0,1], scale by 255, round to nearest (ties away from zero here;
 * WGSL/Vulkan leave the tie rule to the driver -- hence the +-1 LSB tolerance of X7). NaN -> 0. */
static inline uint8_t unorm8(float v) {
    if (!(v > 0.0f)) return 0;
    if (v >= 1.0f) return 255;
    return (uint8_t)floorf(v * 255.0f + 0.5f);
}

/* map (x0.5) -> sigmoid / inv_sigmoid -> x SENSITIVITY(5): dips_shader.wgsl:97-118, :217-229 */
static float visual_chain(float diff, int filter, float sig_scalar) {
    diff = diff * ((0.5f - -0.5f) / (1.0f - -1.0f));                     /* map(), :97-105, :217 */
    if (filter == DIPSO_FILTER_SIGMOID)
        diff = 1.0f / (1.0f + expf(-sig_scalar * diff)) - 0.5f;          /* sigmoid, :108-112 */
    else if (filter == DIPSO_FILTER_INV_SIGMOID)
        diff = (-logf((1.0f / (diff + 0.5f)) - 1.0f)) / sig_scalar;      /* inv_sigmoid, :114-118 */
    return diff * 5.0f;                                                  /* SENSITIVITY, :25, :229 */
}

float dipso_visual_diff(int32_t s_i2, int filter, float sig_scalar) {
    return visual_chain((float)s_i2 / 510.0f, filter, sig_scalar);
}

static void visual_from_diff(float diff, int colorize, uint8_t out[4]) {
    float rgb[3];
    if (colorize) {                                                      /* diff_to_color, :30-36 */
        if (diff < 0) hsl_to_rgb(0.0f, fabsf(diff), 0.5f, rgb);
        else hsl_to_rgb(120.0f, diff, 0.5f, rgb);
    } else {
        rgb[0] = rgb[1] = rgb[2] = 0.5f - diff;                          /* :236 */
    }
    out[0] = unorm8(rgb[0]); out[1] = unorm8(rgb[1]); out[2] = unorm8(rgb[2]); out[3] = 255; /* :239 */
}

void dipso_visual_pixel(int32_t s_i2, int colorize, int filter, float sig_scalar, uint8_t out[4]) {
    visual_from_diff(dipso_visual_diff(s_i2, filter, sig_scalar), colorize, out);
}

void dipso_visual_frame(const uint16_t *ref, const uint16_t *cur, size_t npx, int colorize, int filter,
                        float sig_scalar, uint8_t *out_rgba) {
    for (size_t p = 0; p < npx; ++p)
        dipso_visual_pixel((int32_t)ref[p] - (int32_t)cur[p], colorize, filter, sig_scalar, out_rgba + 4 * p);
}

/* ---- synthetic clips -------------------------------------------------------------------- */

uint64_t dipso_mix64(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}

#define DIPSO_GOLD 0x9E3779B97F4A7C15ull
#define DIPSO_BG_SALT 0xB5AD4ECEDA1CE2A9ull

static inline unsigned hash_byte(uint64_t seed, uint64_t index) {
    const uint64_t v = dipso_mix64(seed + ((index >> 3) + 1) * DIPSO_GOLD);
    return (unsigned)(v >> (8 * (index & 7))) & 0xFFu;
}

/* byte i of frame t.  g = t*frame_bytes + i is the byte's index in the tightly packed clip. */
static inline uint8_t synth_byte(uint64_t seed, int profile, uint64_t t, uint64_t i, uint64_t frame_bytes,
                                 uint32_t W, uint32_t H, int bpp) {
    const uint64_t g = t * frame_bytes + i;
    if (profile == DIPSO_SYNTH_UNIFORM) return (uint8_t)hash_byte(seed, g);
    /* scene: static background + noise in [-8,8] + moving block (+64), clamped */
    const int bg = (int)hash_byte(seed ^ DIPSO_BG_SALT, i);
    const int noise = (int)(hash_byte(seed, g) % 17u) - 8;
    const uint64_t p = i / (uint64_t)bpp;
    const uint32_t x = (uint32_t)(p % W), y = (uint32_t)(p / W);
    const uint32_t bx = (uint32_t)((t * 7u) % W), by = (uint32_t)((t * 3u) % H);
    const uint32_t bw = W * 5u / 16u, bh = H * 5u / 16u;
    const uint32_t dx = (x + W - bx) % W, dy = (y + H - by) % H;
    int v = bg + noise + ((dx < bw && dy < bh) ? 64 : 0);
    v = v < 0 ? 0 : (v > 255 ? 255 : v);
    return (uint8_t)v;
}

void dipso_synth_fill(uint8_t *dst, uint64_t first_frame, uint64_t n_frames, uint32_t width,
                      uint32_t height, int fmt, uint64_t seed, int profile, int nthreads) {
    const int bpp = dipso_bytes_per_pixel(fmt);
    const uint64_t fb = (uint64_t)width * height * (uint64_t)bpp;
    int nt = 1;
#ifdef _OPENMP
    nt = nthreads > 0 ? nthreads : omp_get_max_threads();
#else
    (void)nthreads;
#endif
    (void)nt;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nt) schedule(static)
#endif
    for (int64_t k = 0; k < (int64_t)n_frames; ++k) {
        uint8_t *f = dst + (uint64_t)k * fb;
        for (uint64_t i = 0; i < fb; ++i)
            f[i] = synth_byte(seed, profile, first_frame + (uint64_t)k, i, fb, width, height, bpp);
    }
}

/* ---- reference-flavour `dips` ComputeState (N1) ------------------------------------------- */

struct dipso_cs {
    uint32_t w, h;
    int colorize, filter, chroma;
    float sig_scalar;
    size_t seen;          /* frames pushed so far */
    int initialized;      /* main bind groups initialised (dips/src/gpu/mod.rs:191-214) */
    unsigned ring_index;  /* starting_temporal_index, bind_groups.rs:371, :407-427 */
    uint8_t *ring[4];     /* RGBA8 textures, TEMPORAL_BUFFER_SIZE = 4 (bind_groups.rs:18) */
    uint8_t *fifo[4];     /* host VecDeque of the last 4 frames (dips/src/gpu/mod.rs:171-175) */
    uint8_t *start;       /* start texture, RGBA8 grey */
};

dipso_cs *dipso_cs_new(uint32_t width, uint32_t height, int colorize, int filter, float sig_scalar,
                       int chroma) {
    dipso_cs *cs = (dipso_cs *)calloc(1, sizeof(*cs));
    const size_t fb = (size_t)width * height * 4;
    cs->w = width; cs->h = height; cs->colorize = colorize; cs->filter = filter;
    cs->chroma = chroma; cs->sig_scalar = sig_scalar;
    for (int k = 0; k < 4; ++k) { cs->ring[k] = (uint8_t *)malloc(fb); cs->fifo[k] = (uint8_t *)malloc(fb); }
    cs->start = (uint8_t *)malloc(fb);
    return cs;
}

void dipso_cs_free(dipso_cs *cs) {
    if (!cs) return;
    for (int k = 0; k < 4; ++k) { free(cs->ring[k]); free(cs->fifo[k]); }
    free(cs->start);
    free(cs);
}

/* get_intensity on an RGBA8 texel as f32 (dips_shader.wgsl:64-82) */
static inline float intensity_f32(const uint8_t *px, int chroma) {
    const float r = px[0] / 255.0f, g = px[1] / 255.0f, b = px[2] / 255.0f;
    if (chroma == DIPSO_CHROMA_RED) return r;
    if (chroma == DIPSO_CHROMA_GREEN) return g;
    if (chroma == DIPSO_CHROMA_BLUE) return b;
    const float mx = fmaxf(fmaxf(r, g), b), mn = fminf(fminf(r, g), b);
    return (mx + mn) / 2.0f;
}

static inline void sort4_f32(float v[4]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3 - i; ++j)
            if (v[j] > v[j + 1]) { float t = v[j]; v[j] = v[j + 1]; v[j + 1] = t; }
}

int dipso_cs_frame(dipso_cs *cs, const uint8_t *rgba_in, uint8_t *rgba_out) {
    const size_t npx = (size_t)cs->w * cs->h, fb = npx * 4;
    /* add_texture: push_back, pop_front when > 4 (dips/src/gpu/mod.rs:171-175) */
    if (cs->seen < 4) memcpy(cs->fifo[cs->seen], rgba_in, fb);
    else {
        uint8_t *old = cs->fifo[0];
        cs->fifo[0] = cs->fifo[1]; cs->fifo[1] = cs->fifo[2]; cs->fifo[2] = cs->fifo[3]; cs->fifo[3] = old;
        memcpy(cs->fifo[3], rgba_in, fb);
    }
    cs->seen++;
    if (cs->seen < 4) {                         /* dispatch() == None -> passthrough, lib.rs:241-245 */
        memcpy(rgba_out, rgba_in, fb);
        return 1;
    }
    if (!cs->initialized) {
        /* run_precompute_pipeline (mod.rs:218-304) over the 4 buffered frames, WINDOW_SIZE == 1:
         * start = grey(upper median of 4 intensities), stored rgba8unorm (pre_compute_shader.wgsl:103-131) */
        for (size_t p = 0; p < npx; ++p) {
            float v[4];
            for (int k = 0; k < 4; ++k) v[k] = intensity_f32(cs->fifo[k] + 4 * p, cs->chroma);
            sort4_f32(v);
            const uint8_t q = unorm8(v[2]);
            cs->start[4 * p] = cs->start[4 * p + 1] = cs->start[4 * p + 2] = q; cs->start[4 * p + 3] = 255;
        }
        /* MainComputeBindGroups::initialize uploads the 4 frames into ring slots 0..3, index 0 */
        for (int k = 0; k < 4; ++k) memcpy(cs->ring[k], cs->fifo[k], fb);
        cs->ring_index = 0;
        cs->initialized = 1;
    } else {
        /* update_temporal_texture (bind_groups.rs:407-427): write slot idx, publish idx, idx += 1 */
        memcpy(cs->ring[cs->ring_index], rgba_in, fb);
    }
    const unsigned slot = cs->ring_index;       /* value of the starting_index uniform for this dispatch */
    if (cs->seen > 4) cs->ring_index = (cs->ring_index + 1) % 4;
    /* compute_main (dips_shader.wgsl:172-240), WINDOW_SIZE == 1 */
    for (size_t p = 0; p < npx; ++p) {
        uint8_t *nw = cs->ring[slot] + 4 * p;
        const uint8_t q = unorm8(intensity_f32(nw, cs->chroma));       /* :187 in-place grey store */
        nw[0] = nw[1] = nw[2] = q; nw[3] = 255;
        float v[4];
        for (int k = 0; k < 4; ++k) v[k] = intensity_f32(cs->ring[k] + 4 * p, cs->chroma);   /* :191-193 */
        sort4_f32(v);                                                                          /* :196-211 */
        const float diff = cs->start[4 * p] / 255.0f - v[2];                                   /* :213-214 */
        visual_from_diff(visual_chain(diff, cs->filter, cs->sig_scalar), cs->colorize, rgba_out + 4 * p);
    }
    return 0;
}

/* ---- reference-flavour `dips_alt` DiPsCompute (N1) ----------------------------------------- */

struct dipso_alt {
    uint32_t w, h;
    int colorize, filter, chroma, intended;
    float sig_scalar;
    unsigned index;       /* texture_index, UCircularIndex(0, NUM_TEXTURES), dips_alt/src/dips_compute/mod.rs:519 */
    uint8_t *ring[2];     /* input_textures, RGBA8, zero-initialised */
    uint8_t *snap;        /* snapshot_texture, RGBA8, zero-initialised */
};

dipso_alt *dipso_alt_new(uint32_t width, uint32_t height, int colorize, int filter, float sig_scalar, int chroma,
                         int intended_median) {
    dipso_alt *a = (dipso_alt *)calloc(1, sizeof(*a));
    const size_t fb = (size_t)width * height * 4;
    a->w = width; a->h = height; a->colorize = colorize; a->filter = filter; a->chroma = chroma;
    a->intended = intended_median; a->sig_scalar = sig_scalar;
    a->ring[0] = (uint8_t *)calloc(fb, 1); a->ring[1] = (uint8_t *)calloc(fb, 1); a->snap = (uint8_t *)calloc(fb, 1);
    return a;
}

void dipso_alt_free(dipso_alt *a) {
    if (!a) return;
    free(a->ring[0]); free(a->ring[1]); free(a->snap); free(a);
}

void dipso_alt_frame(dipso_alt *a, const uint8_t *rgba_in, int snapshot, uint8_t *rgba_out) {
    const size_t npx = (size_t)a->w * a->h;
    memcpy(a->ring[a->index], rgba_in, npx * 4);             /* write_texture into slot index, then index += 1 (:507-521) */
    a->index = (a->index + 1) % 2;
    for (size_t p = 0; p < npx; ++p) {
        /* median_array[i] = intensity of texture i (WINDOW_SIZE == 1), 16-slot array otherwise zero (:201-211) */
        float v[3] = { intensity_f32(a->ring[0] + 4 * p, a->chroma), intensity_f32(a->ring[1] + 4 * p, a->chroma), 0.0f };
        float med;
        if (a->intended) {
            med = v[0] > v[1] ? v[0] : v[1];                 /* sorted[N/2] of the two real values */
        } else {
            /* as shipped: i,j < NUM_TEXTURES compare [j],[j+1] -- the zero in slot 2 takes part (:212-227) */
            for (int i = 0; i < 2; ++i) {
                int swapped = 0;
                for (int j = 0; j < 2; ++j)
                    if (v[j] > v[j + 1]) { float t = v[j]; v[j] = v[j + 1]; v[j + 1] = t; swapped = 1; }
                if (!swapped) break;
            }
            med = v[1];                                      /* median_array[NUM_TEXTURES / 2] */
        }
        if (snapshot) {                                      /* :231-235 */
            const uint8_t q = unorm8(med);
            uint8_t *s = a->snap + 4 * p, *o = rgba_out + 4 * p;
            s[0] = s[1] = s[2] = q; s[3] = 255;
            o[0] = o[1] = o[2] = q; o[3] = 255;
        } else {                                             /* :236-262 */
            const float diff = a->snap[4 * p] / 255.0f - med;
            visual_from_diff(visual_chain(diff, a->filter, a->sig_scalar), a->colorize, rgba_out + 4 * p);
        }
    }
}
