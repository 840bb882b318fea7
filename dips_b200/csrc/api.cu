// api.cu -- the C ABI of libdips_b200.so (include/dips_b200.h): context, planning, launches, copies.
// Host-side runtime only; the device code is in clip_kernel.cu / aux_kernels.cu.  No CPU fallback anywhere: every
// result is produced by a CUDA kernel, and dipsb_create fails when there is no usable device.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "dipsb_ctx.h"

namespace dipsb {
uint64_t launch_count_value();
}
using namespace dipsb;

int dipsb::bpp_of(int format) { return (format == DIPSB_FMT_RGB8 || format == DIPSB_FMT_BGR8) ? 3 : 4; }

static int chan_byte_of(int format, int chroma) {
    if (chroma == DIPSB_CHROMA_NONE) return -1;
    const bool bgr = (format == DIPSB_FMT_BGR8 || format == DIPSB_FMT_BGRX8);
    const int rgb_index = chroma - 1;  // 0 R, 1 G, 2 B
    return bgr ? 2 - rgb_index : rgb_index;
}

static bool windowed(const dipsb_ctx* c);
static int32_t ensure_i2_scratch(dipsb_ctx* c);
// spatially filtered intensity plane of one tightly packed frame (N4): raw I2 -> correct median of the w x w window
static int32_t filtered_plane(dipsb_ctx* c, const uint8_t* d_frame, int format, uint16_t* out);

// ---- planning ------------------------------------------------------------------------------------------------------
// Measured on B200 (profiles/r01_sweeps.md): the clip kernels run fastest with ONE large block per SM -- every
// SM streams one contiguous 40-56 KB slice of each frame through a 3-4 deep TMA ring -- and they are co-limited by HBM and
// by instruction issue, so every SM must get the same number of pixels.  Plan: tile_px = npx / (num_sms * waves) rounded
// up to 16 pixels, with the fewest waves that fit a block (clip_kernel_ws: 992 consumer threads at 64 registers;
// clip_kernel: 896 threads at 72 registers or 1024 at 64); 4 pipeline stages unless the ring would exceed 200 KB.  Frames
// too small to give every SM a 2048-pixel tile keep 2048-pixel tiles, run several blocks per SM and are split into frame
// segments instead (plan_segments).
static bool plan_geometry(Geometry& g, uint32_t force_stages, uint32_t force_tile_px, uint32_t force_regs = 0, int kernel = 0) {
    const uint64_t min_tile = std::min<uint64_t>(2048, (g.npx + 15) / 16 * 16);
    if (kernel == 1) {   // warp-specialised variant: 64 registers, 16 px/thread, block = consumers + one producer warp
        if (force_regs && force_regs != 64) return false;
        if (force_stages && force_stages != 3 && force_stages != 4) return false;
        force_regs = 64;
    }
    for (int regs : {72, 64, 80, 96, 128}) {
        if (force_regs ? (uint32_t)regs != force_regs : regs > 72) continue;   // 80/96/128 only on request (tuning)
        const int groups = clip_groups(regs);
        const uint32_t px_thr = (uint32_t)(kPxPerThread * groups);
        const uint32_t max_thr = (uint32_t)clip_max_threads_per_sm(regs) - (kernel == 1 ? 32u : 0u);
        const uint64_t max_slots = (uint64_t)max_thr * px_thr;
        uint64_t tile_px;
        if (force_tile_px) tile_px = force_tile_px;
        else {
            const uint64_t waves = (g.npx + g.num_sms * max_slots - 1) / (g.num_sms * max_slots);
            tile_px = (g.npx + g.num_sms * waves - 1) / (g.num_sms * waves);
            tile_px = std::max<uint64_t>((tile_px + 15) / 16 * 16, min_tile);
            // 72 registers (896 threads) unless only the 64-register variant (1024 threads) saves a whole wave
            if (!force_regs && regs == 72 && kernel == 0) {
                const uint64_t slots64 = 1024ull * kPxPerThread;
                const uint64_t waves64 = (g.npx + g.num_sms * slots64 - 1) / (g.num_sms * slots64);
                if (waves64 < waves) continue;
            }
        }
        if (tile_px > max_slots) continue;
        const uint32_t thr = (uint32_t)((tile_px + px_thr * 32 - 1) / (px_thr * 32)) * 32;
        uint32_t stages = force_stages ? force_stages : 4;
        // measured (profiles/r01_sweeps.md): 4 stages beat 3 for 42-49 KB slices, but a 4 x 56 KB ring (4K RGBx, 224 KB: all
        // of the SM's shared memory) is slower than 3 x 56 KB -- keep the ring <= 200 KB
        if (!force_stages && clip_smem_bytes(thr, g.bpp, stages, regs) > 200 * 1024) stages = 3;
        int occ = clip_occupancy(thr + (kernel == 1 ? 32u : 0u), g.bpp, stages, regs);
        while (!force_stages && occ <= 0 && stages > (kernel == 1 ? 3u : 2u)) occ = clip_occupancy(thr + (kernel == 1 ? 32u : 0u), g.bpp, --stages, regs);
        if (occ <= 0) continue;
        g.threads = thr;
        g.tile_px = (uint32_t)tile_px;
        g.n_tiles = (uint32_t)((g.npx + tile_px - 1) / tile_px);
        g.n_elems = (uint64_t)g.n_tiles * thr * px_thr;
        g.state_elems = g.npx + (uint64_t)thr * px_thr;
        g.stages = stages;
        g.blocks_per_sm = (uint32_t)occ;
        g.regs = regs;
        g.groups = groups;
        g.kernel = kernel;
        return true;
    }
    return false;
}

static uint32_t plan_segments(const dipsb_ctx* c, uint64_t n_frames) {
    if (c->tune_segments) return (uint32_t)std::min<uint64_t>(c->tune_segments, n_frames);
    const uint64_t slots = (uint64_t)c->g.blocks_per_sm * c->g.num_sms;
    if (c->g.n_tiles >= slots) return 1;
    uint64_t segs = slots / c->g.n_tiles;                 // fill the resident slots once
    const uint64_t by_len = std::max<uint64_t>(1, n_frames / 32);   // keep segments >= 32 frames (flush + ref overhead)
    segs = std::min(segs, by_len);
    return (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(segs, n_frames));
}

static bool windowed(const dipsb_ctx* c) { return c->cfg.spatial_window > 1; }

static int32_t ensure_i2_scratch(dipsb_ctx* c) {
    if (!c->i2_scratch) CK(c, cudaMalloc(&c->i2_scratch, 5 * c->g.npx * sizeof(uint16_t)));
    return DIPSB_OK;
}

static int32_t filtered_plane(dipsb_ctx* c, const uint8_t* d_frame, int format, uint16_t* out) {
    const Geometry& g = c->g;
    int32_t rc0 = ensure_i2_scratch(c);
    if (rc0) return rc0;
    Geometry gf = g;                       // the frame may come in another pixel format than the context's (push_frame)
    gf.bpp = bpp_of(format);
    gf.chan_byte = chan_byte_of(format, c->cfg.chroma);
    CK(c, launch_prime(gf, d_frame, c->i2_scratch, c->stream));
    CK(c, launch_spatial_median(g, c->i2_scratch, out, c->cfg.spatial_window, c->stream));
    return DIPSB_OK;
}

// clip_kernel_ws (producer warp, stage-unrolled; +7-10 % sustained on 3 B/px clips, equal elsewhere: profiles/r01_sweeps.md)
// exists for 64 registers and 3 or 4 stages; any other forced tuning selects clip_kernel
// (and not for 4 B/px frames with a chroma filter in per-frame mode: those six instantiations do not fit 64 registers)
static int pick_kernel(const dipsb_ctx* c, int requested, uint32_t stages, uint32_t regs) {
    const bool ws_exists = clip_ws_available(c->g.bpp, c->g.chan_byte, c->cfg.mode);
    if (requested >= 0) return (requested == 1 && !ws_exists) ? -1 : requested;
    return (ws_exists && (regs == 0 || regs == 64) && (stages == 0 || stages == 3 || stages == 4)) ? 1 : 0;
}

// ---- lifetime ------------------------------------------------------------------------------------------------------
extern "C" int32_t dipsb_abi_version(void) { return DIPSB_ABI_VERSION; }

extern "C" void dipsb_default_config(dipsb_config* cfg) {
    if (!cfg) return;
    memset(cfg, 0, sizeof *cfg);
    cfg->struct_size = (uint32_t)sizeof *cfg;
    cfg->format = DIPSB_FMT_RGBX8;
    cfg->mode = DIPSB_MODE_OVERALL;
    cfg->chroma = DIPSB_CHROMA_NONE;
    cfg->filter = DIPSB_FILTER_NONE;       // DiPsProperties::new(): Unfiltered, sensitivity 5, window 1 (dips/src/lib.rs:75-86)
    cfg->sigmoid_scalar = 5.0f;
    cfg->spatial_window = 1;
}

static void free_all(dipsb_ctx* c) {
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (int k = 0; k < 2; ++k) {
        if (c->state[k]) cudaFree(c->state[k]);
        if (c->ev_copy[k]) cudaEventDestroy(c->ev_copy[k]);
        if (c->ev_done[k]) cudaEventDestroy(c->ev_done[k]);
        if (c->h_chunk[k]) cudaFreeHost(c->h_chunk[k]);
        if (c->d_chunk[k]) cudaFree(c->d_chunk[k]);
    }
    cudaFree(c->acc); cudaFree(c->planar); cudaFree(c->d_sad); cudaFree(c->d_cnt); cudaFree(c->partials);
    cudaFree(c->d_repack); cudaFree(c->d_frame); cudaFree(c->d_rgba); cudaFree(c->ring); cudaFree(c->i2_scratch); cudaFree(c->xchg);
    if (c->h_pin) cudaFreeHost(c->h_pin);
    if (c->h_stat) cudaFreeHost(c->h_stat);
    for (cudaEvent_t e : c->tev) cudaEventDestroy(e);
    for (cudaEvent_t e : c->pev) cudaEventDestroy(e);
    if (c->ev_switch) cudaEventDestroy(c->ev_switch);
    for (auto& sl : c->slot) {
        if (sl.h_in) cudaFreeHost(sl.h_in);
        if (sl.h_out) cudaFreeHost(sl.h_out);
        if (sl.h_stat) cudaFreeHost(sl.h_stat);
        cudaFree(sl.d_in); cudaFree(sl.d_out);
        if (sl.ev_done) cudaEventDestroy(sl.ev_done);
        for (auto& e : sl.ev_in) if (e) cudaEventDestroy(e);
        for (auto& e : sl.ev_k) if (e) cudaEventDestroy(e);
        for (auto& e : sl.ev_out) if (e) cudaEventDestroy(e);
    }
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->out_stream) cudaStreamDestroy(c->out_stream);
}

static int32_t alloc_planes(dipsb_ctx* c) {
    const Geometry& g = c->g;
    for (int k = 0; k < 2; ++k) {
        CK(c, cudaMalloc(&c->state[k], g.state_elems * sizeof(uint16_t)));
        CK(c, cudaMemsetAsync(c->state[k], 0, g.state_elems * sizeof(uint16_t), c->stream));
    }
    CK(c, cudaMalloc(&c->acc, 2 * g.n_elems * sizeof(uint32_t)));
    CK(c, cudaMemsetAsync(c->acc, 0, 2 * g.n_elems * sizeof(uint32_t), c->stream));
    c->acc_zero_pending = false;
    CK(c, cudaMalloc(&c->planar, 2 * g.npx * sizeof(uint32_t)));
    if (c->cfg.flavor != DIPSB_FLAVOR_FRAME0 && !c->ring) {
        CK(c, cudaMalloc(&c->ring, 4 * g.npx * sizeof(uint16_t)));
        CK(c, cudaMemsetAsync(c->ring, 0, 4 * g.npx * sizeof(uint16_t), c->stream));   // wgpu textures start zeroed
    }
    return DIPSB_OK;
}

extern "C" int32_t dipsb_create(const dipsb_config* cfg, dipsb_ctx** out) {
    if (!cfg || !out) return fail(nullptr, DIPSB_ERR_INVALID, "dipsb_create: null argument");
    *out = nullptr;
    if (cfg->struct_size != sizeof(dipsb_config))
        return fail(nullptr, DIPSB_ERR_INVALID, "dipsb_create: struct_size %u != %zu", cfg->struct_size, sizeof(dipsb_config));
    if (cfg->width == 0 || cfg->height == 0) return fail(nullptr, DIPSB_ERR_INVALID, "dipsb_create: empty frame");
    if (cfg->format < 0 || cfg->format > 3) return fail(nullptr, DIPSB_ERR_INVALID, "dipsb_create: bad format %d", cfg->format);
    if (cfg->mode != DIPSB_MODE_OVERALL && cfg->mode != DIPSB_MODE_PERFRAME)
        return fail(nullptr, DIPSB_ERR_INVALID, "dipsb_create: bad mode %d", cfg->mode);
    if (cfg->chroma < 0 || cfg->chroma > 3) return fail(nullptr, DIPSB_ERR_INVALID, "dipsb_create: bad chroma %d", cfg->chroma);
    if (cfg->flavor < 0 || cfg->flavor > 3) return fail(nullptr, DIPSB_ERR_INVALID, "dipsb_create: bad flavor %d", cfg->flavor);
    if (cfg->flavor != DIPSB_FLAVOR_FRAME0 && cfg->mode != DIPSB_MODE_OVERALL)
        return fail(nullptr, DIPSB_ERR_INVALID, "dipsb_create: the ring flavours are overall-mode only");
    if (cfg->spatial_window != 0 && cfg->spatial_window != 1 && cfg->spatial_window != 3 && cfg->spatial_window != 5 && cfg->spatial_window != 7)
        return fail(nullptr, DIPSB_ERR_INVALID, "dipsb_create: spatial_window %d not one of 1, 3, 5, 7", cfg->spatial_window);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, DIPSB_ERR_CUDA, "dipsb_create: no CUDA device (%s); there is no CPU fallback", cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, DIPSB_ERR_INVALID, "dipsb_create: device %d of %d", cfg->device, ndev);

    dipsb_ctx* c = new (std::nothrow) dipsb_ctx();
    if (!c) return fail(nullptr, DIPSB_ERR_NOMEM, "dipsb_create: out of host memory");
    c->cfg = *cfg;
    c->device = cfg->device;
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(c->device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, c->device)) != cudaSuccess) {
        delete c;
        return fail(nullptr, DIPSB_ERR_CUDA, "dipsb_create: cudaSetDevice/GetDeviceProperties: %s", cudaGetErrorString(e));
    }
    if (prop.major < 10) {
        delete c;
        return fail(nullptr, DIPSB_ERR_CUDA, "dipsb_create: device sm_%d%d is not Blackwell (sm_100a required)", prop.major, prop.minor);
    }
    Geometry& g = c->g;
    g.width = cfg->width; g.height = cfg->height; g.npx = (uint64_t)cfg->width * cfg->height;
    g.format = cfg->format; g.bpp = bpp_of(cfg->format); g.chan_byte = chan_byte_of(cfg->format, cfg->chroma);
    g.num_sms = (uint32_t)prop.multiProcessorCount;
    if (!plan_geometry(g, 0, 0, 0, clip_ws_available(g.bpp, g.chan_byte, cfg->mode) ? 1 : 0)) {
        delete c;
        return fail(nullptr, DIPSB_ERR_INVALID, "dipsb_create: no kernel geometry fits %ux%u", cfg->width, cfg->height);
    }
    int32_t rc = DIPSB_OK;
    if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->out_stream, cudaStreamNonBlocking) != cudaSuccess)
        rc = fail(nullptr, DIPSB_ERR_CUDA, "dipsb_create: stream creation failed");
    c->stream = c->own_stream;
    for (int k = 0; k < 2 && rc == DIPSB_OK; ++k)
        if (cudaEventCreateWithFlags(&c->ev_copy[k], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->ev_done[k], cudaEventDisableTiming) != cudaSuccess)
            rc = fail(nullptr, DIPSB_ERR_CUDA, "dipsb_create: event creation failed");
    if (rc == DIPSB_OK && cudaEventCreateWithFlags(&c->ev_switch, cudaEventDisableTiming) != cudaSuccess)
        rc = fail(nullptr, DIPSB_ERR_CUDA, "dipsb_create: event creation failed");
    if (rc == DIPSB_OK) rc = alloc_planes(c);
    if (rc == DIPSB_OK && cudaMallocHost(&c->h_stat, 2 * sizeof(uint64_t)) != cudaSuccess)
        rc = fail(c, DIPSB_ERR_NOMEM, "dipsb_create: pinned allocation failed");
    if (rc != DIPSB_OK) {
        g_create_err = c->err.empty() ? g_create_err : c->err;
        free_all(c);
        delete c;
        return rc;
    }
    *out = c;
    return DIPSB_OK;
}

extern "C" void dipsb_destroy(dipsb_ctx* c) {
    if (!c) return;
    comm_detach(c);
    free_all(c);
    delete c;
}

extern "C" const char* dipsb_last_error(const dipsb_ctx* c) { return c ? c->err.c_str() : g_create_err.c_str(); }

int32_t dipsb::finalize_pending(dipsb_ctx* c) {
    if (!c->fin.pending) return DIPSB_OK;
    c->fin.pending = false;
    CK(c, launch_finalize_scalars(c->g, c->partials, c->fin.n, c->fin.words, c->d_sad + c->fin.first, c->d_cnt + c->fin.first, c->stream));
    return DIPSB_OK;
}

int32_t dipsb::ensure_acc_zero(dipsb_ctx* c) {
    if (!c->acc_zero_pending) return DIPSB_OK;
    CK(c, cudaMemsetAsync(c->acc, 0, 2 * c->g.n_elems * sizeof(uint32_t), c->stream));
    c->acc_zero_pending = false;
    return DIPSB_OK;
}

extern "C" int32_t dipsb_reset(dipsb_ctx* c) {
    if (!c) return DIPSB_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    const Geometry& g = c->g;
    // The accumulator planes (8 bytes per pixel) are not cleared here: the clip kernel that normally follows clears each
    // thread's words in its prologue, under the fill of its TMA ring; every other user of the planes clears them first.
    c->acc_zero_pending = true;
    if (c->ring) {
        CK(c, cudaMemsetAsync(c->ring, 0, 4 * g.npx * sizeof(uint16_t), c->stream));
        for (int k = 0; k < 2; ++k) CK(c, cudaMemsetAsync(c->state[k], 0, g.state_elems * sizeof(uint16_t), c->stream));
    }
    c->ring_seen = 0; c->ring_index = 0;
    // the state planes need no clearing: the next run primes [0, npx) and the zero padding past npx is never written
    if (c->scal_cap) {
        CK(c, cudaMemsetAsync(c->d_sad, 0, c->scal_cap * sizeof(uint64_t), c->stream));
        CK(c, cudaMemsetAsync(c->d_cnt, 0, c->scal_cap * sizeof(uint64_t), c->stream));
    }
    c->state_valid = false; c->snapshot_pending = false; c->state_cur = 0;
    c->frames_processed = 0; c->stream_index = 0; c->scal_hi = 0;
    c->acc_sharded = false;
    return DIPSB_OK;
}

extern "C" int32_t dipsb_set_threshold(dipsb_ctx* c, uint32_t threshold) {
    if (!c) return DIPSB_ERR_INVALID;
    c->cfg.threshold = threshold;
    return DIPSB_OK;
}

extern "C" int32_t dipsb_set_stream(dipsb_ctx* c, void* stream) {
    if (!c) return DIPSB_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    if ((cudaStream_t)stream == c->stream) return DIPSB_OK;
    // keep the ordering of already issued work without blocking the host: the new stream waits for the old one
    CK(c, cudaEventRecord(c->ev_switch, c->stream));
    CK(c, cudaStreamWaitEvent((cudaStream_t)stream, c->ev_switch, 0));
    c->stream = (cudaStream_t)stream;
    return DIPSB_OK;
}

extern "C" int32_t dipsb_adopt_stream(dipsb_ctx* c, void* stream) {
    if (!c) return DIPSB_ERR_INVALID;
    c->stream = (cudaStream_t)stream;   // no ordering: the caller has ordered the two streams with its own events
    return DIPSB_OK;
}

extern "C" int32_t dipsb_use_private_stream(dipsb_ctx* c) {
    if (!c) return DIPSB_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    if (c->stream == c->own_stream) return DIPSB_OK;
    CK(c, cudaEventRecord(c->ev_switch, c->stream));
    CK(c, cudaStreamWaitEvent(c->own_stream, c->ev_switch, 0));
    c->stream = c->own_stream;
    return DIPSB_OK;
}

extern "C" int32_t dipsb_synchronize(dipsb_ctx* c) {
    if (!c) return DIPSB_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaStreamSynchronize(c->stream));
    return DIPSB_OK;
}

extern "C" int32_t dipsb_set_tuning(dipsb_ctx* c, uint32_t stages, uint32_t tile_px, uint32_t segments, uint32_t regs) {
    if (!c) return DIPSB_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    if (stages > (uint32_t)kMaxStages || (stages && stages < 2)) return fail(c, DIPSB_ERR_INVALID, "set_tuning: stages %u outside [2,%d]", stages, kMaxStages);
    if (tile_px && (tile_px % 16 || tile_px > 1024u * kPxPerThread))
        return fail(c, DIPSB_ERR_INVALID, "set_tuning: tile_px %u must be a multiple of 16 and <= %d", tile_px, 1024 * kPxPerThread);
    if (regs && regs != 64 && regs != 72 && regs != 80 && regs != 96 && regs != 128) return fail(c, DIPSB_ERR_INVALID, "set_tuning: regs %u not one of 64/72/80/96/128", regs);
    const bool regeo = (stages != c->tune_stages) || (tile_px != c->tune_tile_px) || (regs != c->tune_regs);
    c->tune_segments = segments;
    if (!regeo) return DIPSB_OK;
    if (c->frames_processed != 0) return fail(c, DIPSB_ERR_STATE, "set_tuning: geometry can only change on a fresh or reset context");
    if (c->comm) return fail(c, DIPSB_ERR_STATE, "set_tuning: the geometry is fixed once the context has a communicator (its planes are mapped by the peers)");
    CK(c, cudaStreamSynchronize(c->stream));
    Geometry g = c->g;
    if (!plan_geometry(g, stages, tile_px, regs, pick_kernel(c, c->tune_kernel, stages, regs) < 0 ? 0 : pick_kernel(c, c->tune_kernel, stages, regs))) return fail(c, DIPSB_ERR_INVALID, "set_tuning: tile_px %u / stages %u / regs %u do not fit", tile_px, stages, regs);
    for (int k = 0; k < 2; ++k) { cudaFree(c->state[k]); c->state[k] = nullptr; }
    cudaFree(c->acc); c->acc = nullptr;
    cudaFree(c->planar); c->planar = nullptr;
    cudaFree(c->xchg); c->xchg = nullptr;
    c->g = g;
    c->tune_stages = stages; c->tune_tile_px = tile_px; c->tune_regs = regs;
    c->state_valid = false;
    return alloc_planes(c);
}

// host-only: the plan the library would use for a geometry on a device with num_sms SMs (no device is touched)
extern "C" int32_t dipsb_plan_query(uint32_t width, uint32_t height, int32_t format, uint32_t num_sms, uint32_t out[8]) {
    if (!out || !width || !height || format < 0 || format > 3 || !num_sms) return DIPSB_ERR_INVALID;
    Geometry g{};
    g.width = width; g.height = height; g.npx = (uint64_t)width * height;
    g.format = format; g.bpp = bpp_of(format); g.chan_byte = -1; g.num_sms = num_sms;
    if (!plan_geometry(g, 0, 0, 0, 1)) return DIPSB_ERR_INVALID;
    out[0] = g.n_tiles; out[1] = 0; out[2] = g.threads; out[3] = g.stages | ((uint32_t)g.kernel << 16); out[4] = g.blocks_per_sm; out[5] = g.tile_px;
    out[6] = (uint32_t)clip_smem_bytes(g.threads, g.bpp, g.stages, g.regs) | ((uint32_t)g.regs << 24); out[7] = clip_active_warps(g);
    return DIPSB_OK;
}

extern "C" int32_t dipsb_set_kernel(dipsb_ctx* c, int32_t kernel) {
    if (!c) return DIPSB_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    if (kernel < -1 || kernel > 1) return fail(c, DIPSB_ERR_INVALID, "set_kernel: %d is not -1 (automatic), 0 (clip_kernel) or 1 (clip_kernel_ws)", kernel);
    const int requested = kernel;
    kernel = pick_kernel(c, requested, c->tune_stages, c->tune_regs);
    if (kernel < 0) return fail(c, DIPSB_ERR_INVALID, "set_kernel: clip_kernel_ws has no variant for 4 B/px frames with a chroma filter in per-frame mode");
    if (kernel == c->g.kernel) { c->tune_kernel = requested; return DIPSB_OK; }
    if (c->frames_processed != 0) return fail(c, DIPSB_ERR_STATE, "set_kernel: only on a fresh or reset context");
    if (c->comm) return fail(c, DIPSB_ERR_STATE, "set_kernel: the geometry is fixed once the context has a communicator");
    CK(c, cudaStreamSynchronize(c->stream));
    Geometry g = c->g;
    if (!plan_geometry(g, c->tune_stages, c->tune_tile_px, c->tune_regs, kernel))
        return fail(c, DIPSB_ERR_INVALID, "set_kernel: the current tuning (stages %u, regs %u) does not fit kernel %d", c->tune_stages, c->tune_regs, kernel);
    for (int k = 0; k < 2; ++k) { cudaFree(c->state[k]); c->state[k] = nullptr; }
    cudaFree(c->acc); c->acc = nullptr;
    cudaFree(c->planar); c->planar = nullptr;
    cudaFree(c->xchg); c->xchg = nullptr;
    c->g = g;
    c->tune_kernel = requested;
    c->state_valid = false;
    return alloc_planes(c);
}

extern "C" int32_t dipsb_last_plan(const dipsb_ctx* c, uint32_t out[8]) {
    if (!c || !out) return DIPSB_ERR_INVALID;
    memcpy(out, c->last_plan, sizeof c->last_plan);
    out[0] = c->g.n_tiles; out[2] = c->g.threads; out[3] = c->g.stages | ((uint32_t)c->g.kernel << 16); out[4] = c->g.blocks_per_sm; out[5] = c->g.tile_px;
    out[6] = (uint32_t)clip_smem_bytes(c->g.threads, c->g.bpp, c->g.stages, c->g.regs) | ((uint32_t)c->g.regs << 24);
    return DIPSB_OK;
}

// ---- state plane ---------------------------------------------------------------------------------------------------
extern "C" int32_t dipsb_prime_device(dipsb_ctx* c, const void* d_frame) {
    if (!c || !d_frame) return DIPSB_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    if (windowed(c)) {
        int32_t rc = filtered_plane(c, (const uint8_t*)d_frame, c->g.format, c->state[c->state_cur]);
        if (rc) return rc;
    } else {
        CK(c, launch_prime(c->g, (const uint8_t*)d_frame, c->state[c->state_cur], c->stream));
    }
    c->state_valid = true;
    c->snapshot_pending = false;
    return DIPSB_OK;
}

extern "C" int32_t dipsb_prime_median4_device(dipsb_ctx* c, const void* d_frames, uint64_t stride) {
    if (!c || !d_frames) return DIPSB_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    if (stride < c->g.npx * c->g.bpp) return fail(c, DIPSB_ERR_INVALID, "prime_median4: stride smaller than a frame");
    if (windowed(c)) {   // each start frame is spatially filtered first (pre_compute_shader.wgsl:104-107), then the upper median
        int32_t rc0 = ensure_i2_scratch(c);
        if (rc0) return rc0;
        for (int k = 0; k < 4; ++k) {
            int32_t rc = filtered_plane(c, (const uint8_t*)d_frames + k * stride, c->g.format, c->i2_scratch + (1 + k) * c->g.npx);
            if (rc) return rc;
        }
        CK(c, launch_median4_planes(c->g, c->i2_scratch + c->g.npx, c->state[c->state_cur], c->stream));
    } else {
        CK(c, launch_prime_median4(c->g, (const uint8_t*)d_frames, stride, c->state[c->state_cur], c->stream));
    }
    c->state_valid = true;
    c->snapshot_pending = false;
    return DIPSB_OK;
}

static int32_t ensure_frame_staging(dipsb_ctx* c, size_t in_bytes) {
    const size_t rgba = c->g.npx * 4;
    const size_t need_pin = std::max(in_bytes, rgba);
    if (c->d_frame_bytes < in_bytes) {
        cudaFree(c->d_frame); c->d_frame = nullptr;
        CK(c, cudaMalloc(&c->d_frame, in_bytes));
        c->d_frame_bytes = in_bytes;
    }
    if (!c->d_rgba) CK(c, cudaMalloc(&c->d_rgba, rgba));
    if (c->h_pin_bytes < need_pin) {
        if (c->h_pin) cudaFreeHost(c->h_pin);
        c->h_pin = nullptr;
        CK(c, cudaMallocHost(&c->h_pin, need_pin));
        c->h_pin_bytes = need_pin;
    }
    return DIPSB_OK;
}

extern "C" int32_t dipsb_prime_host(dipsb_ctx* c, const uint8_t* frame) {
    if (!c || !frame) return DIPSB_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    const size_t fb = c->g.npx * c->g.bpp;
    int32_t rc = ensure_frame_staging(c, fb);
    if (rc) return rc;
    host_copy2d(c->h_pin, fb, frame, fb, fb, 1);
    CK(c, cudaMemcpyAsync(c->d_frame, c->h_pin, fb, cudaMemcpyHostToDevice, c->stream));
    rc = dipsb_prime_device(c, c->d_frame);
    if (rc) return rc;
    CK(c, cudaStreamSynchronize(c->stream));
    return DIPSB_OK;
}

extern "C" int32_t dipsb_state_plane_device(dipsb_ctx* c, void** d_state) {
    if (!c || !d_state) return DIPSB_ERR_INVALID;
    *d_state = c->state[c->state_cur];
    return DIPSB_OK;
}

extern "C" int32_t dipsb_mark_state_valid(dipsb_ctx* c, int32_t valid) {
    if (!c) return DIPSB_ERR_INVALID;
    c->state_valid = valid != 0;
    return DIPSB_OK;
}

extern "C" int32_t dipsb_get_state_plane(dipsb_ctx* c, uint16_t* out) {
    if (!c || !out) return DIPSB_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaMemcpyAsync(out, c->state[c->state_cur], c->g.npx * sizeof(uint16_t), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return DIPSB_OK;
}

// ---- scalars storage -----------------------------------------------------------------------------------------------
int32_t dipsb::ensure_scalars(dipsb_ctx* c, uint64_t upto) {
    if (upto <= c->scal_cap) return DIPSB_OK;
    uint64_t cap = std::max<uint64_t>(1024, c->scal_cap);
    while (cap < upto) cap *= 2;
    uint64_t *ns = nullptr, *nc = nullptr;
    CK(c, cudaMalloc(&ns, cap * sizeof(uint64_t)));
    CK(c, cudaMalloc(&nc, cap * sizeof(uint64_t)));
    CK(c, cudaMemsetAsync(ns, 0, cap * sizeof(uint64_t), c->stream));
    CK(c, cudaMemsetAsync(nc, 0, cap * sizeof(uint64_t), c->stream));
    if (c->scal_cap) {
        CK(c, cudaMemcpyAsync(ns, c->d_sad, c->scal_cap * sizeof(uint64_t), cudaMemcpyDeviceToDevice, c->stream));
        CK(c, cudaMemcpyAsync(nc, c->d_cnt, c->scal_cap * sizeof(uint64_t), cudaMemcpyDeviceToDevice, c->stream));
        CK(c, cudaStreamSynchronize(c->stream));
        cudaFree(c->d_sad); cudaFree(c->d_cnt);
    }
    c->d_sad = ns; c->d_cnt = nc; c->scal_cap = cap;
    return DIPSB_OK;
}

// ---- batch ---------------------------------------------------------------------------------------------------------
// What the next frame does to a reference-flavour ring (N1): which slot it overwrites, whether the start plane / snapshot is
// (re)computed, whether a difference is produced.  Advances the context's ring bookkeeping; returns true while the frame is
// only passed through (the `dips` warm-up).  Shared by the per-frame call and the batch call.
static bool ring_next_frame(dipsb_ctx* c, RingArgs& r) {
    r.grey_slot = -1; r.compute_start = 0; r.snapshot = 0; r.median_is_max = 0; r.do_diff = 1;
    if (c->cfg.flavor == DIPSB_FLAVOR_DIPS_RING4) {
        if (c->snapshot_pending) { c->ring_seen = 0; c->ring_index = 0; c->snapshot_pending = false; }   // restart the warm-up
        r.n_slots = 4;
        const uint32_t seen = ++c->ring_seen;                     // frames including this one
        if (seen < 4) {                                            // dispatch() == None: passthrough (mod.rs:394-396)
            r.write_slot = (int)seen - 1; r.do_diff = 0;
        } else if (seen == 4) {                                    // pre-compute + first dispatch on slot 0 (mod.rs:177-214)
            r.write_slot = 3; r.compute_start = 1; r.grey_slot = 0;
        } else {                                                   // update_temporal_texture (bind_groups.rs:407-427)
            r.write_slot = r.grey_slot = (int)c->ring_index;
            c->ring_index = (c->ring_index + 1) % 4;
        }
        if (seen > 4) c->ring_seen = 5;                            // saturate
        return seen < 4;
    }
    r.n_slots = 2;
    r.median_is_max = c->cfg.flavor == DIPSB_FLAVOR_ALT_RING2_MEDIAN;
    r.write_slot = (int)c->ring_index;                            // texture_index, mod.rs:507-521
    c->ring_index = (c->ring_index + 1) % 2;
    r.snapshot = c->snapshot_pending ? 1 : 0;
    c->snapshot_pending = false;
    return false;                                                  // dips_alt always returns a computed frame
}

// The extra trailing frame of a per-frame shard on a layout the clip kernel cannot stream in place: wait for its arrival
// with a one-thread kernel, then difference it as an ordinary one-frame call (the state plane chains it to frame n-1).
static int32_t run_extra_frame(dipsb_ctx* c, const ShardExtra& x, uint64_t index) {
    if (x.flag) CK(c, launch_wait_flag(x.flag, x.epoch, x.timeout_ns, x.status, c->stream));
    return run_clip_on_stream(c, x.frame, 1, (c->g.npx * c->g.bpp + 15) & ~15ull, index, true, nullptr);
}

// Batch execution of the reference-exact flavours (DIPS_RING4 / ALT_RING2[_MEDIAN]) over a device-resident clip
// (accumulators and per-frame scalars; no visual output, no host copies).  The frames that change what the ring machine
// does -- the `dips` warm-up and start-plane frame, a frame that takes a snapshot -- go through the per-frame ring kernel,
// exactly as dipsb_push_frame runs them; the steady state after them is one launch of ring_clip_kernel per run of frames
// (ring in registers, every frame byte read once).  Spatial windows and layouts without aligned 8-pixel units stay per frame.
static int32_t grow_partials(dipsb_ctx* c, uint64_t need) {
    if (need <= c->partial_cap) return DIPSB_OK;
    CK(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->partials); c->partials = nullptr; c->partial_cap = 0;
    CK(c, cudaMalloc(&c->partials, need * sizeof(uint32_t)));
    c->partial_cap = need;
    return DIPSB_OK;
}
static int32_t run_ring_steady(dipsb_ctx* c, const uint8_t* d_frames, uint64_t m, uint64_t stride, uint64_t first, uint32_t seg_frames) {
    const Geometry& g = c->g;
    const int ns = c->cfg.flavor == DIPSB_FLAVOR_DIPS_RING4 ? 4 : 2;
    const uint32_t words = ring_clip_words_per_frame(g), pitch = (words + 3u) & ~3u;
    // scalar scratch rows of one launch: at most 64 M words, a whole number of ring turns
    const uint64_t max_rows = std::max<uint64_t>(8 * ns, ((64ull << 20) / pitch) / ns * ns);
    for (uint64_t done = 0; done < m;) {
        const uint64_t cnt = std::min(m - done, max_rows);
        int32_t rc = grow_partials(c, cnt * pitch);
        if (rc) return rc;
        RingClipArgs a;
        a.frames = d_frames + done * stride; a.stride = stride; a.n_frames = (uint32_t)cnt;
        a.ring = c->ring; a.n_slots = ns; a.first_slot = (int)c->ring_index;
        a.median_is_max = c->cfg.flavor == DIPSB_FLAVOR_ALT_RING2_MEDIAN;
        a.start = c->state[c->state_cur];
        a.acc_sum = c->acc; a.acc_cnt = c->acc + g.n_elems; a.partials = c->partials;
        a.tau = c->cfg.threshold; a.seg_frames = seg_frames;
        CK(c, launch_ring_clip(g, a, c->stream));
        CK(c, launch_finalize_scalars(g, c->partials, (uint32_t)cnt, words, c->d_sad + first + done, c->d_cnt + first + done, c->stream));
        c->ring_index = (uint32_t)((c->ring_index + cnt) % ns);
        done += cnt;
    }
    return DIPSB_OK;
}
static int32_t run_clip_ring(dipsb_ctx* c, const uint8_t* d_frames, uint64_t n, uint64_t stride, uint64_t first) {
    const Geometry& g = c->g;
    if (c->frames_processed + n > DIPSB_MAX_ACCUMULATED_FRAMES) return fail(c, DIPSB_ERR_STATE, "run_clip: the u32 sums would overflow; read the results and dipsb_reset");
    if (stride < g.npx * g.bpp) return fail(c, DIPSB_ERR_INVALID, "run_clip: stride %llu smaller than a frame", (unsigned long long)stride);
    int32_t rc = ensure_scalars(c, first + n);
    if (rc) return rc;
    if ((rc = ensure_acc_zero(c))) return rc;
    // measurement / test hooks: DIPSB_RING_BATCH=0 keeps every frame on the per-frame kernel; DIPSB_RING_SEG_FRAMES forces
    // the frame-segment length of ring_clip_kernel (small clips would otherwise run as one segment)
    const char* env_batch = getenv("DIPSB_RING_BATCH");
    const char* env_seg = getenv("DIPSB_RING_SEG_FRAMES");
    const bool batchable = !windowed(c) && !(env_batch && env_batch[0] == '0') && ring_clip_available(g, d_frames, stride);
    const uint64_t min_run = c->cfg.flavor == DIPSB_FLAVOR_DIPS_RING4 ? 8 : 4;
    uint64_t k = 0;
    for (; k < n; ++k) {
        const bool steady = !c->snapshot_pending && (c->cfg.flavor != DIPSB_FLAVOR_DIPS_RING4 || c->ring_seen >= 4);
        if (batchable && steady && n - k >= min_run) break;
        RingArgs r;
        r.frame = d_frames + k * stride; r.pitch = (uint64_t)g.width * g.bpp; r.format = g.format; r.chan_byte = g.chan_byte;
        if (windowed(c)) {
            if ((rc = ensure_i2_scratch(c))) return rc;
            if ((rc = filtered_plane(c, r.frame, g.format, c->i2_scratch + g.npx))) return rc;
            r.i2src = c->i2_scratch + g.npx;
        }
        r.ring = c->ring; r.start = c->state[c->state_cur];
        r.acc_sum = c->acc; r.acc_cnt = c->acc + g.n_elems; r.sad = c->d_sad + first + k; r.cnt = c->d_cnt + first + k;
        r.out_rgba = nullptr;
        r.tau = c->cfg.threshold; r.colorize = c->cfg.colorize; r.filter = c->cfg.filter; r.sig_scalar = c->cfg.sigmoid_scalar;
        ring_next_frame(c, r);
        CK(c, cudaMemsetAsync(r.sad, 0, sizeof(uint64_t), c->stream));
        CK(c, cudaMemsetAsync(r.cnt, 0, sizeof(uint64_t), c->stream));
        CK(c, launch_ring(g, r, c->stream));
    }
    if (k < n) {
        if (c->cfg.flavor == DIPSB_FLAVOR_DIPS_RING4) c->ring_seen = 5;
        if ((rc = run_ring_steady(c, d_frames + k * stride, n - k, stride, first + k, env_seg ? (uint32_t)strtoul(env_seg, nullptr, 10) : 0u))) return rc;
    }
    c->state_valid = true;
    c->frames_processed += n;
    c->scal_hi = std::max(c->scal_hi, first + n);
    c->stream_index = first + n;
    c->last_plan[1] = 0; c->last_plan[7] = k < n ? 2 : 0;
    return DIPSB_OK;
}

// `zero_padded`: the rows come from the library's own re-packed buffers -- pitch a multiple of 16 and zero bytes between the
// end of a frame and its pitch -- so the clip kernel may round its last bulk copy of a frame up to 16 bytes.
int32_t dipsb::run_clip_on_stream(dipsb_ctx* c, const uint8_t* d_frames, uint64_t n, uint64_t stride, uint64_t first,
                                  bool zero_padded, const ShardExtra* extra) {
    const Geometry& g = c->g;
    if (n == 0) return DIPSB_OK;
    const ShardPush* push = (extra && extra->push.nranks) ? &extra->push : nullptr;
    bool* pushed = extra ? extra->pushed : nullptr;
    const bool defer_fin = extra && extra->defer_finalize;
    if (extra && (!extra->frame || c->cfg.mode != DIPSB_MODE_PERFRAME)) extra = nullptr;   // no trailing frame
    const uint64_t n_scal = n + (extra ? 1 : 0);   // scalar rows this call produces
    if (c->cfg.flavor != DIPSB_FLAVOR_FRAME0) return run_clip_ring(c, d_frames, n, stride, first);
    if (n > 0x7FFFFFFFull) return fail(c, DIPSB_ERR_INVALID, "run_clip: too many frames in one call");
    if (c->frames_processed + n > DIPSB_MAX_ACCUMULATED_FRAMES)
        return fail(c, DIPSB_ERR_STATE, "run_clip: %llu frames accumulated, %llu more would overflow the u32 sums; read the results and dipsb_reset",
                    (unsigned long long)c->frames_processed, (unsigned long long)n);
    if (stride < g.npx * g.bpp) return fail(c, DIPSB_ERR_INVALID, "run_clip: stride %llu smaller than a frame", (unsigned long long)stride);
    int32_t rc = ensure_scalars(c, first + n_scal);
    if (rc) return rc;
    // the TMA bulk copies of the clip kernel need 16-byte aligned addresses and sizes
    // (a spatial window > 1 needs a filtered intensity plane per frame: per-frame kernels, not the clip kernel)
    const uint64_t fb = g.npx * g.bpp;
    const bool aligned = !windowed(c) && (((uintptr_t)d_frames | stride) & 15u) == 0 && ((fb & 15u) == 0 || (zero_padded && stride >= ((fb + 15) & ~15ull)));
    if (!aligned && !windowed(c) && !zero_padded && n >= 2) {
        // Unaligned base, pitch or frame size: re-pack through an aligned, zero-padded scratch (one extra read + write of
        // the clip on the device; repack_kernel -- a 2-D cudaMemcpy with an odd pitch runs at a third of its rate) and
        // stream that, instead of one small kernel per frame.
        const uint64_t dpitch = (fb + 15) & ~15ull;
        const uint64_t per_chunk = std::max<uint64_t>(1, std::min<uint64_t>(n, (256ull << 20) / dpitch));
        if (c->repack_bytes < per_chunk * dpitch) {
            CK(c, cudaStreamSynchronize(c->stream));
            cudaFree(c->d_repack); c->d_repack = nullptr; c->repack_bytes = 0;
            CK(c, cudaMalloc(&c->d_repack, per_chunk * dpitch));
            c->repack_bytes = per_chunk * dpitch;
            CK(c, cudaMemsetAsync(c->d_repack, 0, c->repack_bytes, c->stream));
        }
        for (uint64_t done = 0; done < n;) {
            const uint64_t m = std::min(per_chunk, n - done);
            CK(c, launch_repack(g, d_frames + done * stride, stride, fb, m, c->d_repack, dpitch, c->stream));
            rc = run_clip_on_stream(c, c->d_repack, m, dpitch, first + done, true, nullptr);
            if (rc) return rc;
            done += m;
        }
        return extra ? run_extra_frame(c, *extra, first + n) : DIPSB_OK;
    }
    if (!c->state_valid) {   // frame 0 of the call is the reference (overall) / has no predecessor (per-frame): D = 0
        if (windowed(c)) {
            rc = filtered_plane(c, d_frames, g.format, c->state[c->state_cur]);
            if (rc) return rc;
        } else {
            CK(c, launch_prime(g, d_frames, c->state[c->state_cur], c->stream));
        }
        c->state_valid = true;
    }
    const uint32_t tau = c->cfg.threshold;
    if (aligned) {
        const uint32_t segs = plan_segments(c, n);
        const uint32_t words = g.n_tiles * clip_active_warps(g);
        const uint64_t need = n_scal * (uint64_t)((words + 3u) & ~3u);   // rows pitched to 4 words (16-byte aligned)
        if (need >= (1ull << 32)) return fail(c, DIPSB_ERR_INVALID, "run_clip: %llu frames x %u warps exceed the per-call scalar scratch; split the call", (unsigned long long)n, words);
        if (need > c->partial_cap) {
            CK(c, cudaStreamSynchronize(c->stream));
            cudaFree(c->partials); c->partials = nullptr; c->partial_cap = 0;
            CK(c, cudaMalloc(&c->partials, need * sizeof(uint32_t)));
            c->partial_cap = need;
        }
        ClipArgs a;
        a.frames = d_frames; a.stride = stride; a.n_frames = (uint32_t)n; a.n_segments = segs;
        a.state_in = c->state[c->state_cur]; a.state_out = c->state[c->state_cur ^ 1];
        a.acc_sum = c->acc; a.acc_cnt = c->acc + g.n_elems; a.partials = c->partials;
        a.tau = tau; a.mode = c->cfg.mode;
        if (extra) {
            a.extra_frame = extra->frame; a.halo_flag = extra->flag; a.halo_epoch = extra->epoch;
            a.wait_timeout_ns = extra->timeout_ns; a.status = extra->status;
        }
        if (push && clip_can_push(g, segs)) {
            a.push = push;
            if (pushed) *pushed = true;
        }
        if (c->acc_zero_pending) {
            if (clip_can_store_first(g, segs)) { a.first_store = true; c->acc_zero_pending = false; }
            else if ((rc = ensure_acc_zero(c))) return rc;
        }
        if (c->timing) {
            if (c->tev_used + 2 > c->tev.size()) {
                cudaEvent_t e0, e1;
                CK(c, cudaEventCreate(&e0));
                CK(c, cudaEventCreate(&e1));
                c->tev.push_back(e0); c->tev.push_back(e1);
            }
            CK(c, cudaEventRecord(c->tev[c->tev_used], c->stream));
        }
        CK(c, launch_clip(g, a, c->stream));
        if (c->timing) {
            CK(c, cudaEventRecord(c->tev[c->tev_used + 1], c->stream));
            c->tev_used += 2;
        }
        if (defer_fin) { c->fin.pending = true; c->fin.n = (uint32_t)n_scal; c->fin.words = words; c->fin.first = first; }
        else CK(c, launch_finalize_scalars(g, c->partials, (uint32_t)n_scal, words, c->d_sad + first, c->d_cnt + first, c->stream));
        if (c->cfg.mode == DIPSB_MODE_PERFRAME) c->state_cur ^= 1;
        c->last_plan[1] = segs;
        c->last_plan[7] = 1;
    } else {   // spatial window, or a single unaligned frame: per-frame kernels
        if ((rc = ensure_acc_zero(c))) return rc;
        CK(c, cudaMemsetAsync(c->d_sad + first, 0, n * sizeof(uint64_t), c->stream));
        CK(c, cudaMemsetAsync(c->d_cnt + first, 0, n * sizeof(uint64_t), c->stream));
        for (uint64_t k = 0; k < n; ++k) {
            FrameArgs f;
            f.frame = d_frames + k * stride; f.pitch = (uint64_t)g.width * g.bpp; f.format = g.format; f.chan_byte = g.chan_byte;
            if (windowed(c)) {
                rc = ensure_i2_scratch(c);
                if (rc) return rc;
                rc = filtered_plane(c, f.frame, g.format, c->i2_scratch + g.npx);
                if (rc) return rc;
                f.i2src = c->i2_scratch + g.npx;
            }
            f.state_in = c->state[c->state_cur];
            f.state_out = c->cfg.mode == DIPSB_MODE_PERFRAME ? c->state[c->state_cur] : nullptr;
            f.acc_sum = c->acc; f.acc_cnt = c->acc + g.n_elems;
            f.sad = c->d_sad + first + k; f.cnt = c->d_cnt + first + k; f.out_rgba = nullptr;
            f.tau = tau; f.accumulate = 1; f.colorize = 0; f.filter = DIPSB_FILTER_NONE; f.sig_scalar = 5.0f;
            CK(c, launch_frame(g, f, c->stream));
        }
        c->last_plan[1] = 0;
        c->last_plan[7] = 0;
        c->frames_processed += n;
        c->scal_hi = std::max(c->scal_hi, first + n);
        c->stream_index = first + n;
        return extra ? run_extra_frame(c, *extra, first + n) : DIPSB_OK;
    }
    c->frames_processed += n_scal;
    c->scal_hi = std::max(c->scal_hi, first + n_scal);
    c->stream_index = first + n_scal;
    return DIPSB_OK;
}

extern "C" int32_t dipsb_run_clip_device(dipsb_ctx* c, const void* d_frames, uint64_t n_frames, uint64_t stride, uint64_t first) {
    if (!c || (!d_frames && n_frames)) return DIPSB_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    return run_clip_on_stream(c, (const uint8_t*)d_frames, n_frames, stride, first, false, nullptr);
}

extern "C" int32_t dipsb_run_clip_host(dipsb_ctx* c, const uint8_t* frames, uint64_t n, uint64_t stride, uint64_t first) {
    if (!c || (!frames && n)) return DIPSB_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    return run_clip_host_impl(c, frames, n, stride, first, nullptr);
}

int32_t dipsb::run_clip_host_impl(dipsb_ctx* c, const uint8_t* frames, uint64_t n, uint64_t stride, uint64_t first,
                                  const HostClipHooks* hooks) {
    const Geometry& g = c->g;
    const uint64_t fb = g.npx * g.bpp;
    if (n == 0) return DIPSB_OK;
    if (stride < fb) return fail(c, DIPSB_ERR_INVALID, "run_clip_host: stride smaller than a frame");
    // frames are re-packed on the device side at a 16-byte aligned pitch so that the TMA path is always taken
    const uint64_t dpitch = (fb + 15) & ~15ull;
    const uint64_t per_chunk = std::max<uint64_t>(1, std::min<uint64_t>(n, (256ull << 20) / dpitch));
    const size_t cbytes = per_chunk * dpitch;
    cudaPointerAttributes attr;
    bool pinned = false;
    if (cudaPointerGetAttributes(&attr, frames) == cudaSuccess) pinned = (attr.type == cudaMemoryTypeHost);
    else cudaGetLastError();
    if (c->chunk_bytes < cbytes) {
        CK(c, cudaStreamSynchronize(c->stream));
        CK(c, cudaStreamSynchronize(c->copy_stream));
        for (int k = 0; k < 2; ++k) {
            if (c->h_chunk[k]) cudaFreeHost(c->h_chunk[k]);
            if (c->d_chunk[k]) cudaFree(c->d_chunk[k]);
            c->h_chunk[k] = nullptr; c->d_chunk[k] = nullptr;
            c->chunk_used[k] = false;
        }
        c->chunk_bytes = 0;
        for (int k = 0; k < 2; ++k) {
            CK(c, cudaMalloc(&c->d_chunk[k], cbytes));
            CK(c, cudaMallocHost(&c->h_chunk[k], cbytes));
            CK(c, cudaMemsetAsync(c->d_chunk[k], 0, cbytes, c->copy_stream));   // row padding stays zero from here on
            memset(c->h_chunk[k], 0, cbytes);
        }
        c->chunk_bytes = cbytes;
    }
    // The slot flags live in the context: a second call without a synchronisation in between (reset + run loops, a long
    // video fed in several calls) must still wait for the kernels and copies of the previous call that use the same slot.
    uint64_t done = 0;
    int slot = c->chunk_slot;
    int last_slot = -1;
    while (done < n) {
        const uint64_t m = std::min(per_chunk, n - done);
        const uint8_t* src = frames + done * stride;
        if (c->chunk_used[slot]) CK(c, cudaStreamWaitEvent(c->copy_stream, c->ev_done[slot], 0));   // kernels finished with d_chunk[slot]
        if (pinned) {
            CK(c, cudaMemcpy2DAsync(c->d_chunk[slot], dpitch, src, stride, fb, m, cudaMemcpyHostToDevice, c->copy_stream));
        } else {
            if (c->chunk_used[slot]) CK(c, cudaEventSynchronize(c->ev_copy[slot]));        // previous H2D from h_chunk[slot] done
            host_copy2d(c->h_chunk[slot], dpitch, src, stride, fb, m);
            CK(c, cudaMemcpyAsync(c->d_chunk[slot], c->h_chunk[slot], m * dpitch, cudaMemcpyHostToDevice, c->copy_stream));
        }
        CK(c, cudaEventRecord(c->ev_copy[slot], c->copy_stream));
        CK(c, cudaStreamWaitEvent(c->stream, c->ev_copy[slot], 0));
        if (done == 0 && hooks && hooks->after_first_upload) {
            int32_t hrc = hooks->after_first_upload(c, c->d_chunk[slot], hooks->user);
            if (hrc) return hrc;
        }
        const ShardExtra* extra = (hooks && done + m == n) ? hooks->extra : nullptr;
        int32_t rc = run_clip_on_stream(c, c->d_chunk[slot], m, dpitch, first + done, true, extra);
        if (rc) return rc;
        CK(c, cudaEventRecord(c->ev_done[slot], c->stream));
        c->chunk_used[slot] = true;
        last_slot = slot;
        slot ^= 1;
        done += m;
    }
    c->chunk_slot = slot;
    // the caller's frames are borrowed for the call only: a page-locked clip is read by the copy engine directly, so wait
    // for the last upload (the earlier ones precede it on the copy stream); the kernels of the last chunk stay asynchronous
    if (pinned && last_slot >= 0) CK(c, cudaEventSynchronize(c->ev_copy[last_slot]));
    return DIPSB_OK;
}

// ---- streaming -----------------------------------------------------------------------------------------------------
extern "C" int32_t dipsb_snapshot(dipsb_ctx* c) {
    if (!c) return DIPSB_ERR_INVALID;
    c->snapshot_pending = true;
    return DIPSB_OK;
}

static int32_t ensure_slot(dipsb_ctx* c, dipsb_ctx::FrameSlot& sl, size_t in_bytes) {
    const size_t rgba = c->g.npx * 4;
    if (sl.in_bytes < in_bytes) {
        if (sl.h_in) cudaFreeHost(sl.h_in);
        cudaFree(sl.d_in);
        sl.h_in = nullptr; sl.d_in = nullptr; sl.in_bytes = 0;
        CK(c, cudaMallocHost(&sl.h_in, in_bytes));
        CK(c, cudaMalloc(&sl.d_in, in_bytes));
        sl.in_bytes = in_bytes;
    }
    if (!sl.d_out) CK(c, cudaMalloc(&sl.d_out, rgba));
    if (!sl.h_out) CK(c, cudaMallocHost(&sl.h_out, rgba));
    if (!sl.h_stat) CK(c, cudaMallocHost(&sl.h_stat, 2 * sizeof(uint64_t)));
    if (!sl.ev_done) CK(c, cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming));
    for (auto& e : sl.ev_in) if (!e) CK(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : sl.ev_k) if (!e) CK(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : sl.ev_out) if (!e) CK(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return DIPSB_OK;
}

// frames of 2 MB and more are worked on in up to 4 row bands so that copies, transfers and kernels overlap
static uint32_t stage_pieces(uint64_t bytes, uint32_t rows) {
    static const uint64_t max_bands = [] {
        const char* e = getenv("DIPSB_FRAME_BANDS");           // 1 .. 8, default 4 (measured: profiles/r01_sweeps.md)
        const long v = e ? strtol(e, nullptr, 10) : 0;
        return (uint64_t)(v >= 1 && v <= 8 ? v : 4);
    }();
    const uint64_t p = std::min<uint64_t>(max_bands, bytes >> 20);
    return (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(p, rows));
}

// page-locked host memory (dipsb_host_alloc, cudaHostAlloc/cudaHostRegister, torch pin_memory): the copy engine can
// reach it directly, so the staging memcpy -- the dominant cost of a 1080p per-frame call -- is skipped
static bool host_pinned(const void* p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return attr.type == cudaMemoryTypeHost;
}

static int32_t start_readback(dipsb_ctx* c, dipsb_ctx::FrameSlot& sl, uint8_t* out_rgba, bool want_stats);

// Stage one frame into `sl` and enqueue upload, kernel and read-back.  `overlap`: upload on the copy stream so that it runs
// concurrently with the previous frame's kernel / read-back (pipelined mode); otherwise everything on the context's stream.
// `out_direct`: page-locked caller buffer the difference frame is read back into (synchronous mode), `defer_out`: leave
// the read-back to collect_frame.
static int32_t submit_frame(dipsb_ctx* c, dipsb_ctx::FrameSlot& sl, const uint8_t* px, uint32_t width, uint32_t height,
                            uint32_t stride, int32_t format, bool want_rgba, bool overlap, uint8_t* out_direct = nullptr,
                            bool defer_out = false, dipsb_ctx::FrameSlot* prev = nullptr, uint8_t* prev_out = nullptr,
                            bool prev_stats = false) {
    const Geometry& g = c->g;
    if (width != g.width || height != g.height) return fail(c, DIPSB_ERR_INVALID, "push_frame: %ux%u does not match the context's %ux%u", width, height, g.width, g.height);
    if (format < 0 || format > 3) return fail(c, DIPSB_ERR_INVALID, "push_frame: bad format %d", format);
    const int bpp = bpp_of(format);
    const uint64_t row = (uint64_t)width * bpp;
    if (stride < row) return fail(c, DIPSB_ERR_INVALID, "push_frame: stride %u smaller than a row (%llu)", stride, (unsigned long long)row);
    const size_t fb = row * height;
    if (c->frames_processed + 1 > DIPSB_MAX_ACCUMULATED_FRAMES)
        return fail(c, DIPSB_ERR_STATE, "push_frame: %llu frames accumulated, one more would overflow the u32 sums; read the results and dipsb_reset",
                    (unsigned long long)c->frames_processed);
    int32_t rc = ensure_slot(c, sl, fb);
    if (rc) return rc;
    // Pipelined call with a page-locked frame: the two big transfers of the call go out before anything else -- this frame's
    // upload on the copy stream, then the previous frame's read-back on the context's stream (ahead of this frame's kernel) --
    // so that the ~40 us of bookkeeping below run under them instead of in front of them.
    // (One transfer, not row bands: four smaller DMAs plus per-band launches measured 3.7 k calls/s against 4.35 k.)
    uint32_t uploaded = 0;     // bands already on their way
    if (overlap && px && stride == row && host_pinned(px)) {
        uploaded = 1u;
        for (uint32_t k = 0; k < uploaded; ++k) {
            const uint32_t r0 = (uint32_t)((uint64_t)height * k / uploaded), r1 = (uint32_t)((uint64_t)height * (k + 1) / uploaded);
            CK(c, cudaMemcpyAsync(sl.d_in + (uint64_t)r0 * row, px + (uint64_t)r0 * row, (uint64_t)(r1 - r0) * row, cudaMemcpyHostToDevice, c->copy_stream));
            CK(c, cudaEventRecord(sl.ev_in[k], c->copy_stream));
        }
    }
    if (prev) {
        rc = start_readback(c, *prev, prev_out, prev_stats);
        if (rc) return rc;
    }
    rc = ensure_scalars(c, c->stream_index + 1);
    if (rc) return rc;
    if ((rc = ensure_acc_zero(c))) return rc;
    const uint16_t* i2src = nullptr;
    if (windowed(c)) {   // N4: spatially filtered intensity of this frame (dips_shader.wgsl:187), computed below
        rc = ensure_i2_scratch(c);
        if (rc) return rc;
        i2src = c->i2_scratch + g.npx;
    }
    // The input slice is borrowed for the call only (frame_extractor.rs:224-226): it is either staged now or, when it is
    // page-locked, uploaded straight from the caller's buffer and the upload awaited before this call returns.
    // px == nullptr: the frame was staged and its upload started by dipsb_stage_frame (slot 0); only kernels and read-back follow
    const bool pre_staged = px == nullptr;
    const bool in_direct = uploaded != 0 || (!pre_staged && stride == row && host_pinned(px));
    // The synchronous call works in row bands: upload of band k+1 (copy stream), kernels of band k (the context's stream)
    // and read-back of band k-1 (read-back stream) run concurrently, and so do the CPU staging copies on either side.  In
    // the pipelined call the neighbouring frames already overlap and the extra launches only cost; a spatial window
    // needs the whole frame.
    const uint32_t bands = pre_staged ? c->staged_bands : uploaded ? uploaded
                                      : (overlap || windowed(c)) ? 1u : stage_pieces(std::max<uint64_t>(fb, g.npx * 4), height);
    const bool banded = bands > 1;
    const bool side_upload = overlap || banded;
    cudaStream_t up = side_upload ? c->copy_stream : c->stream;
    // (a pipelined frame whose read-back is deferred to the next call keeps everything of this call on the context's stream)
    cudaStream_t tail = (banded && !(overlap && defer_out)) ? c->out_stream : c->stream;
    const uint64_t idx = c->stream_index;
    CK(c, cudaMemsetAsync(c->d_sad + idx, 0, sizeof(uint64_t), c->stream));
    CK(c, cudaMemsetAsync(c->d_cnt + idx, 0, sizeof(uint64_t), c->stream));

    // ---- what this frame does to the state machine (once per frame; the launches below follow per band) ----
    bool establishes;
    const bool ring_flavour = c->cfg.flavor != DIPSB_FLAVOR_FRAME0;
    FrameArgs f;
    RingArgs r;
    if (!ring_flavour) {
        establishes = !c->state_valid || c->snapshot_pending;
        f.i2src = i2src;
        f.frame = sl.d_in; f.pitch = row; f.format = format; f.chan_byte = chan_byte_of(format, c->cfg.chroma);
        f.acc_sum = c->acc; f.acc_cnt = c->acc + g.n_elems;
        f.sad = c->d_sad + idx; f.cnt = c->d_cnt + idx;
        f.tau = c->cfg.threshold; f.colorize = c->cfg.colorize; f.filter = c->cfg.filter; f.sig_scalar = c->cfg.sigmoid_scalar;
        f.state_in = c->state[c->state_cur];
        if (establishes) {
            // this frame becomes the reference: D = 0 for it, output is the input passed through (dips/src/lib.rs:241-245)
            f.state_out = c->state[c->state_cur]; f.out_rgba = nullptr; f.accumulate = 0;
            c->snapshot_pending = false;
        } else {
            f.state_out = c->cfg.mode == DIPSB_MODE_PERFRAME ? c->state[c->state_cur] : nullptr;
            f.out_rgba = want_rgba ? sl.d_out : nullptr; f.accumulate = 1;
        }
    } else {
        r.i2src = i2src;
        r.frame = sl.d_in; r.pitch = row; r.format = format; r.chan_byte = chan_byte_of(format, c->cfg.chroma);
        r.ring = c->ring; r.start = c->state[c->state_cur];
        r.acc_sum = c->acc; r.acc_cnt = c->acc + g.n_elems; r.sad = c->d_sad + idx; r.cnt = c->d_cnt + idx;
        r.out_rgba = want_rgba ? sl.d_out : nullptr;
        r.tau = c->cfg.threshold; r.colorize = c->cfg.colorize; r.filter = c->cfg.filter; r.sig_scalar = c->cfg.sigmoid_scalar;
        establishes = ring_next_frame(c, r);
    }
    c->state_valid = true;
    sl.out_direct = want_rgba && out_direct != nullptr;
    sl.out_deferred = want_rgba && !sl.out_direct && defer_out;
    sl.out_pieces = 0;
    const bool readback = want_rgba && !sl.out_deferred;

    for (uint32_t k = 0; k < bands; ++k) {
        const uint32_t r0 = (uint32_t)((uint64_t)height * k / bands), r1 = (uint32_t)((uint64_t)height * (k + 1) / bands);
        const uint64_t p0 = (uint64_t)r0 * width, p1 = (uint64_t)r1 * width;
        // upload
        if (pre_staged) {
            CK(c, cudaStreamWaitEvent(c->stream, sl.ev_in[k], 0));   // recorded by dipsb_stage_frame behind the upload of band k
        } else if (uploaded) {
            CK(c, cudaStreamWaitEvent(c->stream, sl.ev_in[k], 0));
        } else {
            const uint8_t* src = px + (uint64_t)r0 * stride;
            if (!in_direct) {
                host_copy2d(sl.h_in + (uint64_t)r0 * row, row, src, stride, row, r1 - r0);
                src = sl.h_in + (uint64_t)r0 * row;
            }
            CK(c, cudaMemcpyAsync(sl.d_in + (uint64_t)r0 * row, src, (uint64_t)(r1 - r0) * row, cudaMemcpyHostToDevice, up));
            if (side_upload) {
                CK(c, cudaEventRecord(sl.ev_in[k], c->copy_stream));
                CK(c, cudaStreamWaitEvent(c->stream, sl.ev_in[k], 0));
            }
        }
        // kernels
        if (windowed(c)) {
            rc = filtered_plane(c, sl.d_in, format, c->i2_scratch + g.npx);
            if (rc) return rc;
        }
        if (!ring_flavour) {
            f.p_begin = p0; f.p_end = p1;
            CK(c, launch_frame(g, f, c->stream));
        } else {
            r.p_begin = p0; r.p_end = p1;
            CK(c, launch_ring(g, r, c->stream));
        }
        if (establishes && want_rgba) CK(c, launch_passthrough_rgba(g, sl.d_in, row, format, sl.d_out, c->stream, p0, p1));
        // read-back
        if (banded && tail != c->stream) {
            CK(c, cudaEventRecord(sl.ev_k[k], c->stream));
            CK(c, cudaStreamWaitEvent(tail, sl.ev_k[k], 0));
        }
        sl.out_off[k] = p0 * 4;
        if (readback) {
            uint8_t* dst = sl.out_direct ? out_direct : sl.h_out;
            CK(c, cudaMemcpyAsync(dst + p0 * 4, sl.d_out + p0 * 4, (p1 - p0) * 4, cudaMemcpyDeviceToHost, tail));
            if (!sl.out_direct) CK(c, cudaEventRecord(sl.ev_out[k], tail));   // collect_frame copies band k out behind it
        }
    }
    sl.out_off[bands] = g.npx * 4;
    if (readback && !sl.out_direct) sl.out_pieces = (int)bands;
    if (sl.out_deferred) {
        // kernels of this frame done: its read-back (and, if asked for, its two scalars) may start -- start_readback, next call
        CK(c, cudaEventRecord(sl.ev_k[0], c->stream));
    } else {
        CK(c, cudaMemcpyAsync(&sl.h_stat[0], c->d_sad + idx, sizeof(uint64_t), cudaMemcpyDeviceToHost, tail));
        CK(c, cudaMemcpyAsync(&sl.h_stat[1], c->d_cnt + idx, sizeof(uint64_t), cudaMemcpyDeviceToHost, tail));
        CK(c, cudaEventRecord(sl.ev_done, tail));
    }
    sl.pending = true; sl.want_rgba = want_rgba; sl.idx = idx; sl.status = establishes ? DIPSB_NOT_READY : DIPSB_OK;
    c->stream_index = idx + 1;
    c->frames_processed += 1;
    c->scal_hi = std::max(c->scal_hi, idx + 1);
    if (in_direct && overlap) CK(c, cudaEventSynchronize(sl.ev_in[bands - 1]));   // the caller's buffer is free again on return
    return DIPSB_OK;
}

// deferred read-back (pipelined mode, page-locked caller): enqueue it now, behind whatever the stream already holds
static int32_t start_readback(dipsb_ctx* c, dipsb_ctx::FrameSlot& sl, uint8_t* out_rgba, bool want_stats) {
    if (!sl.pending || !sl.out_deferred) return DIPSB_OK;
    if (!out_rgba) {   // nobody wants the frame: only (perhaps) its scalars
        CK(c, cudaStreamWaitEvent(c->out_stream, sl.ev_k[0], 0));
        if (want_stats) {
            CK(c, cudaMemcpyAsync(&sl.h_stat[0], c->d_sad + sl.idx, sizeof(uint64_t), cudaMemcpyDeviceToHost, c->out_stream));
            CK(c, cudaMemcpyAsync(&sl.h_stat[1], c->d_cnt + sl.idx, sizeof(uint64_t), cudaMemcpyDeviceToHost, c->out_stream));
        }
        CK(c, cudaEventRecord(sl.ev_done, c->out_stream));
        sl.out_deferred = false; sl.want_rgba = false;
        return DIPSB_OK;
    }
    const bool direct = host_pinned(out_rgba);
    // on the read-back stream, behind the frame's own kernels only: on the context's stream the 8 MB copy would sit between
    // the kernels of consecutive frames and make (read-back + kernels + scalar copies) the per-frame critical chain
    CK(c, cudaStreamWaitEvent(c->out_stream, sl.ev_k[0], 0));
    CK(c, cudaMemcpyAsync(direct ? out_rgba : sl.h_out, sl.d_out, c->g.npx * 4, cudaMemcpyDeviceToHost, c->out_stream));
    if (want_stats) {   // the frame's two scalars ride behind it, and only when the caller asked for them
        CK(c, cudaMemcpyAsync(&sl.h_stat[0], c->d_sad + sl.idx, sizeof(uint64_t), cudaMemcpyDeviceToHost, c->out_stream));
        CK(c, cudaMemcpyAsync(&sl.h_stat[1], c->d_cnt + sl.idx, sizeof(uint64_t), cudaMemcpyDeviceToHost, c->out_stream));
    }
    CK(c, cudaEventRecord(sl.ev_done, c->out_stream));
    sl.out_deferred = false;
    sl.out_direct = direct;
    return DIPSB_OK;
}

// wait for the frame in `sl` and hand its output to the caller; returns the frame's own status (OK / NOT_READY)
static int32_t collect_frame(dipsb_ctx* c, dipsb_ctx::FrameSlot& sl, uint8_t* out_rgba, dipsb_frame_stats* stats) {
    if (!sl.pending) return fail(c, DIPSB_ERR_STATE, "no frame in flight");
    int32_t rb = start_readback(c, sl, out_rgba, stats != nullptr);
    if (rb) return rb;
    const bool copy_out = out_rgba && sl.want_rgba && !sl.out_direct && !sl.out_deferred;
    if (copy_out && sl.out_pieces > 1) {
        for (int k = 0; k < sl.out_pieces; ++k) {
            const uint64_t b0 = sl.out_off[k], b1 = sl.out_off[k + 1];
            CK(c, cudaEventSynchronize(sl.ev_out[k]));
            host_copy2d(out_rgba + b0, b1 - b0, sl.h_out + b0, b1 - b0, b1 - b0, 1);
        }
        CK(c, cudaEventSynchronize(sl.ev_done));
    } else {
        CK(c, cudaEventSynchronize(sl.ev_done));
        if (copy_out) host_copy2d(out_rgba, c->g.npx * 4, sl.h_out, c->g.npx * 4, c->g.npx * 4, 1);
    }
    if (stats) { stats->frame_index = sl.idx; stats->sad = sl.h_stat[0]; stats->count = sl.h_stat[1]; }
    sl.pending = false;
    return sl.status;
}

extern "C" int32_t dipsb_push_frame(dipsb_ctx* c, const uint8_t* px, uint32_t width, uint32_t height, uint32_t stride,
                                    int32_t format, uint8_t* out_rgba, dipsb_frame_stats* stats) {
    if (!c || !px) return DIPSB_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    if (c->slot[0].pending || c->slot[1].pending) return fail(c, DIPSB_ERR_STATE, "push_frame: a pipelined frame is in flight; call dipsb_flush_frame first");
    if (c->staged) { CK(c, cudaEventSynchronize(c->slot[0].ev_in[c->staged_bands - 1])); c->staged = false; }   // a staged frame is dropped
    int32_t rc = submit_frame(c, c->slot[0], px, width, height, stride, format, out_rgba != nullptr, false,
                              out_rgba && host_pinned(out_rgba) ? out_rgba : nullptr);
    if (rc) return rc;
    return collect_frame(c, c->slot[0], out_rgba, stats);
}

// The reference splits its frame call in two -- ComputeState::add_texture (dips/src/gpu/mod.rs:170: keeps the borrowed frame)
// and dispatch (:306: computes and returns it).  Staging copies the frame straight into the library's page-locked input slot
// (threaded host copy) and starts its upload, so a wrapper need not keep a copy of its own for dispatch.
extern "C" int32_t dipsb_stage_frame(dipsb_ctx* c, const uint8_t* px, uint32_t width, uint32_t height, uint32_t stride, int32_t format) {
    if (!c || !px) return DIPSB_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    const Geometry& g = c->g;
    if (c->slot[0].pending || c->slot[1].pending) return fail(c, DIPSB_ERR_STATE, "stage_frame: a pipelined frame is in flight; call dipsb_flush_frame first");
    if (width != g.width || height != g.height) return fail(c, DIPSB_ERR_INVALID, "stage_frame: %ux%u does not match the context's %ux%u", width, height, g.width, g.height);
    if (format < 0 || format > 3) return fail(c, DIPSB_ERR_INVALID, "stage_frame: bad format %d", format);
    const uint64_t row = (uint64_t)width * bpp_of(format);
    if (stride < row) return fail(c, DIPSB_ERR_INVALID, "stage_frame: stride %u smaller than a row (%llu)", stride, (unsigned long long)row);
    dipsb_ctx::FrameSlot& sl = c->slot[0];
    int32_t rc = ensure_slot(c, sl, row * height);
    if (rc) return rc;
    if (c->staged) CK(c, cudaEventSynchronize(sl.ev_in[c->staged_bands - 1]));   // a staged frame that was never dispatched: its upload still reads h_in
    // in the same row bands as dipsb_push_frame: the upload of band k runs while the CPU copies band k+1
    const uint32_t bands = windowed(c) ? 1u : stage_pieces(std::max<uint64_t>(row * height, g.npx * 4), height);
    for (uint32_t k = 0; k < bands; ++k) {
        const uint32_t r0 = (uint32_t)((uint64_t)height * k / bands), r1 = (uint32_t)((uint64_t)height * (k + 1) / bands);
        host_copy2d(sl.h_in + (uint64_t)r0 * row, row, px + (uint64_t)r0 * stride, stride, row, r1 - r0);
        CK(c, cudaMemcpyAsync(sl.d_in + (uint64_t)r0 * row, sl.h_in + (uint64_t)r0 * row, (uint64_t)(r1 - r0) * row, cudaMemcpyHostToDevice, c->copy_stream));
        CK(c, cudaEventRecord(sl.ev_in[k], c->copy_stream));
    }
    c->staged = true; c->staged_format = format; c->staged_bands = bands;
    return DIPSB_OK;
}

extern "C" int32_t dipsb_dispatch_staged(dipsb_ctx* c, uint8_t* out_rgba, dipsb_frame_stats* stats) {
    if (!c) return DIPSB_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    if (!c->staged) return fail(c, DIPSB_ERR_STATE, "dispatch_staged: no frame staged (dipsb_stage_frame)");
    c->staged = false;
    const int bpp = bpp_of(c->staged_format);
    int32_t rc = submit_frame(c, c->slot[0], nullptr, c->g.width, c->g.height, c->g.width * (uint32_t)bpp, c->staged_format, out_rgba != nullptr, false,
                              out_rgba && host_pinned(out_rgba) ? out_rgba : nullptr);
    if (rc) return rc;
    return collect_frame(c, c->slot[0], out_rgba, stats);
}

extern "C" int32_t dipsb_push_frame_pipelined(dipsb_ctx* c, const uint8_t* px, uint32_t width, uint32_t height, uint32_t stride,
                                              int32_t format, uint8_t* out_rgba_prev, dipsb_frame_stats* stats_prev) {
    if (!c || !px) return DIPSB_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    dipsb_ctx::FrameSlot& cur = c->slot[c->next_slot];
    dipsb_ctx::FrameSlot& prev = c->slot[c->next_slot ^ 1];
    if (cur.pending) return fail(c, DIPSB_ERR_STATE, "push_frame_pipelined: slot still in flight");
    // output planes are requested for every frame in this mode (the caller decides at collection time); a caller that
    // hands in page-locked output buffers gets the read-back issued at collection, straight into its buffer
    if (out_rgba_prev) c->out_pinned_hint = host_pinned(out_rgba_prev);
    // the previous frame's read-back and this frame's upload run side by side (separate copy engines); submit_frame issues both
    // before it does anything else
    int32_t rc = submit_frame(c, cur, px, width, height, stride, format, true, true, nullptr, c->out_pinned_hint, &prev, out_rgba_prev,
                              stats_prev != nullptr);
    if (rc) return rc;
    c->next_slot ^= 1;
    if (!prev.pending) return DIPSB_NOT_READY;          // first call: nothing to hand back yet
    rc = collect_frame(c, prev, out_rgba_prev, stats_prev);
    return rc < 0 ? rc : (rc == DIPSB_NOT_READY ? 2 : DIPSB_OK);
}

extern "C" int32_t dipsb_flush_frame(dipsb_ctx* c, uint8_t* out_rgba, dipsb_frame_stats* stats) {
    if (!c) return DIPSB_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    dipsb_ctx::FrameSlot& last = c->slot[c->next_slot ^ 1];
    if (!last.pending) return DIPSB_NOT_READY;
    int32_t rc = collect_frame(c, last, out_rgba, stats);
    return rc < 0 ? rc : (rc == DIPSB_NOT_READY ? 2 : DIPSB_OK);
}

// ---- results -------------------------------------------------------------------------------------------------------
extern "C" uint64_t dipsb_frames_processed(const dipsb_ctx* c) { return c ? c->frames_processed : 0; }

extern "C" int32_t dipsb_get_accumulators(dipsb_ctx* c, uint32_t* acc_sum, uint32_t* acc_cnt) {
    if (!c) return DIPSB_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    { int32_t zrc = ensure_acc_zero(c); if (zrc) return zrc; }
    if (c->acc_sharded) return fail(c, DIPSB_ERR_STATE, "%s: the totals are sharded by pixel range over the ranks; call dipsb_gather_accumulators (collective) first", __func__);
    const Geometry& g = c->g;
    if (acc_sum) {
        CK(c, launch_unpermute(g, c->acc, c->planar, c->stream));
        CK(c, cudaMemcpyAsync(acc_sum, c->planar, g.npx * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    }
    if (acc_cnt) {
        CK(c, launch_unpermute(g, c->acc + g.n_elems, c->planar + g.npx, c->stream));
        CK(c, cudaMemcpyAsync(acc_cnt, c->planar + g.npx, g.npx * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    }
    CK(c, cudaStreamSynchronize(c->stream));
    return DIPSB_OK;
}

extern "C" int32_t dipsb_set_accumulators(dipsb_ctx* c, const uint32_t* acc_sum, const uint32_t* acc_cnt) {
    if (!c) return DIPSB_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    { int32_t zrc = ensure_acc_zero(c); if (zrc) return zrc; }
    if (c->acc_sharded) return fail(c, DIPSB_ERR_STATE, "%s: the totals are sharded by pixel range over the ranks; call dipsb_gather_accumulators (collective) first", __func__);
    const Geometry& g = c->g;
    if (acc_sum) {
        CK(c, cudaMemcpyAsync(c->planar, acc_sum, g.npx * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
        CK(c, launch_permute(g, c->planar, c->acc, c->stream));
    }
    if (acc_cnt) {
        CK(c, cudaMemcpyAsync(c->planar + g.npx, acc_cnt, g.npx * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
        CK(c, launch_permute(g, c->planar + g.npx, c->acc + g.n_elems, c->stream));
    }
    CK(c, cudaStreamSynchronize(c->stream));
    return DIPSB_OK;
}

extern "C" int32_t dipsb_accumulators_device(dipsb_ctx* c, void** d_acc, uint64_t* n_elems) {
    if (!c || !d_acc || !n_elems) return DIPSB_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    { int32_t zrc = ensure_acc_zero(c); if (zrc) return zrc; }
    *d_acc = c->acc;
    *n_elems = c->g.n_elems;
    return DIPSB_OK;
}

int dipsb::bit_length(uint64_t v) { int b = 0; while (v) { ++b; v >>= 1; } return b; }

extern "C" int32_t dipsb_pack_accumulators_device(dipsb_ctx* c, uint64_t total_frames, void** d_packed, uint64_t* n_words) {
    if (!c || !d_packed || !n_words || total_frames == 0) return DIPSB_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    const Geometry& g = c->g;
    { int32_t zrc = ensure_acc_zero(c); if (zrc) return zrc; }
    if (total_frames < c->frames_processed)
        return fail(c, DIPSB_ERR_INVALID, "pack_accumulators: total_frames %llu < the %llu frames this context accumulated since its last reset "
                    "(it must bound every frame accumulated on ALL ranks, or the packed fields carry into each other)",
                    (unsigned long long)total_frames, (unsigned long long)c->frames_processed);
    // after the sum over all ranks: acc_sum <= 510*total_frames, acc_cnt <= total_frames (every frame counts at most once)
    const int sum_bits = bit_length(510ull * total_frames), cnt_bits = bit_length(total_frames);
    int layout;
    if (sum_bits + cnt_bits <= 32) layout = 1;
    else if (total_frames < 65536 && sum_bits <= 32) layout = 2;
    else {   // totals too large for a packed format: exchange the planes themselves
        c->xchg_layout = 0;
        *d_packed = c->acc;
        *n_words = 2 * g.n_elems;
        return DIPSB_OK;
    }
    if (!c->xchg) CK(c, cudaMalloc(&c->xchg, (g.n_elems + g.n_elems / 2) * sizeof(uint32_t)));
    CK(c, launch_pack_acc(g, c->acc, c->xchg, layout, sum_bits, c->stream));
    c->xchg_layout = layout; c->xchg_sum_bits = sum_bits;
    *d_packed = c->xchg;
    *n_words = layout == 1 ? g.n_elems : g.n_elems + g.n_elems / 2;
    return DIPSB_OK;
}

extern "C" int32_t dipsb_unpack_accumulators_device(dipsb_ctx* c) {
    if (!c) return DIPSB_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    if (c->xchg_layout == 0) return DIPSB_OK;   // the planes were exchanged in place
    CK(c, launch_unpack_acc(c->g, c->xchg, c->acc, c->xchg_layout, c->xchg_sum_bits, c->stream));
    c->xchg_layout = 0;
    return DIPSB_OK;
}

extern "C" int32_t dipsb_get_scalars(dipsb_ctx* c, uint64_t first, uint64_t n, uint64_t* sad, uint64_t* cnt) {
    if (!c) return DIPSB_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    if (n == 0) return DIPSB_OK;
    if (first + n > c->scal_hi) return fail(c, DIPSB_ERR_INVALID, "get_scalars: frames [%llu,%llu) not processed (have %llu)", (unsigned long long)first, (unsigned long long)(first + n), (unsigned long long)c->scal_hi);
    if (sad) CK(c, cudaMemcpyAsync(sad, c->d_sad + first, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    if (cnt) CK(c, cudaMemcpyAsync(cnt, c->d_cnt + first, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return DIPSB_OK;
}

extern "C" int32_t dipsb_get_intensity_map(dipsb_ctx* c, uint64_t n_eff, float* out) {
    if (!c || !out) return DIPSB_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    { int32_t zrc = ensure_acc_zero(c); if (zrc) return zrc; }
    if (c->acc_sharded) return fail(c, DIPSB_ERR_STATE, "%s: the totals are sharded by pixel range over the ranks; call dipsb_gather_accumulators (collective) first", __func__);
    float* d = reinterpret_cast<float*>(c->planar);
    CK(c, launch_intensity_map(c->g, c->acc, n_eff, d, c->stream));
    CK(c, cudaMemcpyAsync(out, d, c->g.npx * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return DIPSB_OK;
}

extern "C" int32_t dipsb_get_frame_means(dipsb_ctx* c, uint64_t first, uint64_t n, float* out) {
    if (!c || (!out && n)) return DIPSB_ERR_INVALID;
    if (n == 0) return DIPSB_OK;
    uint64_t* tmp = new (std::nothrow) uint64_t[n];
    if (!tmp) return fail(c, DIPSB_ERR_NOMEM, "get_frame_means: out of host memory");
    int32_t rc = dipsb_get_scalars(c, first, n, tmp, nullptr);
    if (rc == DIPSB_OK) {
        const double den = 510.0 * (double)c->g.npx;   // one division of the exact integer, as the oracle does
        for (uint64_t i = 0; i < n; ++i) out[i] = (float)((double)tmp[i] / den);
    }
    delete[] tmp;
    return rc;
}

// ---- utilities -----------------------------------------------------------------------------------------------------
extern "C" int32_t dipsb_synth_fill_device(int32_t device, void* d_dst, uint64_t first_frame, uint64_t n_frames, uint32_t width,
                                           uint32_t height, int32_t format, uint64_t seed, int32_t profile, void* stream) {
    if (!d_dst || format < 0 || format > 3) return DIPSB_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return DIPSB_ERR_CUDA;
    cudaError_t e = launch_synth((uint8_t*)d_dst, first_frame, n_frames, width, height, bpp_of(format), seed, profile, (cudaStream_t)stream);
    return e == cudaSuccess ? DIPSB_OK : fail(nullptr, DIPSB_ERR_CUDA, "synth_fill: %s", cudaGetErrorString(e));
}

extern "C" int32_t dipsb_host_alloc(int32_t device, uint64_t bytes, void** out) {
    if (!out || bytes == 0) return fail(nullptr, DIPSB_ERR_INVALID, "host_alloc: null or empty request");
    *out = nullptr;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return fail(nullptr, DIPSB_ERR_CUDA, "host_alloc: no CUDA device %d", device); }
    cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) { cudaGetLastError(); *out = nullptr; return fail(nullptr, DIPSB_ERR_NOMEM, "host_alloc(%llu): %s", (unsigned long long)bytes, cudaGetErrorString(e)); }
    return DIPSB_OK;
}

extern "C" int32_t dipsb_host_free(void* p) {
    if (!p) return DIPSB_OK;
    cudaError_t e = cudaFreeHost(p);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(nullptr, DIPSB_ERR_INVALID, "host_free: %s", cudaGetErrorString(e)); }
    return DIPSB_OK;
}

extern "C" int32_t dipsb_host_register(int32_t device, void* p, uint64_t bytes) {
    if (!p || bytes == 0) return fail(nullptr, DIPSB_ERR_INVALID, "host_register: null or empty range");
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return fail(nullptr, DIPSB_ERR_CUDA, "host_register: no CUDA device %d", device); }
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(nullptr, e == cudaErrorMemoryAllocation ? DIPSB_ERR_NOMEM : DIPSB_ERR_INVALID, "host_register(%p, %llu): %s", p, (unsigned long long)bytes, cudaGetErrorString(e));
    }
    return DIPSB_OK;
}

extern "C" int32_t dipsb_host_unregister(void* p) {
    if (!p) return fail(nullptr, DIPSB_ERR_INVALID, "host_unregister: null pointer");
    cudaError_t e = cudaHostUnregister(p);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(nullptr, DIPSB_ERR_INVALID, "host_unregister(%p): %s", p, cudaGetErrorString(e)); }
    return DIPSB_OK;
}

extern "C" int32_t dipsb_host_copy2d(void* dst, uint64_t dpitch, const void* src, uint64_t spitch, uint64_t row_bytes, uint64_t rows) {
    if (!rows || !row_bytes) return DIPSB_OK;
    if (!dst || !src || dpitch < row_bytes || spitch < row_bytes) return fail(nullptr, DIPSB_ERR_INVALID, "host_copy2d: null pointer or pitch smaller than a row");
    host_copy2d(dst, dpitch, src, spitch, row_bytes, rows);
    return DIPSB_OK;
}

extern "C" uint32_t dipsb_host_copy_threads(void) { return host_copy_threads(); }

extern "C" int32_t dipsb_enable_timing(dipsb_ctx* c, int32_t on) {
    if (!c) return DIPSB_ERR_INVALID;
    c->timing = on != 0;
    return DIPSB_OK;
}

extern "C" int32_t dipsb_clip_kernel_time(dipsb_ctx* c, double* total_ms, uint64_t* launches) {
    if (!c || !total_ms || !launches) return DIPSB_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaStreamSynchronize(c->stream));
    double sum = 0.0;
    for (size_t i = 0; i + 1 < c->tev_used; i += 2) {
        float ms = 0.f;
        CK(c, cudaEventElapsedTime(&ms, c->tev[i], c->tev[i + 1]));
        sum += ms;
    }
    *total_ms = sum;
    *launches = c->tev_used / 2;
    c->tev_used = 0;
    return DIPSB_OK;
}

// Roofline probe (measurement aid): stream the clip through the clip kernel's TMA ring without computing anything and
// report the mean kernel time of `reps` launches; results of the context are not touched.
extern "C" int32_t dipsb_stream_probe(dipsb_ctx* c, const void* d_frames, uint64_t n_frames, uint64_t stride, uint32_t reps, float* ms) {
    if (!c || !d_frames || !ms || !reps || !n_frames) return DIPSB_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    const Geometry& g = c->g;
    if ((((uintptr_t)d_frames | stride | (g.npx * g.bpp)) & 15u) != 0) return fail(c, DIPSB_ERR_INVALID, "stream_probe: clip not 16-byte aligned");
    cudaEvent_t e0, e1;
    CK(c, cudaEventCreate(&e0));
    CK(c, cudaEventCreate(&e1));
    CK(c, launch_stream_probe(g, (const uint8_t*)d_frames, stride, (uint32_t)n_frames, c->stream));   // warm-up
    CK(c, cudaEventRecord(e0, c->stream));
    for (uint32_t r = 0; r < reps; ++r) CK(c, launch_stream_probe(g, (const uint8_t*)d_frames, stride, (uint32_t)n_frames, c->stream));
    CK(c, cudaEventRecord(e1, c->stream));
    CK(c, cudaEventSynchronize(e1));
    float total = 0.f;
    CK(c, cudaEventElapsedTime(&total, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *ms = total / (float)reps;
    return DIPSB_OK;
}

extern "C" uint64_t dipsb_launch_count(void) { return launch_count_value(); }
