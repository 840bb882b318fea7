/*
 * dips_b200.h -- C ABI of libdips_b200.so: the B200 (sm_100a) implementation of the DiPs per-pixel
 * frame-difference hot path.  Plain pointers and sizes only; no C++/torch types cross this boundary.
 *
 * What each entry point replaces in the reference (paths relative to the RubenMovsesyan/DiPs root):
 *   dipsb_create / dipsb_destroy      ComputeState::new                  dips/src/gpu/mod.rs:59-165
 *                                     DiPsCompute::new                   dips_alt/src/dips_compute/mod.rs:270-496
 *   dipsb_push_frame                  frame_callback = add_texture+dispatch   dips/src/lib.rs:233-246,
 *                                                                        dips/src/gpu/mod.rs:170-216, :306-397
 *                                     DiPsCompute::send_frame            dips_alt/src/dips_compute/mod.rs:498-646
 *   dipsb_snapshot                    send_frame(.., snapshot = Some(())) dips_alt/src/lib.rs:222-225, :636-639
 *                                     and the refresh markers            dips_alt/src/lib.rs:668-670
 *   dipsb_prime_device                pre_compute_main ("start" plane)   dips/src/gpu/shaders/pre_compute_shader.wgsl:92-132
 *   dipsb_run_clip_device / _host     compute_main applied to a whole clip    dips/src/gpu/shaders/dips_shader.wgsl:172-240
 *                                     (the per-frame dispatch loop of    dips/src/frame_extractor.rs:206-276)
 *   dipsb_config fields               pipeline-overridable constants     dips/src/gpu/mod.rs:101-109, dips_shader.wgsl:15-21
 *
 * Semantics (integer contract):
 *   I2 = max(r,g,b)+min(r,g,b) in [0,510]  (2*channel with a chroma filter)
 *   overall:   D_t = |I2_t - I2_ref|;  per-frame: D_t = |I2_t - I2_{t-1}|, D_0 = 0
 *   M_t = D_t > threshold;  acc_sum[p] += D_t(p);  acc_cnt[p] += M_t(p);  sad[t] = sum_p D_t;  cnt[t] = sum_p M_t
 *
 * The context keeps a "state plane" (u16 I2 per pixel): the reference plane in overall mode, the I2 of the
 * last processed frame in per-frame mode.  Consecutive run/push calls therefore chain: a clip may be fed in
 * chunks, and a frame-range shard on another GPU is primed with frame 0 (overall) or with the one-frame halo
 * t0-1 (per-frame) through dipsb_prime_device before its first run call.
 *
 * Threading: one caller at a time per context; the context may migrate between threads (every entry point
 * selects its device).  Errors: 0 = ok, 1 = DIPSB_NOT_READY (warm-up / passthrough), <0 = error with a message
 * in dipsb_last_error().  Nothing throws or aborts across this boundary, and there is no CPU fallback: without
 * a usable CUDA device dipsb_create fails.
 */
#ifndef DIPS_B200_H
#define DIPS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DIPSB_ABI_VERSION 2
/* frames one context may accumulate between resets: 510 * this still fits the u32 per-pixel sums (~39 h of 60 fps video);
 * a call that would exceed it fails with DIPSB_ERR_STATE instead of wrapping */
#define DIPSB_MAX_ACCUMULATED_FRAMES 8421504ull

typedef struct dipsb_ctx dipsb_ctx;

enum dipsb_status {
    DIPSB_OK = 0,
    DIPSB_NOT_READY = 1,
    DIPSB_ERR_INVALID = -1,
    DIPSB_ERR_CUDA = -2,
    DIPSB_ERR_NOMEM = -3,
    DIPSB_ERR_STATE = -4
};

enum dipsb_format { DIPSB_FMT_RGB8 = 0, DIPSB_FMT_RGBX8 = 1, DIPSB_FMT_BGR8 = 2, DIPSB_FMT_BGRX8 = 3 };
enum dipsb_mode { DIPSB_MODE_OVERALL = 0, DIPSB_MODE_PERFRAME = 1 };
/* numbering of dips/src/lib.rs:52-61 */
enum dipsb_chroma { DIPSB_CHROMA_NONE = 0, DIPSB_CHROMA_RED = 1, DIPSB_CHROMA_GREEN = 2, DIPSB_CHROMA_BLUE = 3 };
/* numbering of dips/src/lib.rs:32-41 */
enum dipsb_filter { DIPSB_FILTER_SIGMOID = 0, DIPSB_FILTER_INV_SIGMOID = 1, DIPSB_FILTER_NONE = 255 };
enum dipsb_synth { DIPSB_SYNTH_UNIFORM = 0, DIPSB_SYNTH_SCENE = 1 };
/*
 * Temporal semantics of dipsb_push_frame:
 *   FRAME0        north star: the reference is the first frame pushed (or a prime / snapshot), temporal window 1.
 *   DIPS_RING4    the `dips` crate as shipped: the first 3 frames pass through; start plane = grey(upper median of the
 *                 first 4 frames); every output = start - median(ring of 4), the newest ring slot quantised to grey in
 *                 place (dips/src/gpu/mod.rs:170-216, bind_groups.rs:407-427, dips_shader.wgsl:187-214).
 *   ALT_RING2     `dips_alt` as shipped: ring of 2 frames, "median" = min of the two (zero-sentinel sort,
 *                 pre_compute_shader.wgsl:212-227); dipsb_snapshot() makes the next frame store and return the grey
 *                 snapshot (dips_alt/src/lib.rs:222-225); until then the snapshot plane is zero.
 *   ALT_RING2_MEDIAN  the same with the in-bounds median (max of the two).
 * The ring flavours exist for drop-in parity with the crates as shipped.  dipsb_run_clip_device / _host accept them too
 * (accumulators and scalars, no visual output): the warm-up / start-plane / snapshot frames run frame by frame exactly as
 * dipsb_push_frame runs them, the steady state after them in one launch per run of frames with the ring held in registers
 * (ring_clip_kernel).  Per-frame calls and batch calls may be mixed on one context.  Sharded passes need FRAME0.
 */
enum dipsb_flavor { DIPSB_FLAVOR_FRAME0 = 0, DIPSB_FLAVOR_DIPS_RING4 = 1, DIPSB_FLAVOR_ALT_RING2 = 2, DIPSB_FLAVOR_ALT_RING2_MEDIAN = 3 };

typedef struct dipsb_config {
    uint32_t struct_size;      /* sizeof(dipsb_config), for ABI evolution */
    int32_t device;            /* CUDA device ordinal */
    uint32_t width, height;    /* pixels */
    int32_t format;            /* dipsb_format of the frames fed to run_clip / push_frame */
    int32_t mode;              /* dipsb_mode */
    int32_t chroma;            /* dipsb_chroma */
    uint32_t threshold;        /* tau in I2 units [0,510]; a float theta in [0,1] maps to floor(theta*510) */
    /* visual output of dipsb_push_frame (reference compute_main colour mapping) */
    int32_t colorize;          /* COLORIZE override, dips_shader.wgsl:15 */
    int32_t filter;            /* dipsb_filter, FILTER_TYPE override */
    float sigmoid_scalar;      /* SIGMOID_HORIZONTAL_SCALAR override (UI "sensitivity") */
    int32_t spatial_window;    /* WINDOW_SIZE override: 1, 3, 5 or 7.  > 1 applies a CORRECT zero-padded median of the full
                                * window to every frame's intensity (the reference's own loop is defective);
                                * such contexts run frame by frame (no clip kernel) */
    int32_t flavor;            /* dipsb_flavor */
    uint32_t reserved[3];
} dipsb_config;

typedef struct dipsb_frame_stats {
    uint64_t frame_index;
    uint64_t sad;              /* sum_p D_t(p) */
    uint64_t count;            /* sum_p M_t(p) */
} dipsb_frame_stats;

/* ---- lifetime ------------------------------------------------------------------------------- */
int32_t dipsb_abi_version(void);
void dipsb_default_config(dipsb_config *cfg);
int32_t dipsb_create(const dipsb_config *cfg, dipsb_ctx **out);
void dipsb_destroy(dipsb_ctx *ctx);
const char *dipsb_last_error(const dipsb_ctx *ctx);      /* ctx may be NULL: last error of dipsb_create */
/* zero the accumulators and scalars, forget the state plane and the frame counter */
int32_t dipsb_reset(dipsb_ctx *ctx);
/* change tau / mode between clips (accumulators are kept; call dipsb_reset to clear) */
int32_t dipsb_set_threshold(dipsb_ctx *ctx, uint32_t threshold);
/* all further work of this context is issued on `stream` (a cudaStream_t; NULL = the CUDA default stream); it is ordered
 * after the work already issued (the new stream waits on an event of the old one; the host is not blocked) */
int32_t dipsb_set_stream(dipsb_ctx *ctx, void *stream);
/* same without any ordering: for callers that pipeline several contexts over several streams and order them with their
 * own events (an implicit wait here would serialise unrelated work that was queued on the old stream meanwhile) */
int32_t dipsb_adopt_stream(dipsb_ctx *ctx, void *stream);
/* go back to the context's private non-blocking stream (the state after dipsb_create) */
int32_t dipsb_use_private_stream(dipsb_ctx *ctx);
int32_t dipsb_synchronize(dipsb_ctx *ctx);

/* ---- state plane (reference frame / halo) --------------------------------------------------- */
/* state := I2(frame) for one raw frame resident on the device (K1). */
int32_t dipsb_prime_device(dipsb_ctx *ctx, const void *d_frame);
/* state := upper median of the I2 of 4 consecutive raw frames (reference "start" plane). */
int32_t dipsb_prime_median4_device(dipsb_ctx *ctx, const void *d_frames, uint64_t frame_stride_bytes);
int32_t dipsb_prime_host(dipsb_ctx *ctx, const uint8_t *frame);
/* device pointer to the u16[width*height] state plane (for an NCCL broadcast / send-recv) and a setter
 * that marks it valid after a peer wrote it. */
int32_t dipsb_state_plane_device(dipsb_ctx *ctx, void **d_state);
int32_t dipsb_mark_state_valid(dipsb_ctx *ctx, int32_t valid);
int32_t dipsb_get_state_plane(dipsb_ctx *ctx, uint16_t *out);

/* ---- batch: a clip (or a frame-range shard of one) ------------------------------------------ */
/*
 * Frames k = 0..n_frames-1 live at d_frames + k*frame_stride_bytes, tightly packed rows.  They are frames
 * first_frame_index .. first_frame_index+n_frames-1 of the logical clip; per-frame scalars are stored under
 * that index.  If the state plane is not valid it is primed from frame 0 of this call (so D of that frame is 0).
 * Asynchronous with respect to the host; ordered on the context's stream.  Fastest when d_frames, frame_stride_bytes and
 * width*height*bpp are multiples of 16 bytes (the clip is then streamed in place); any other layout is first re-packed on
 * the device into an aligned scratch of at most 256 MB, chunk by chunk -- same results.
 */
int32_t dipsb_run_clip_device(dipsb_ctx *ctx, const void *d_frames, uint64_t n_frames,
                              uint64_t frame_stride_bytes, uint64_t first_frame_index);
/* Same from pageable or pinned HOST memory: chunks are staged through pinned buffers and uploaded on a copy
 * stream overlapped with the kernels of the previous chunk.  Returns after the last kernel was issued; `frames` is borrowed
 * for the call only (ordinary memory has been staged, a page-locked clip has been uploaded when the call returns).  Calls
 * may follow each other without a synchronisation in between: the staging slots are guarded across calls. */
int32_t dipsb_run_clip_host(dipsb_ctx *ctx, const uint8_t *frames, uint64_t n_frames,
                            uint64_t frame_stride_bytes, uint64_t first_frame_index);

/* ---- streaming: one frame per call (the reference's per-frame boundary) --------------------- */
/*
 * px: host frame, `stride` bytes per row.  out_rgba (nullable): width*height*4 bytes, receives the reference's
 * visual frame for this input (RGBA8, alpha 255).  stats (nullable).  Returns DIPSB_NOT_READY and copies the
 * input through (converted to RGBA8) while there is no reference yet, as frame_callback does during warm-up.
 * Both buffers are borrowed for the call only.  Ordinary host memory is staged through page-locked buffers by the threaded
 * host copy; page-locked buffers (dipsb_host_alloc) are read and written by the copy engine directly.  Frames of 2 MB and
 * more are processed in up to 4 row bands (DIPSB_FRAME_BANDS) so that upload, kernels and read-back overlap.
 */
int32_t dipsb_push_frame(dipsb_ctx *ctx, const uint8_t *px, uint32_t width, uint32_t height, uint32_t stride,
                         int32_t format, uint8_t *out_rgba, dipsb_frame_stats *stats);
/*
 * Pipelined variant (opt-in, one frame of latency): submits frame t -- staging copy, upload on a copy
 * stream, kernel, read-back -- and hands back the output of frame t-1, so the host-side copies and the upload of frame t
 * overlap the GPU work of frame t-1.  Returns DIPSB_NOT_READY on the first call (nothing to return yet), DIPSB_OK when
 * out_rgba_prev/stats_prev hold the difference frame of t-1, 2 (DIPSB_PASSTHROUGH) when they hold a passed-through
 * warm-up frame.  dipsb_flush_frame collects the last frame in flight.  Do not mix with dipsb_push_frame without flushing.
 */
#define DIPSB_PASSTHROUGH 2
int32_t dipsb_push_frame_pipelined(dipsb_ctx *ctx, const uint8_t *px, uint32_t width, uint32_t height, uint32_t stride,
                                   int32_t format, uint8_t *out_rgba_prev, dipsb_frame_stats *stats_prev);
int32_t dipsb_flush_frame(dipsb_ctx *ctx, uint8_t *out_rgba, dipsb_frame_stats *stats);
/*
 * dipsb_push_frame in the two steps the reference's ComputeState takes (add_texture dips/src/gpu/mod.rs:170 keeps the
 * borrowed frame, dispatch :306 computes and returns it): stage copies the frame straight into the library's page-locked
 * input slot and starts the upload (px is free again on return), dispatch runs the kernels and hands back what
 * dipsb_push_frame would have.  A wrapper built on this pair keeps no frame copy of its own.
 */
int32_t dipsb_stage_frame(dipsb_ctx *ctx, const uint8_t *px, uint32_t width, uint32_t height, uint32_t stride, int32_t format);
int32_t dipsb_dispatch_staged(dipsb_ctx *ctx, uint8_t *out_rgba, dipsb_frame_stats *stats);
/* the next pushed frame becomes the reference (dips_alt snapshot / refresh marker) */
int32_t dipsb_snapshot(dipsb_ctx *ctx);

/* ---- results -------------------------------------------------------------------------------- */
uint64_t dipsb_frames_processed(const dipsb_ctx *ctx);
/* planar row-major u32[width*height]; either pointer may be NULL.  Synchronises the stream. */
int32_t dipsb_get_accumulators(dipsb_ctx *ctx, uint32_t *acc_sum, uint32_t *acc_cnt);
int32_t dipsb_set_accumulators(dipsb_ctx *ctx, const uint32_t *acc_sum, const uint32_t *acc_cnt);
/*
 * Device view of the accumulators for collectives: one contiguous u32[2*n_elems] buffer (sum plane then count
 * plane, in the library's internal tile order -- identical on every context of the same geometry, so an
 * element-wise NCCL sum across ranks is exact).  Finalises pending work first.
 */
int32_t dipsb_accumulators_device(dipsb_ctx *ctx, void **d_acc, uint64_t *n_elems);
/*
 * The same for a cheaper cross-GPU sum: packs the accumulators of this context into an exchange buffer whose element-wise
 * int32 sum over all ranks equals the packed sum of the fields -- `total_frames` = number of frames of the WHOLE clip (all
 * ranks), which bounds the totals: one u32 per element (sum | count << bits) while both fit 32 bits (<= 2056 frames),
 * else the sum plane + counts as u16 pairs (6 bytes per element, < 65536 frames), else the planes themselves.  All-reduce
 * (sum, int32) the n_words returned, then call dipsb_unpack_accumulators_device on every rank.
 */
int32_t dipsb_pack_accumulators_device(dipsb_ctx *ctx, uint64_t total_frames, void **d_packed, uint64_t *n_words);
int32_t dipsb_unpack_accumulators_device(dipsb_ctx *ctx);
/* per-frame scalars of logical frames first..first+n-1; either pointer may be NULL. */
int32_t dipsb_get_scalars(dipsb_ctx *ctx, uint64_t first, uint64_t n, uint64_t *sad, uint64_t *cnt);
/* X6 float outputs: acc_sum/(510*n_eff) per pixel, sad[t]/(510*W*H) per frame */
int32_t dipsb_get_intensity_map(dipsb_ctx *ctx, uint64_t n_eff, float *out);
int32_t dipsb_get_frame_means(dipsb_ctx *ctx, uint64_t first, uint64_t n, float *out);

/* ---- several GPUs: frame-range shards of one clip ------------------------------------------- */
/*
 * The reference drives one adapter (dips/src/gpu/mod.rs:71-78); long clips shard naturally by frame range over the GPUs of
 * a box.  Rank r of R owns the contiguous frames dipsb_shard_range gives it and runs them through its own
 * context; per clip the library exchanges
 *   overall mode    the u16 reference plane of frame 0, before the pass: rank 0's prime kernel stores slice j of it straight
 *                   into rank j over NVLink and starts at once, the other ranks forward their slices to each other
 *                   (scatter + all-gather over peer memory); or one ncclBroadcast (DIPSB_REDUCE_NCCL);
 *   per-frame mode  the one-frame halo: every rank starts at once from its own first frame while its copy engine pushes
 *                   that frame over NVLink to the previous rank, whose clip kernel differences it as one extra trailing
 *                   frame (nothing is exchanged before the pass; the boundary frame's scalars are handed to the rank
 *                   that owns the frame with the accumulator exchange);
 *   at the end      the per-pixel accumulators: a reduce-scatter written by the library's own kernels over peer memory
 *                   (packed partial sums stored straight into the owner's window over NVLink), after which every rank
 *                   holds the totals of the pixel range it owns until dipsb_gather_accumulators (a collective all-gather
 *                   over the same peer memory) completes the planes everywhere.  Without peer memory, or on request
 *                   (DIPSB_REDUCE_NCCL): pack -> ncclAllReduce -> unpack, totals replicated at once.
 * Per-frame scalars stay on the rank that owns the frame.  Integer sums: the result is bit-identical to one GPU.
 *
 * Two ways to form the ranks:
 *   one process per GPU   rank 0 calls dipsb_comm_unique_id and hands the 128 bytes to the others by any means (a file,
 *                         MPI, torchrun's store); every rank calls dipsb_comm_init_rank on its context (collective:
 *                         ncclCommInitRank + mapping of the peers' windows with cudaIpc*), then, per clip,
 *                         dipsb_reset + dipsb_run_clip_sharded_device/_host.
 *   one process, all GPUs dipsb_create_group makes one context per device, ncclCommInitAll and peer access; the
 *                         dipsb_group_* calls drive all ranks from the calling thread.
 * NCCL is loaded at run time (dlopen libnccl.so.2) by these calls only.
 */
#define DIPSB_UNIQUE_ID_BYTES 128
enum dipsb_reduce_path { DIPSB_REDUCE_AUTO = 0, DIPSB_REDUCE_P2P = 1, DIPSB_REDUCE_NCCL = 2 };
typedef struct dipsb_group dipsb_group;

/* frames [*first, *first + *count) of a total_frames clip owned by `rank` of `nranks` (contiguous, disjoint, covering) */
void dipsb_shard_range(uint64_t total_frames, uint32_t nranks, uint32_t rank, uint64_t *first, uint64_t *count);
/* host-only: exchange format of the peer-memory reduce for such a clip: out[0] bytes per accumulator element (4: sum |
 * count << out[1] in one u32, 8: two u32), out[1] sum bits, out[2] frames a rank may difference (ceil(total/nranks) + 1),
 * out[3] accumulator elements owned per rank for planes of n_elems elements */
int32_t dipsb_xchg_plan_query(uint64_t total_frames, uint32_t nranks, uint64_t n_elems, uint64_t out[4]);
int32_t dipsb_comm_unique_id(void *id128);
int32_t dipsb_comm_init_rank(dipsb_ctx *ctx, uint32_t nranks, uint32_t rank, const void *id128);
int32_t dipsb_comm_destroy(dipsb_ctx *ctx);
/* out: [0] ranks, [1] rank, [2] 1 = peers' windows mapped (peer-memory kernels available), [3] NCCL version, [4] the
 * dipsb_reduce_path in effect, [5] 1 = single-process group, [6] 1 = the accumulators currently hold totals only in this
 * rank's owned range, [7] 1 = NCCL communicator present */
int32_t dipsb_comm_info(const dipsb_ctx *ctx, uint32_t out[8]);
int32_t dipsb_comm_set_reduce(dipsb_ctx *ctx, int32_t path);
/* DIPSB_ERR_STATE if a bounded wait for a peer GPU timed out since the last check (DIPSB_COMM_TIMEOUT_MS, default 5000):
 * a rank never arrived and the results are invalid.  Synchronises the stream. */
int32_t dipsb_comm_check(dipsb_ctx *ctx);
/*
 * One pass over this rank's shard of a total_frames clip: the n_frames frames starting at logical index
 * first_frame_index (at most ceil(total_frames / ranks)); collective -- every rank of the communicator calls it once per
 * dipsb_reset.  Asynchronous; ordered on the context's stream (the halo push uses the context's copy stream).
 */
int32_t dipsb_run_clip_sharded_device(dipsb_ctx *ctx, const void *d_frames, uint64_t n_frames, uint64_t frame_stride_bytes,
                                      uint64_t first_frame_index, uint64_t total_frames);
int32_t dipsb_run_clip_sharded_host(dipsb_ctx *ctx, const uint8_t *frames, uint64_t n_frames, uint64_t frame_stride_bytes,
                                    uint64_t first_frame_index, uint64_t total_frames);
/* with dipsb_enable_timing: summed milliseconds of the three phases of the sharded passes since the last call -- [0] reference
 * / halo exchange before the pass, [1] the pass (prime, clip kernel, scalars), [2] accumulator exchange -- and their number.
 * Event pairs on the context's stream; a phase that waits for a slower rank contains that wait.  Synchronises. */
int32_t dipsb_comm_phase_times(dipsb_ctx *ctx, double out_ms[3], uint64_t *passes);
/* measurement aid (collective, multi-process communicators): mean milliseconds of `reps` back-to-back exchanges without a
 * pass in between -- what 0: accumulator reduce-scatter over peer memory, 1: reference-plane ncclBroadcast, 2: all-gather of
 * the totals, 3: pack + ncclAllReduce + unpack, 4: reference-plane scatter + all-gather over peer memory.  Leaves the
 * accumulators and the state plane undefined: dipsb_reset afterwards. */
int32_t dipsb_comm_probe(dipsb_ctx *ctx, int32_t what, uint64_t total_frames, uint32_t reps, float *ms);
/* collective: complete the accumulator planes on every rank (no-op when they already are); then dipsb_get_accumulators,
 * dipsb_get_intensity_map ... work as on one GPU.  Those calls fail with DIPSB_ERR_STATE while the totals are sharded. */
int32_t dipsb_gather_accumulators(dipsb_ctx *ctx);

/* single process: ndev contexts of the same configuration on devices[0..ndev-1] (NULL: 0..ndev-1; cfg->device is ignored).
 * Listing one device several times gives a loopback group (no NCCL, ranks share the device) for tests on one GPU. */
int32_t dipsb_create_group(const dipsb_config *cfg, uint32_t ndev, const int32_t *devices, dipsb_group **out);
void dipsb_destroy_group(dipsb_group *grp);
uint32_t dipsb_group_size(const dipsb_group *grp);
dipsb_ctx *dipsb_group_ctx(dipsb_group *grp, uint32_t rank);    /* borrowed: per-rank setters, scalars, tuning queries */
const char *dipsb_group_last_error(const dipsb_group *grp);      /* grp may be NULL: last error of dipsb_create_group */
int32_t dipsb_group_reset(dipsb_group *grp);
/* shard i = n_frames[i] frames at d_frames[i] (resident on device i), in clip order; total = their sum */
int32_t dipsb_group_run_clip_device(dipsb_group *grp, const void *const *d_frames, const uint64_t *n_frames,
                                    uint64_t frame_stride_bytes);
int32_t dipsb_group_gather_accumulators(dipsb_group *grp);
int32_t dipsb_group_synchronize(dipsb_group *grp);
/* the clip's combined maps (gathers first when needed) and the scalars of its frames, each from the rank that owns them */
int32_t dipsb_group_get_accumulators(dipsb_group *grp, uint32_t *acc_sum, uint32_t *acc_cnt);
int32_t dipsb_group_get_scalars(dipsb_group *grp, uint64_t first, uint64_t n, uint64_t *sad, uint64_t *cnt);

/* ---- utilities ------------------------------------------------------------------------------ */
/* deterministic synthetic clip generated on the device (same bytes as the oracle's generator) */
int32_t dipsb_synth_fill_device(int32_t device, void *d_dst, uint64_t first_frame, uint64_t n_frames,
                                uint32_t width, uint32_t height, int32_t format, uint64_t seed, int32_t profile,
                                void *stream);
/*
 * Page-locked host buffers for the frame paths.  dipsb_push_frame / dipsb_push_frame_pipelined / dipsb_run_clip_host
 * recognise page-locked pointers (these, cudaHostAlloc / cudaHostRegister memory, torch pin_memory tensors) and let the
 * copy engine read and write them directly; any other host pointer is staged through an internal page-locked buffer with
 * a CPU memcpy first, which for a 1080p frame costs several times the PCIe transfer itself.  A decoder that writes its
 * frames into such a buffer (the role of frame_extractor.rs:216-226's mapped gst buffer) removes that copy.
 */
int32_t dipsb_host_alloc(int32_t device, uint64_t bytes, void **out);
int32_t dipsb_host_free(void *p);
/*
 * The same for memory the caller already owns (a decoder's buffer pool, an mmap'ed file, a Vec kept for the whole run):
 * page-lock [p, p + bytes) in place so that the frame calls take the direct path for any frame inside it.  Costs about a
 * millisecond per 8 MB, so register buffers that are reused, once, not every frame.  The range must stay mapped until
 * dipsb_host_unregister(p) (same p); unregistering memory that was not registered is an error, freeing registered memory
 * without unregistering it leaves the pages locked until the process ends.
 */
int32_t dipsb_host_register(int32_t device, void *p, uint64_t bytes);
int32_t dipsb_host_unregister(void *p);
/*
 * The staging copy the frame calls use for ordinary host memory: rows of row_bytes from src (pitch spitch) to dst (pitch
 * dpitch) on a small persistent thread pool -- the caller plus DIPSB_COPY_THREADS-1 helpers (environment variable, default
 * min(4, cores/2); 1 = the calling thread only); copies under 1 MB stay on the caller.  While copies follow each other at a
 * per-frame pace (less than 8 x DIPSB_COPY_SPIN_US apart) the helpers poll for DIPSB_COPY_SPIN_US microseconds (default 500,
 * 0 = never) after a copy before they block, so that a per-frame caller finds them awake; an isolated copy leaves them
 * asleep.  Host only, no device involved.
 */
int32_t dipsb_host_copy2d(void *dst, uint64_t dpitch, const void *src, uint64_t spitch, uint64_t row_bytes, uint64_t rows);
uint32_t dipsb_host_copy_threads(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
uint64_t dipsb_launch_count(void);
/* last clip kernel geometry, for reports: [0] tiles, [1] frame segments, [2] threads per block, [3] stages | kernel << 16, [4] blocks per
 * SM, [5] pixels per tile, [6] dynamic shared memory bytes | registers << 24, [7] 1 = TMA clip kernel / 2 = ring clip kernel
 * (reference-flavour rings in registers) / 0 = per-frame kernels */
int32_t dipsb_last_plan(const dipsb_ctx *ctx, uint32_t out[8]);
/* optional device-side timing of the clip kernel alone (cudaEvent pairs on the context's stream around each launch).
 * dipsb_clip_kernel_time synchronises, returns the summed milliseconds and launch count since the last call, and resets. */
int32_t dipsb_enable_timing(dipsb_ctx *ctx, int32_t on);
int32_t dipsb_clip_kernel_time(dipsb_ctx *ctx, double *total_ms, uint64_t *launches);
/* roofline probe (measurement aid): streams the clip through the clip kernel's TMA ring with the context's tile plan but
 * computes nothing; *ms = mean kernel time over `reps` launches.  Gives the bandwidth ceiling of the access pattern. */
int32_t dipsb_stream_probe(dipsb_ctx *ctx, const void *d_frames, uint64_t n_frames, uint64_t frame_stride_bytes,
                           uint32_t reps, float *ms);
/* which clip kernel runs the batch path: -1 = automatic (default: clip_kernel_ws whenever the tuning allows it),
 * 0 = clip_kernel (thread 0 of each block issues the TMA copies; any stage count / register variant), 1 = clip_kernel_ws
 * (dedicated producer warp, frame loop unrolled over the pipeline stages; 64 registers, 3 or 4 stages).  Same results. */
int32_t dipsb_set_kernel(dipsb_ctx *ctx, int32_t kernel);
/* host-only: the plan (same layout as dipsb_last_plan, [7] = active warps per block) the library would choose for a
 * geometry on a device with num_sms SMs; touches no device. */
int32_t dipsb_plan_query(uint32_t width, uint32_t height, int32_t format, uint32_t num_sms, uint32_t out[8]);
/* tuning knobs (0 = automatic): pipeline stages, pixels per tile (multiple of 16), frame segments, register variant of
 * the clip kernel (64/72/80/96 registers with 16 pixels per thread, 128 registers with 32).  Geometry knobs can only change on a fresh or reset context. */
int32_t dipsb_set_tuning(dipsb_ctx *ctx, uint32_t stages, uint32_t tile_px, uint32_t segments, uint32_t regs);

#ifdef __cplusplus
}
#endif
#endif
