// Internal declarations shared by the translation units of libdips_b200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace dipsb {

constexpr int kPxPerThread = 16;        // pixels owned by one thread of the clip kernel (8 packed u16x2 registers)
constexpr int kFlushFrames = 128;       // 510*128 < 65536: packed u16 accumulators are flushed to u32 every 128 frames
constexpr int kMaxStages = 8;
constexpr int kMaxRanks = 16;          // GPUs one clip can be sharded over (one box)

// Geometry of one context (fixed at create time so that the internal accumulator order never changes).
struct Geometry {
    uint32_t width, height;
    uint64_t npx;            // width*height
    int bpp;                 // 3 or 4
    int format;              // dipsb_format
    int chan_byte;           // -1: all channels (max+min); else byte offset of the selected channel inside a pixel
    uint32_t threads;        // threads per block (multiple of 32); a block has 16*groups*threads accumulator slots
    uint32_t tile_px;        // pixels per tile: multiple of 16, <= slots (chosen so the tiles fill the SMs evenly)
    uint32_t n_tiles;        // ceil(npx / tile_px)
    uint64_t n_elems;        // n_tiles * slots: length of each accumulator plane in internal (tile) order
    uint64_t state_elems;    // npx + slots: length of a state plane (zero padded past npx)
    uint32_t blocks_per_sm;  // resident blocks per SM the plan assumes
    uint32_t stages;         // pipeline depth
    int regs;                // kernel variant: 64/72/80/96 registers with 16 pixels per thread, 128 registers with 32
    int groups;              // groups of 16 pixels per thread (1 or 2) == clip_groups(regs)
    int kernel;              // 0: clip_kernel (thread 0 of the block produces); 1: clip_kernel_ws (producer warp, stage-unrolled)
    uint32_t num_sms;
};

// Accumulator exchange fused into the clip kernel's last flush (comm.cu): where each owner receives this rank's partials.
struct ShardPush {
    uint32_t nranks = 0, rank = 0;
    uint64_t chunk = 0;              // accumulator elements owned per rank
    int fmt = 1, sum_bits = 0;       // 1: sum | cnt << sum_bits in one u32; 2: two u32 (counts at +chunk)
    uint32_t* recv[kMaxRanks] = {};  // owner r's receive slot for this rank, this pass (unused for r == rank)
};
// true when launch_clip would fuse the push for this geometry / call shape
bool clip_can_push(const Geometry& g, uint32_t n_segments);
// true when launch_clip can take over the zeroing of the accumulator planes for this call shape
bool clip_can_store_first(const Geometry& g, uint32_t n_segments);

struct ClipArgs {
    const uint8_t* frames;       // frame k at frames + k*stride
    uint64_t stride;             // bytes
    uint32_t n_frames;           // frames in this call
    uint32_t n_segments;         // grid.y
    const uint16_t* state_in;    // u16[n_elems >= npx], planar pixel order
    uint16_t* state_out;         // per-frame mode: I2 of the last frame (written by the last segment)
    uint32_t* acc_sum;           // u32[n_elems], internal order, RED target
    uint32_t* acc_cnt;
    uint32_t* partials;          // u32[n_frames][n_tiles*warps]: sad | cnt<<20 per warp per frame
    uint32_t tau;                // clamped to <= 511
    int mode;                    // dipsb_mode
    // frame-range shards, per-frame mode (comm.cu): an extra trailing frame -- the next shard's first frame -- differenced
    // after frame n_frames-1 (scalar row n_frames); the kernel waits for *halo_flag >= halo_epoch before reading it
    const uint8_t* extra_frame = nullptr;
    const unsigned long long* halo_flag = nullptr;
    unsigned long long halo_epoch = 0, wait_timeout_ns = 0;
    uint32_t* status = nullptr;
    const ShardPush* push = nullptr; // last launch of a sharded pass: hand the non-owned totals to their owners
    bool first_store = false;        // the planes were not zeroed: the launch clears them itself (only honoured when clip_can_store_first)
};

// Which pixel of its tile does register slot k (0 .. 16*groups-1) of thread `thread` hold?
//   3 B/px: group g = 16 consecutive pixels at 16*(g*threads + thread); slot k = 16*g + (pixel within the group).
//   4 B/px: quad q = 4 consecutive pixels at 4*(q*threads + thread), q = 0 .. 4*groups-1; slot k = 4*q + (pixel within the quad).
// index of pixel p in the internal accumulator order: tile*(16*groups*threads) + k*threads + thread.
__host__ __device__ inline uint64_t tile_order_index(uint64_t p, uint32_t tile_px, uint32_t threads, int bpp, int groups) {
    const uint64_t tile = p / tile_px;
    const uint32_t q = (uint32_t)(p - tile * tile_px);
    uint32_t thread, k;
    if (bpp == 3) {
        const uint32_t grp = q / kPxPerThread;
        thread = grp % threads;
        k = (grp / threads) * kPxPerThread + q % kPxPerThread;
    } else {
        const uint32_t quad = q / 4u;
        thread = quad % threads;
        k = 4u * (quad / threads) + q % 4u;
    }
    return tile * ((uint64_t)threads * kPxPerThread * groups) + (uint64_t)k * threads + thread;
}

int clip_groups(int regs);   // groups of 16 pixels per thread of a register variant (128 registers -> 2, else 1)
size_t clip_smem_bytes(uint32_t threads, int bpp, uint32_t stages, int regs);
// resident threads per SM allowed by a register variant of the clip kernel (64 -> 1024, 72 -> 896, 80 -> 800, 96 -> 672, 128 -> 512)
int clip_max_threads_per_sm(int regs);
// resident blocks/SM for (threads, stages, register variant), or 0 when it does not fit
int clip_occupancy(uint32_t threads, int bpp, uint32_t stages, int regs);
uint32_t clip_active_warps(const Geometry& g);
// clip_kernel_ws is instantiated for every (bytes per pixel, channel, mode) but 4 B/px + chroma filter + per-frame mode
bool clip_ws_available(int bpp, int chan_byte, int mode);
cudaError_t launch_clip(const Geometry& g, const ClipArgs& a, cudaStream_t s);
cudaError_t launch_stream_probe(const Geometry& g, const uint8_t* frames, uint64_t stride, uint32_t n_frames, cudaStream_t s);

// Scatter step of the reference-plane broadcast, fused into rank 0's prime kernel (comm.cu): the plane is cut into nranks-1
// slices, slice k is also stored into rank k+1's state plane over NVLink, and the kernel's last block stamps every peer.
struct PlaneScatter {
    uint32_t nranks = 0;                          // 0: no scatter
    uint64_t slice_px = 0;                        // pixels per slice (multiple of 16); slice k belongs to rank k + 1
    uint16_t* plane_peer[kMaxRanks] = {};
    unsigned long long* stamp_peer[kMaxRanks] = {};   // null: nobody to tell
    unsigned long long epoch = 0;
    uint32_t* blocks_done = nullptr;
};
// true when the vectorised prime kernel (16 pixels per thread, 128-bit loads) can take this frame
bool prime_fast_path(const Geometry& g, const uint8_t* frame);
cudaError_t launch_prime(const Geometry& g, const uint8_t* frame, uint16_t* state, cudaStream_t s, const PlaneScatter* scatter = nullptr);
cudaError_t launch_prime_median4(const Geometry& g, const uint8_t* frames, uint64_t stride, uint16_t* state,
                                 cudaStream_t s);
cudaError_t launch_finalize_scalars(const Geometry& g, const uint32_t* partials, uint32_t n_frames,
                                    uint32_t words_per_frame, uint64_t* sad, uint64_t* cnt, cudaStream_t s);
cudaError_t launch_unpermute(const Geometry& g, const uint32_t* acc_internal, uint32_t* planar, cudaStream_t s);
cudaError_t launch_permute(const Geometry& g, const uint32_t* planar, uint32_t* acc_internal, cudaStream_t s);

struct FrameArgs {
    const uint8_t* frame;        // one frame, row pitch `pitch` bytes
    uint64_t pitch;
    int format;                  // format of this frame (may differ from the context's in push_frame)
    int chan_byte;
    const uint16_t* i2src = nullptr;   // spatially filtered intensity plane to use instead of the raw frame (window > 1)
    const uint16_t* state_in;
    uint16_t* state_out;         // nullptr: do not update
    uint32_t* acc_sum;
    uint32_t* acc_cnt;
    uint64_t* sad;               // single slots, pre-zeroed
    uint64_t* cnt;
    uint8_t* out_rgba;           // nullable: visual frame
    uint32_t tau;
    int accumulate;              // 0: only prime/visual
    int colorize, filter;
    float sig_scalar;
    uint64_t p_begin = 0, p_end = 0;   // pixel range of this launch (row band); p_end == 0: the whole frame
};
cudaError_t launch_frame(const Geometry& g, const FrameArgs& a, cudaStream_t s);
// reference-flavour temporal rings (N1): one frame against a ring of u16 I2 planes
struct RingArgs {
    const uint8_t* frame; uint64_t pitch; int format; int chan_byte;
    const uint16_t* i2src = nullptr;   // see FrameArgs
    uint16_t* ring;              // n_slots planes of npx u16
    int n_slots;                 // 4 (dips) or 2 (dips_alt)
    int write_slot;              // slot that receives the raw I2 of this frame
    int grey_slot;               // slot quantised to grey in place (dips), -1: none
    int compute_start;           // dips: start plane := grey(upper median of the 4 raw slots)
    int snapshot;                // dips_alt: snapshot plane := grey(median), output = that grey
    int median_is_max;           // dips_alt: 0 = as shipped (min of two), 1 = in-bounds median (max of two)
    int do_diff;                 // produce output / accumulate (0 during warm-up)
    uint16_t* start;             // start / snapshot plane (u16, I2 units, even values)
    uint32_t* acc_sum; uint32_t* acc_cnt; uint64_t* sad; uint64_t* cnt; uint8_t* out_rgba;
    uint32_t tau; int colorize, filter; float sig_scalar;
    uint64_t p_begin = 0, p_end = 0;   // see FrameArgs
};
cudaError_t launch_ring(const Geometry& g, const RingArgs& a, cudaStream_t s);
// ring_clip.cu: the steady state of a reference-flavour ring over a run of frames in one launch (ring in registers)
struct RingClipArgs {
    const uint8_t* frames; uint64_t stride; uint32_t n_frames;
    uint16_t* ring;              // n_slots planes of npx u16: loaded by the first frame segment, stored back by the last
    int n_slots;                 // 4 (dips: slots hold grey-quantised I2) or 2 (dips_alt: raw I2)
    int first_slot;              // slot the first frame overwrites; frame j overwrites (first_slot + j) % n_slots
    int median_is_max;           // dips_alt only, see RingArgs
    const uint16_t* start;       // start / snapshot plane, constant over the run
    uint32_t* acc_sum; uint32_t* acc_cnt;
    uint32_t* partials;          // u32[n_frames][pitch4(ring_clip_words_per_frame)]: sad | cnt<<20 per warp per frame
    uint32_t tau;
    uint32_t seg_frames = 0;     // frames per segment (0: planned from the occupancy); test hook
};
// 8-pixel units, aligned vector loads, 32-bit accumulator indices
bool ring_clip_available(const Geometry& g, const uint8_t* frames, uint64_t stride);
uint32_t ring_clip_words_per_frame(const Geometry& g);
cudaError_t launch_ring_clip(const Geometry& g, const RingClipArgs& a, cudaStream_t s);
// N4: correct spatial median (window 3/5/7, zero padded) of a u16 intensity plane; upper median of 4 planes
cudaError_t launch_spatial_median(const Geometry& g, const uint16_t* in, uint16_t* out, int window, cudaStream_t s);
cudaError_t launch_median4_planes(const Geometry& g, const uint16_t* planes, uint16_t* out, cudaStream_t s);
cudaError_t launch_passthrough_rgba(const Geometry& g, const uint8_t* frame, uint64_t pitch, int format,
                                    uint8_t* out_rgba, cudaStream_t s, uint64_t p_begin = 0, uint64_t p_end = 0);
cudaError_t launch_repack(const Geometry& g, const uint8_t* src, uint64_t stride, uint64_t fb, uint64_t n_frames, uint8_t* dst,
                          uint64_t dpitch, cudaStream_t s);
cudaError_t launch_synth(uint8_t* dst, uint64_t first_frame, uint64_t n_frames, uint32_t w, uint32_t h, int bpp,
                         uint64_t seed, int profile, cudaStream_t s);
// accumulator exchange format for the cross-GPU sum (see aux_kernels.cu)
cudaError_t launch_pack_acc(const Geometry& g, const uint32_t* acc, uint32_t* out, int layout, int sum_bits, cudaStream_t s);
cudaError_t launch_unpack_acc(const Geometry& g, const uint32_t* in, uint32_t* acc, int layout, int sum_bits, cudaStream_t s);
cudaError_t launch_intensity_map(const Geometry& g, const uint32_t* acc_internal, uint64_t n_eff, float* out,
                                 cudaStream_t s);

// comm.cu: one-thread kernel that waits (bounded) until *flag >= want
cudaError_t launch_wait_flag(const unsigned long long* flag, unsigned long long want, unsigned long long timeout_ns,
                             uint32_t* status, cudaStream_t s);

void count_launch(uint64_t n = 1);

// host_copy.cu: staging copy on a small persistent thread pool (DIPSB_COPY_THREADS, default min(4, cores/2); 1 = caller only)
void host_copy2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t row_bytes, size_t rows);
unsigned host_copy_threads();

}  // namespace dipsb
