// dips_host.hpp -- C++ host-side mirror of the reference's per-frame API over the C ABI (include/dips_b200.h).
//
// The reference's host side is Rust; this image has no Rust toolchain, so besides the Rust crates shipped as source (rust/)
// the same interface is provided in C++ (header-only) where it can be compiled and tested:
//   dips::ComputeState {ComputeState(...), add_texture, dispatch}   == dips/src/gpu/mod.rs:59, :170, :306
//   dips::frame_callback                                            == dips/src/lib.rs:233-246
//   dips::DiPsProperties / DiPsFilter / ChromaFilter                == dips/src/lib.rs:26-170
//   dips_alt::DiPsCompute {DiPsCompute(...), send_frame}            == dips_alt/src/dips_compute/mod.rs:270, :498
// Same names, argument meaning and error behaviour: constructors throw std::runtime_error where the reference returns
// anyhow::Err, dispatch() returns an empty optional while the reference returns None (warm-up), send_frame throws where the
// reference panics.
#pragma once
#include <cstdint>
#include <memory>
#include <new>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "dips_b200.h"

namespace dips {

// The frame type both mirrors hand back: a std::vector of bytes whose allocator does not zero what the library is about to
// overwrite (a value-initialising vector costs an 8 MB memset per 1080p frame -- the reference's `vec![0; n]` gets its zero
// pages from calloc for free).
template <class T>
struct default_init_allocator : std::allocator<T> {
    template <class U> struct rebind { using other = default_init_allocator<U>; };
    using std::allocator<T>::allocator;
    template <class U> void construct(U* p) noexcept { ::new (static_cast<void*>(p)) U; }
    template <class U, class... A> void construct(U* p, A&&... a) { ::new (static_cast<void*>(p)) U(std::forward<A>(a)...); }
};
using Frame = std::vector<uint8_t, default_init_allocator<uint8_t>>;

enum class DiPsFilter { Unfiltered, Sigmoid, InverseSigmoid };   // dips/src/lib.rs:26-41
enum class ChromaFilter { None, Red, Green, Blue };              // dips/src/lib.rs:44-61

inline int32_t to_ffi(DiPsFilter f) {
    switch (f) {
        case DiPsFilter::Sigmoid: return DIPSB_FILTER_SIGMOID;
        case DiPsFilter::InverseSigmoid: return DIPSB_FILTER_INV_SIGMOID;
        default: return DIPSB_FILTER_NONE;
    }
}
inline int32_t to_ffi(ChromaFilter c) { return static_cast<int32_t>(c); }

// builder defaults of DiPsProperties::new(), dips/src/lib.rs:75-86
struct DiPsProperties {
    bool colorize = false;
    int32_t spatial_window_size = 1;
    float sensitivity = 5.0f;
    DiPsFilter filter_type = DiPsFilter::Unfiltered;
    ChromaFilter chroma_filter = ChromaFilter::None;
};

class ComputeState {
  public:
    // reference_exact = true: the crate's own temporal semantics (median-of-4 ring, 3 passthrough frames);
    // false: north-star semantics (reference = first frame, 1 passthrough frame)
    ComputeState(bool colorize, int32_t spatial_window_size, float sensitivity, DiPsFilter filter_type,
                 ChromaFilter chroma_filter, bool reference_exact = true, int32_t device = 0)
        : colorize_(colorize), window_(spatial_window_size), sensitivity_(sensitivity), filter_(filter_type),
          chroma_(chroma_filter), exact_(reference_exact), device_(device) {
        if (spatial_window_size != 1 && spatial_window_size != 3 && spatial_window_size != 5 && spatial_window_size != 7)
            throw std::runtime_error("spatial_window_size must be 1, 3, 5 or 7");   // > 1: correct zero-padded median
    }
    ComputeState(const ComputeState&) = delete;
    ComputeState& operator=(const ComputeState&) = delete;
    ~ComputeState() { if (ctx_) dipsb_destroy(ctx_); }

    // dips/src/gpu/mod.rs:170 -- frame_data is borrowed for the call only
    void add_texture(uint32_t width, uint32_t height, const uint8_t* frame_data, size_t len) {
        if (len < static_cast<size_t>(width) * height * 4) throw std::runtime_error("add_texture: frame smaller than width*height*4");
        ensure(width, height);
        // no copy of our own (the reference's `to_vec`, mod.rs:171): the frame goes straight into the library's page-locked
        // input slot and its upload starts; dispatch() finishes the call
        if (dipsb_stage_frame(ctx_, frame_data, width, height, width * 4, DIPSB_FMT_RGBX8) != DIPSB_OK)
            throw std::runtime_error(std::string("dipsb_stage_frame: ") + dipsb_last_error(ctx_));
        have_frame_ = true;
    }

    // dips/src/gpu/mod.rs:306 -- empty while warming up
    std::optional<Frame> dispatch() {
        if (!have_frame_ || !ctx_) return std::nullopt;
        have_frame_ = false;
        Frame out(static_cast<size_t>(width_) * height_ * 4);
        const int32_t rc = dipsb_dispatch_staged(ctx_, out.data(), &stats_);
        if (rc < 0) throw std::runtime_error(std::string("dipsb_dispatch_staged: ") + dipsb_last_error(ctx_));
        if (rc == DIPSB_NOT_READY) return std::nullopt;
        return out;
    }

    const dipsb_frame_stats& last_stats() const { return stats_; }
    dipsb_ctx* raw() { return ctx_; }

  private:
    void ensure(uint32_t w, uint32_t h) {
        if (ctx_ && w == width_ && h == height_) return;
        if (ctx_) { dipsb_destroy(ctx_); ctx_ = nullptr; }
        dipsb_config cfg;
        dipsb_default_config(&cfg);
        cfg.device = device_; cfg.width = w; cfg.height = h; cfg.format = DIPSB_FMT_RGBX8; cfg.mode = DIPSB_MODE_OVERALL;
        cfg.chroma = to_ffi(chroma_); cfg.colorize = colorize_; cfg.filter = to_ffi(filter_);
        cfg.sigmoid_scalar = sensitivity_; cfg.spatial_window = window_;
        cfg.flavor = exact_ ? DIPSB_FLAVOR_DIPS_RING4 : DIPSB_FLAVOR_FRAME0;
        if (dipsb_create(&cfg, &ctx_) != DIPSB_OK) throw std::runtime_error(std::string("dipsb_create: ") + dipsb_last_error(nullptr));
        width_ = w; height_ = h;
    }
    bool colorize_; int32_t window_; float sensitivity_; DiPsFilter filter_; ChromaFilter chroma_; bool exact_; int32_t device_;
    dipsb_ctx* ctx_ = nullptr;
    uint32_t width_ = 0, height_ = 0;
    bool have_frame_ = false;
    dipsb_frame_stats stats_{};
};

// dips/src/lib.rs:233-246, verbatim shape
inline Frame frame_callback(uint32_t width, uint32_t height, const uint8_t* frame_data, size_t len, ComputeState& compute) {
    compute.add_texture(width, height, frame_data, len);
    if (auto new_frame = compute.dispatch()) return std::move(*new_frame);   // moved, as Rust moves the Vec out of the Option
    return Frame(frame_data, frame_data + len);
}

// Page-locked frame buffer (dipsb_host_alloc): a decoder that writes here, and a caller that receives the difference frame
// here, let dipsb_push_frame* skip its staging copies (the role of the mapped gst buffer, dips/src/frame_extractor.rs:216-226).
class PinnedFrame {
  public:
    PinnedFrame(size_t len, int32_t device = 0) : len_(len) {
        void* p = nullptr;
        if (dipsb_host_alloc(device, len, &p) != DIPSB_OK) throw std::runtime_error(dipsb_last_error(nullptr));
        ptr_ = static_cast<uint8_t*>(p);
    }
    ~PinnedFrame() { dipsb_host_free(ptr_); }
    PinnedFrame(const PinnedFrame&) = delete;
    PinnedFrame& operator=(const PinnedFrame&) = delete;
    uint8_t* data() { return ptr_; }
    const uint8_t* data() const { return ptr_; }
    size_t size() const { return len_; }

  private:
    uint8_t* ptr_ = nullptr;
    size_t len_ = 0;
};

// The same for a buffer the caller already owns and reuses (a decoder's buffer pool, the vector the output is collected in):
// page-locked in place for the lifetime of this object (dipsb_host_register / dipsb_host_unregister).  The memory must
// outlive it.
class RegisteredFrames {
  public:
    RegisteredFrames(void* p, size_t len, int32_t device = 0) : ptr_(p) {
        if (dipsb_host_register(device, p, len) != DIPSB_OK) throw std::runtime_error(dipsb_last_error(nullptr));
    }
    ~RegisteredFrames() { dipsb_host_unregister(ptr_); }
    RegisteredFrames(const RegisteredFrames&) = delete;
    RegisteredFrames& operator=(const RegisteredFrames&) = delete;

  private:
    void* ptr_ = nullptr;
};

}  // namespace dips

namespace dips_alt {

enum class Filter { Sigmoid = 0, InverseSigmoid = 1 };           // dips_alt/src/dips_compute/mod.rs:151-157
enum class ChromaFilter { All = 0, Red = 1, Green = 2, Blue = 3 };

struct DiPsProperties {                                           // defaults of :179-189, clamps of :218-229
    bool colorize = true;
    uint8_t window_size = 1;
    float sigmoid_horizontal_scalar = 5.0f;
    Filter filter_type = Filter::Sigmoid;
    ChromaFilter chroma_filter = ChromaFilter::All;
    void set_sigmoid_horizontal_scalar(float s) { sigmoid_horizontal_scalar = s < 1.f ? 1.f : (s > 10.f ? 10.f : s); }
    void set_window_size(uint8_t s) { window_size = s < 1 ? 1 : (s > 7 ? 7 : s); if (window_size % 2 == 0) --window_size; }
};

class DiPsCompute {
  public:
    // num_textures is fixed to the reference's FRAME_COUNT = 2 (dips_alt/src/lib.rs:36); the wgpu window/device/queue
    // arguments of the reference constructor have no meaning on the CUDA path
    DiPsCompute(size_t num_textures, uint32_t textures_width, uint32_t textures_height, const DiPsProperties& p,
                bool as_shipped_median = true, int32_t device = 0)
        : width_(textures_width), height_(textures_height) {
        if (num_textures != 2) throw std::runtime_error("DiPsCompute: only num_textures == 2 (FRAME_COUNT) is implemented");

        dipsb_config cfg;
        dipsb_default_config(&cfg);
        cfg.device = device; cfg.width = width_; cfg.height = height_; cfg.format = DIPSB_FMT_RGBX8; cfg.mode = DIPSB_MODE_OVERALL;
        cfg.chroma = static_cast<int32_t>(p.chroma_filter); cfg.colorize = p.colorize; cfg.filter = static_cast<int32_t>(p.filter_type);
        cfg.sigmoid_scalar = p.sigmoid_horizontal_scalar; cfg.spatial_window = p.window_size;   // 1/3/5/7 after set_window_size
        cfg.flavor = as_shipped_median ? DIPSB_FLAVOR_ALT_RING2 : DIPSB_FLAVOR_ALT_RING2_MEDIAN;
        if (dipsb_create(&cfg, &ctx_) != DIPSB_OK) throw std::runtime_error(std::string("dipsb_create: ") + dipsb_last_error(nullptr));
    }
    DiPsCompute(const DiPsCompute&) = delete;
    DiPsCompute& operator=(const DiPsCompute&) = delete;
    ~DiPsCompute() { if (ctx_) dipsb_destroy(ctx_); }

    // dips_alt/src/dips_compute/mod.rs:498-503; snapshot == true is `Some(())`
    dips::Frame send_frame(const uint8_t* frame, size_t len, bool snapshot) {
        if (len < static_cast<size_t>(width_) * height_ * 4) throw std::runtime_error("send_frame: frame smaller than width*height*4");
        if (snapshot) dipsb_snapshot(ctx_);
        dips::Frame out(static_cast<size_t>(width_) * height_ * 4);
        if (dipsb_push_frame(ctx_, frame, width_, height_, width_ * 4, DIPSB_FMT_RGBX8, out.data(), nullptr) < 0)
            throw std::runtime_error(std::string("dipsb_push_frame: ") + dipsb_last_error(ctx_));
        return out;
    }

  private:
    dipsb_ctx* ctx_ = nullptr;
    uint32_t width_, height_;
};

constexpr size_t FRAME_COUNT = 2;                                 // dips_alt/src/lib.rs:36

// When the caller loops of dips_alt ask for a snapshot: the frame on which `index == FRAME_COUNT` (the third frame, and the
// third frame after every refresh marker); `index` saturates one past it and a marker -- a 1-based count of frames
// processed so far -- resets it (dips_alt/src/lib.rs:222-232 live, :560-561 and :662-670 file mode).
class SnapshotSchedule {
  public:
    explicit SnapshotSchedule(std::vector<size_t> refresh_markers = {}) : markers_(std::move(refresh_markers)) {}
    bool snapshot_now() const { return index_ == FRAME_COUNT; }   // ask BEFORE sending the frame
    void frame_sent() {                                           // call AFTER sending it
        if (index_ <= FRAME_COUNT) ++index_;
        ++overall_frame_;
        for (size_t m : markers_) if (m == overall_frame_) { index_ = 0; break; }
    }
    size_t overall_frame() const { return overall_frame_; }

  private:
    std::vector<size_t> markers_;
    size_t index_ = 0, overall_frame_ = 0;
};

// The compute part of run_dips_on_file (dips_alt/src/lib.rs:553-690) without the OpenCV capture/writer around it:
// every RGBA frame goes through send_frame with that schedule; `sink(frame_number, rgba)` receives the difference frames.
template <class Sink>
inline void run_dips_on_frames(DiPsCompute& compute, const uint8_t* frames, size_t n_frames, size_t frame_len,
                               const std::vector<size_t>& refresh_markers, Sink&& sink) {
    SnapshotSchedule schedule(refresh_markers);
    for (size_t t = 0; t < n_frames; ++t) {
        dips::Frame out = compute.send_frame(frames + t * frame_len, frame_len, schedule.snapshot_now());
        schedule.frame_sent();
        sink(t, out);
    }
}

}  // namespace dips_alt
