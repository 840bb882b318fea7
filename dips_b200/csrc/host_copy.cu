// Host-side staging copies, spread over a few threads.
//
// The reference hands every frame to its callback as a borrowed slice of ordinary (pageable) memory
// (dips/src/frame_extractor.rs:216-226), so the library has to copy it into a page-locked staging buffer before the copy
// engine can take it, and copy the difference frame back out the same way.  One core moves ~10 GB/s; at 1080p RGBA that
// is 0.8 ms per direction -- five times the PCIe transfer.  A small persistent pool (the caller + a few helpers) brings the
// copy close to what the memory system gives.  Pure host code: nothing here touches the device or computes anything.
#include "dipsb_internal.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace dipsb {
namespace {

struct CopyJob {
    uint8_t* dst = nullptr; const uint8_t* src = nullptr;
    size_t dpitch = 0, spitch = 0, row_bytes = 0, rows = 0;
    size_t parts = 0;
};

// bytes [b0, b1) of the logical stream rows*row_bytes
void copy_range(const CopyJob& j, size_t b0, size_t b1) {
    while (b0 < b1) {
        const size_t r = b0 / j.row_bytes, off = b0 - r * j.row_bytes;
        const size_t len = std::min(j.row_bytes - off, b1 - b0);
        memcpy(j.dst + r * j.dpitch + off, j.src + r * j.spitch + off, len);
        b0 += len;
    }
}

void copy_part(const CopyJob& j, size_t i) {
    const size_t total = j.rows * j.row_bytes;
    // 4 KB boundaries so that two threads never share a page or cache line of the destination
    size_t b0 = (total / j.parts * i) & ~size_t(4095), b1 = (i + 1 == j.parts) ? total : (total / j.parts * (i + 1)) & ~size_t(4095);
    copy_range(j, b0, b1);
}

class CopyPool {
  public:
    CopyPool(unsigned helpers, long spin_us) : spin_us_(spin_us) {
        for (unsigned k = 0; k < helpers; ++k) workers_.emplace_back([this] { run(); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    unsigned threads() const { return (unsigned)workers_.size() + 1; }

    void copy(const CopyJob& job) {
        std::lock_guard<std::mutex> one_caller(callers_);
        // The helpers only stay hot between copies while copies keep coming at a per-frame pace: a gap of more than a few
        // spin windows since the previous copy means no streaming caller is active, and they go to sleep at once.
        const auto now = std::chrono::steady_clock::now();
        hot_.store(spin_us_ > 0 && now - last_copy_ < std::chrono::microseconds(8 * spin_us_), std::memory_order_relaxed);
        last_copy_ = now;
        std::unique_lock<std::mutex> lk(m_);
        job_ = job; next_ = 0; done_ = 0;
        epoch_.fetch_add(1, std::memory_order_release);
        lk.unlock();
        cv_.notify_all();
        lk.lock();
        while (next_ < job_.parts) {                     // the caller takes parts too
            const size_t i = next_++;
            lk.unlock();
            copy_part(job, i);
            lk.lock();
            ++done_;
        }
        cv_done_.wait(lk, [this] { return done_ == job_.parts; });
        job_.parts = 0; next_ = 0; done_ = 0;
    }

  private:
    void run() {
        std::unique_lock<std::mutex> lk(m_);
        for (;;) {
            if (!stop_ && next_ >= job_.parts && spin_us_ > 0 && hot_.load(std::memory_order_relaxed)) {
                // Stay hot for a moment: a per-frame caller comes back within a few hundred microseconds (upload, kernel,
                // read-back), and a helper that went to sleep in between wakes too late to be of any use.
                const uint64_t seen = epoch_.load(std::memory_order_relaxed);
                lk.unlock();
                const auto t0 = std::chrono::steady_clock::now();
                while (epoch_.load(std::memory_order_acquire) == seen) {
                    for (int k = 0; k < 32; ++k) cpu_relax();
                    if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(spin_us_)) break;
                }
                lk.lock();
            }
            cv_.wait(lk, [this] { return stop_ || next_ < job_.parts; });
            if (stop_) return;
            const size_t i = next_++;
            const CopyJob j = job_;
            lk.unlock();
            copy_part(j, i);
            lk.lock();
            if (++done_ == job_.parts) cv_done_.notify_all();
        }
    }

    static void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#elif defined(__aarch64__)
        asm volatile("yield");
#endif
    }

    const long spin_us_;
    std::atomic<bool> hot_{false};
    std::chrono::steady_clock::time_point last_copy_{};
    std::atomic<uint64_t> epoch_{0};
    std::mutex m_, callers_;
    std::condition_variable cv_, cv_done_;
    std::vector<std::thread> workers_;
    CopyJob job_;
    size_t next_ = 0, done_ = 0;
    bool stop_ = false;
};

long configured_spin_us() {
    if (const char* e = getenv("DIPSB_COPY_SPIN_US")) {
        const long v = strtol(e, nullptr, 10);
        if (v >= 0 && v <= 100000) return v;
    }
    return 500;
}

unsigned configured_threads() {
    if (const char* e = getenv("DIPSB_COPY_THREADS")) {
        const long v = strtol(e, nullptr, 10);
        if (v >= 1 && v <= 64) return (unsigned)v;
    }
    const unsigned hw = std::thread::hardware_concurrency();
    return std::max(1u, std::min(4u, hw / 2));
}

CopyPool* pool() {
    static CopyPool* p = [] {
        const unsigned t = configured_threads();
        return t > 1 ? new CopyPool(t - 1, configured_spin_us()) : nullptr;    // lives for the process: helper threads never outlive their pool
    }();
    return p;
}

}  // namespace

unsigned host_copy_threads() {
    CopyPool* p = pool();
    return p ? p->threads() : 1;
}

void host_copy2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t row_bytes, size_t rows) {
    if (!rows || !row_bytes) return;
    CopyJob j;
    j.dst = (uint8_t*)dst; j.src = (const uint8_t*)src;
    j.dpitch = dpitch; j.spitch = spitch; j.row_bytes = row_bytes; j.rows = rows;
    if (dpitch == row_bytes && spitch == row_bytes) {   // contiguous: one long row
        j.row_bytes = row_bytes * rows; j.rows = 1; j.dpitch = j.spitch = j.row_bytes;
    }
    const size_t total = j.rows * j.row_bytes;
    CopyPool* p = total >= (size_t(1) << 20) ? pool() : nullptr;
    if (!p) {
        copy_range(j, 0, total);
        return;
    }
    // at least 256 KB per part; a few parts per thread even out a late-waking helper
    j.parts = std::max<size_t>(1, std::min<size_t>(size_t(p->threads()) * 2, total >> 18));
    p->copy(j);
}

}  // namespace dipsb
