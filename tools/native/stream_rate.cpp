// stream_rate.cpp -- calls per second of the reference-shaped per-frame boundary, measured in native code (no Python binding):
//   mirror      dips::frame_callback over dips::ComputeState (dips_b200/host/dips_host.hpp): the drop-in for
//               dips/src/lib.rs:233-246 -- add_texture + dispatch, a fresh std::vector out per frame, pageable buffers
//   push        dipsb_push_frame,            pageable / page-locked buffers
//   pipelined   dipsb_push_frame_pipelined,  pageable / page-locked buffers
// 1920x1080 RGBA8 in, RGBA8 difference frame out.  Prints one JSON object.
//   g++ -std=c++17 -O2 -Iinclude -Idips_b200/host tools/native/stream_rate.cpp -o build/stream_rate -Ldips_b200 -ldips_b200 -Wl,-rpath,$PWD/dips_b200
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "dips_host.hpp"

using clk = std::chrono::steady_clock;
static double secs(clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); }

int main(int argc, char** argv) {
    const uint32_t w = 1920, h = 1080;
    const int n = argc > 1 ? atoi(argv[1]) : 200, ring = 8;
    const size_t fb = (size_t)w * h * 4;
    std::vector<uint8_t> frames(ring * fb);
    uint32_t x = 0x44695073u;
    for (size_t i = 0; i < frames.size(); ++i) { x = x * 1664525u + 1013904223u; frames[i] = (uint8_t)(x >> 24); }
    printf("{");
    {   // the C++ mirror of the reference callback (north-star semantics: reference = first frame)
        dips::ComputeState cs(false, 1, 5.0f, dips::DiPsFilter::Unfiltered, dips::ChromaFilter::None, /*reference_exact=*/false);
        size_t sink = 0;
        for (int k = 0; k < 4; ++k) sink += dips::frame_callback(w, h, frames.data() + (k % ring) * fb, fb, cs)[17];
        const auto t0 = clk::now();
        for (int k = 0; k < n; ++k) sink += dips::frame_callback(w, h, frames.data() + (k % ring) * fb, fb, cs)[17];
        printf("\"mirror_frame_callback_fps\": %.1f, ", n / secs(t0, clk::now()));
        double ta = 0, td = 0;       // where a mirror call spends its time: add_texture (stage) vs dispatch (kernels, read-back, fresh vector)
        for (int k = 0; k < n; ++k) {
            const auto a = clk::now();
            cs.add_texture(w, h, frames.data() + (k % ring) * fb, fb);
            const auto b = clk::now();
            auto f = cs.dispatch();
            const auto c2 = clk::now();
            sink += f ? (*f)[17] : 0;
            ta += secs(a, b); td += secs(b, c2);
        }
        printf("\"mirror_add_texture_us\": %.1f, \"mirror_dispatch_us\": %.1f, ", 1e6 * ta / n, 1e6 * td / n);
        if (sink == 0xdeadbeef) printf("\"x\": 0, ");
    }
    dipsb_config cfg;
    dipsb_default_config(&cfg);
    cfg.width = w; cfg.height = h; cfg.format = DIPSB_FMT_RGBX8; cfg.threshold = 32;
    for (int pinned = 0; pinned < 2; ++pinned) {
        uint8_t *in = nullptr, *out = nullptr;
        if (pinned) {
            void* p = nullptr;
            if (dipsb_host_alloc(0, ring * fb, &p)) return 1;
            in = (uint8_t*)p;
            if (dipsb_host_alloc(0, 2 * fb, &p)) return 1;
            out = (uint8_t*)p;
            memcpy(in, frames.data(), ring * fb);
        } else {
            in = frames.data();
            out = (uint8_t*)malloc(2 * fb);
            memset(out, 0, 2 * fb);
        }
        for (int pipelined = 0; pipelined < 2; ++pipelined) {
            dipsb_ctx* c = nullptr;
            if (dipsb_create(&cfg, &c)) { fprintf(stderr, "%s\n", dipsb_last_error(nullptr)); return 1; }
            auto call = [&](int k) {
                return pipelined ? dipsb_push_frame_pipelined(c, in + (k % ring) * fb, w, h, w * 4, DIPSB_FMT_RGBX8, out + (k & 1) * fb, nullptr)
                                 : dipsb_push_frame(c, in + (k % ring) * fb, w, h, w * 4, DIPSB_FMT_RGBX8, out + (k & 1) * fb, nullptr);
            };
            for (int k = 0; k < 4; ++k) if (call(k) < 0) { fprintf(stderr, "%s\n", dipsb_last_error(c)); return 1; }
            const auto t0 = clk::now();
            for (int k = 4; k < n + 4; ++k) if (call(k) < 0) { fprintf(stderr, "%s\n", dipsb_last_error(c)); return 1; }
            if (pipelined) dipsb_flush_frame(c, out, nullptr);
            printf("\"%s_%s_fps\": %.1f%s", pipelined ? "push_frame_pipelined" : "push_frame", pinned ? "pinned" : "pageable",
                   n / secs(t0, clk::now()), (pinned && pipelined) ? "" : ", ");
            dipsb_destroy(c);
        }
        if (pinned) { dipsb_host_free(in); dipsb_host_free(out); } else free(out);
    }
    printf("}\n");
    return 0;
}
