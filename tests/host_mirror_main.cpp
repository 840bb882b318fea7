// Test driver for dips_b200/host/dips_host.hpp (compiled by tests/test_host_mirror.py with g++ -std=c++17).
// usage: host_mirror_main <dips|alt> <width> <height> <n_frames> <in.rgba> <out.rgba>
//   dips: every frame through dips::frame_callback (reference-exact flavour);
//   alt : every frame through dips_alt::DiPsCompute::send_frame with the reference's snapshot rule
//         (snapshot on the call where index == FRAME_COUNT, dips_alt/src/lib.rs:633-666).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>

#include "dips_host.hpp"

int main(int argc, char** argv) {
    if (argc == 2 && !std::strcmp(argv[1], "--compile-check")) { std::puts("ok"); return 0; }
    if (argc != 7) { std::fprintf(stderr, "usage\n"); return 2; }
    const std::string mode = argv[1];
    const uint32_t w = std::atoi(argv[2]), h = std::atoi(argv[3]);
    const int n = std::atoi(argv[4]);
    const size_t fb = static_cast<size_t>(w) * h * 4;
    std::vector<uint8_t> in(fb * n);
    std::ifstream fi(argv[5], std::ios::binary);
    fi.read(reinterpret_cast<char*>(in.data()), in.size());
    if (!fi) { std::fprintf(stderr, "short input\n"); return 3; }
    std::ofstream fo(argv[6], std::ios::binary);
    try {
        if (mode == "dips") {
            dips::DiPsProperties p;                       // builder defaults: grey, window 1, sensitivity 5, unfiltered
            dips::ComputeState cs(p.colorize, p.spatial_window_size, p.sensitivity, p.filter_type, p.chroma_filter);
            for (int t = 0; t < n; ++t) {
                auto out = dips::frame_callback(w, h, in.data() + fb * t, fb, cs);
                fo.write(reinterpret_cast<const char*>(out.data()), out.size());
            }
        } else {
            dips_alt::DiPsProperties p;                   // defaults: colourised sigmoid 5
            dips_alt::DiPsCompute dc(2, w, h, p);
            size_t index = 0;
            const size_t FRAME_COUNT = 2;
            for (int t = 0; t < n; ++t) {
                auto out = dc.send_frame(in.data() + fb * t, fb, index == FRAME_COUNT);
                if (index <= FRAME_COUNT) ++index;
                fo.write(reinterpret_cast<const char*>(out.data()), out.size());
            }
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 4;
    }
    return 0;
}
