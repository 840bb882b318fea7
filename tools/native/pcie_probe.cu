// pcie_probe.cu -- what does the platform give a per-frame caller over PCIe, independent of libdips_b200?  (measurement aid)
// Two page-locked frame buffers, two streams, bare cudaMemcpyAsync: one direction at a time, both directions at once, and
// the per-frame pattern of the pipelined boundary (upload of frame t next to the read-back of frame t-1, one
// synchronisation per frame).   nvcc -O3 -o build/pcie_probe tools/native/pcie_probe.cu && build/pcie_probe [frame MB]
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
using clk = std::chrono::steady_clock;
static double secs(clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); }

int main(int argc, char** argv) {
    const size_t bytes = (size_t)((argc > 1 ? atof(argv[1]) : 8.2944) * 1e6);   // 1920x1080 RGBA8
    const int reps = 400;
    void *h_in, *h_out, *d_in, *d_out;
    CK(cudaHostAlloc(&h_in, bytes, cudaHostAllocDefault)); CK(cudaHostAlloc(&h_out, bytes, cudaHostAllocDefault));
    CK(cudaMalloc(&d_in, bytes)); CK(cudaMalloc(&d_out, bytes));
    cudaStream_t up, down;
    CK(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&down, cudaStreamNonBlocking));
    auto run = [&](const char* name, bool h2d, bool d2h, bool sync_each) {
        for (int w = 0; w < 10; ++w) {
            if (h2d) CK(cudaMemcpyAsync(d_in, h_in, bytes, cudaMemcpyHostToDevice, up));
            if (d2h) CK(cudaMemcpyAsync(h_out, d_out, bytes, cudaMemcpyDeviceToHost, down));
        }
        CK(cudaDeviceSynchronize());
        const auto t0 = clk::now();
        for (int r = 0; r < reps; ++r) {
            if (d2h) CK(cudaMemcpyAsync(h_out, d_out, bytes, cudaMemcpyDeviceToHost, down));
            if (h2d) CK(cudaMemcpyAsync(d_in, h_in, bytes, cudaMemcpyHostToDevice, up));
            if (sync_each) { if (h2d) CK(cudaStreamSynchronize(up)); if (d2h) CK(cudaStreamSynchronize(down)); }
        }
        CK(cudaDeviceSynchronize());
        const double dt = secs(t0, clk::now());
        const double gb = (double)bytes * reps * ((h2d ? 1 : 0) + (d2h ? 1 : 0)) / 1e9;
        printf("%-58s %7.1f us per frame   %6.1f GB/s total%s\n", name, 1e6 * dt / reps, gb / dt,
               (h2d && d2h) ? "  (both directions summed)" : "");
    };
    printf("frame = %.2f MB, %d repetitions\n", bytes / 1e6, reps);
    run("H2D only, back to back", true, false, false);
    run("D2H only, back to back", false, true, false);
    run("H2D and D2H concurrently, back to back", true, true, false);
    run("H2D only, one synchronisation per frame", true, false, true);
    run("D2H only, one synchronisation per frame", false, true, true);
    run("H2D(t) next to D2H(t-1), one synchronisation per frame", true, true, true);
    return 0;
}
