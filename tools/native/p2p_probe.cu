// p2p_probe.cu -- what does one B200 get out of NVLink to a peer, by copy engine and by SM kernels?  (measurement aid)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/p2p_probe tools/native/p2p_probe.cu && build/p2p_probe [MB]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int U>
__global__ void push_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n) {   // local loads, remote stores
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += U * stride) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) if (i + u * stride < n) v[u] = src[i + u * stride];
#pragma unroll
        for (int u = 0; u < U; ++u) if (i + u * stride < n) dst[i + u * stride] = v[u];
    }
}

int main(int argc, char** argv) {
    const size_t mb = argc > 1 ? atoi(argv[1]) : 128, bytes = mb << 20, n = bytes / 16;
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev < 2) { printf("needs 2 GPUs\n"); return 0; }
    uint4 *a0, *b0, *a1;
    CK(cudaSetDevice(1)); CK(cudaMalloc(&a1, bytes)); CK(cudaMemset(a1, 1, bytes));
    CK(cudaDeviceEnablePeerAccess(0, 0));
    CK(cudaSetDevice(0)); CK(cudaMalloc(&a0, bytes)); CK(cudaMalloc(&b0, bytes)); CK(cudaMemset(a0, 2, bytes));
    CK(cudaDeviceEnablePeerAccess(1, 0));
    cudaStream_t s; CK(cudaStreamCreate(&s));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto time = [&](const char* name, auto&& fn) {
        fn(); CK(cudaStreamSynchronize(s));
        CK(cudaEventRecord(e0, s));
        for (int r = 0; r < 10; ++r) fn();
        CK(cudaEventRecord(e1, s)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("%-52s %8.1f GB/s  (%.3f ms per %zu MB)\n", name, bytes * 10 / (ms * 1e6), ms / 10, mb);
    };
    time("copy engine, 0 -> 1 (cudaMemcpyAsync)", [&] { CK(cudaMemcpyAsync(a1, a0, bytes, cudaMemcpyDeviceToDevice, s)); });
    time("copy engine, 1 -> 0 (pull)", [&] { CK(cudaMemcpyAsync(b0, a1, bytes, cudaMemcpyDeviceToDevice, s)); });
    time("copy engine, local 0 -> 0", [&] { CK(cudaMemcpyAsync(b0, a0, bytes, cudaMemcpyDeviceToDevice, s)); });
    for (int blocks : {148, 296, 592, 1184, 2368}) {
        char name[96];
        snprintf(name, sizeof name, "SM push 0 -> 1, %d x 256 threads, 1 x 16 B", blocks);
        time(name, [&] { push_kernel<1><<<blocks, 256, 0, s>>>(a0, a1, n); });
        snprintf(name, sizeof name, "SM push 0 -> 1, %d x 256 threads, 4 x 16 B", blocks);
        time(name, [&] { push_kernel<4><<<blocks, 256, 0, s>>>(a0, a1, n); });
    }
    time("SM pull 1 -> 0, 592 x 256 threads, 4 x 16 B", [&] { push_kernel<4><<<592, 256, 0, s>>>(a1, b0, n); });
    time("SM pull 1 -> 0, 2368 x 256 threads, 4 x 16 B", [&] { push_kernel<4><<<2368, 256, 0, s>>>(a1, b0, n); });
    time("SM local 0 -> 0, 592 x 256 threads, 4 x 16 B", [&] { push_kernel<4><<<592, 256, 0, s>>>(a0, b0, n); });
    // small transfers: latency of one copy-engine copy vs one kernel
    for (size_t kb : {64, 1024, 4096, 16384}) {
        const size_t nb = kb << 10;
        char name[96];
        snprintf(name, sizeof name, "copy engine 0 -> 1, %zu KB", kb);
        fn_small: ;
        cudaEventRecord(e0, s);
        for (int r = 0; r < 20; ++r) CK(cudaMemcpyAsync(a1, a0, nb, cudaMemcpyDeviceToDevice, s));
        cudaEventRecord(e1, s); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-52s %8.1f GB/s  (%.1f us each)\n", name, nb * 20 / (ms * 1e6), ms * 1000 / 20);
        snprintf(name, sizeof name, "SM push 0 -> 1 (592 blocks, x4), %zu KB", kb);
        cudaEventRecord(e0, s);
        for (int r = 0; r < 20; ++r) push_kernel<4><<<592, 256, 0, s>>>(a0, a1, nb / 16);
        cudaEventRecord(e1, s); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("%-52s %8.1f GB/s  (%.1f us each)\n", name, nb * 20 / (ms * 1e6), ms * 1000 / 20);
    }
    return 0;
}
