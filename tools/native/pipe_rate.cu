// pipe_rate.cu -- issue rate of the few instructions the packed-u16 kernels of libdips_b200 are made of, per SM sub-partition:
// 8 independent dependency chains per thread, 8 warps per sub-partition, clock64 around the loop.  Measurement aid for the
// median / clip kernels (which pipe is the scarce one).   nvcc -arch=sm_100a -O3 -o build/pipe_rate pipe_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define OPS 8
#define ITER 4096

template <int OP>
__device__ __forceinline__ uint32_t apply(uint32_t x, uint32_t a, uint32_t b) {
    if constexpr (OP == 0) return __viaddmin_s16x2_relu(x, a, b);            // VIADDMNMX.S16x2.RELU
    else if constexpr (OP == 1) return __vmaxu2(x, a);                        // VIMNMX.U16x2
    else if constexpr (OP == 2) return __vimax3_u16x2(x, a, b);               // VIMNMX3.U16x2
    else if constexpr (OP == 3) { uint32_t r; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(a), "r"(b)); return r; }   // IMAD
    else if constexpr (OP == 4) { uint32_t r; asm volatile("add.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(a)); return r; }                  // IADD3
    else if constexpr (OP == 5) { uint32_t r; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(x), "r"(a), "r"(b)); return r; }   // LOP3
    else if constexpr (OP == 6) { uint32_t r; asm volatile("fma.rn.sat.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(a), "r"(b)); return r; }  // HFMA2.SAT
    else if constexpr (OP == 7) { uint32_t r; asm volatile("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(a)); return r; }             // HADD2
    else if constexpr (OP == 8) return __byte_perm(x, a, b);                  // PRMT
    else if constexpr (OP == 9) return __dp2a_lo(x, a, b);                    // IDP.2A
    else if constexpr (OP == 10) { uint32_t r; asm volatile("shf.r.clamp.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(a), "r"(b & 31)); return r; }  // SHF
    else { uint32_t r; asm volatile("set.ge.u32.f16x2 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(a)); return r; }                                // HSET2
}

template <int OP>
__global__ void rate_kernel(uint32_t a, uint32_t b, uint32_t* sink, long long* cycles) {
    uint32_t x[OPS];
#pragma unroll
    for (int i = 0; i < OPS; ++i) x[i] = threadIdx.x * 7u + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < OPS; ++i) x[i] = apply<OP>(x[i], a, b);
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < OPS; ++i) s ^= x[i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// two instruction kinds interleaved 1:1 (do the pipes overlap?)
template <int OPA, int OPB>
__global__ void pair_kernel(uint32_t a, uint32_t b, uint32_t* sink, long long* cycles) {
    uint32_t x[OPS];
#pragma unroll
    for (int i = 0; i < OPS; ++i) x[i] = threadIdx.x * 7u + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < OPS; i += 2) { x[i] = apply<OPA>(x[i], a, b); x[i + 1] = apply<OPB>(x[i + 1], a, b); }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < OPS; ++i) s ^= x[i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <typename K>
static void run(const char* name, K kernel, int threads) {
    uint32_t* sink; long long* cyc;
    cudaMalloc(&sink, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    kernel<<<148, threads>>>(3u, 0x00010001u, sink, cyc);
    kernel<<<148, threads>>>(3u, 0x00010001u, sink, cyc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)h[i] / 148;
    const double warp_instr_per_smsp = (double)ITER * OPS * (threads / 32) / 4.0;
    printf("%-34s %2d warps/SMSP: %6.3f clk per warp-instruction per sub-partition\n", name, threads / 128, avg / warp_instr_per_smsp);
    cudaFree(sink); cudaFree(cyc);
}

int main() {
    const int T = 1024;
    run("VIADDMNMX.S16x2.RELU", rate_kernel<0>, T);
    run("VIMNMX.U16x2", rate_kernel<1>, T);
    run("VIMNMX3.U16x2", rate_kernel<2>, T);
    run("IMAD", rate_kernel<3>, T);
    run("IADD3", rate_kernel<4>, T);
    run("LOP3", rate_kernel<5>, T);
    run("HFMA2.SAT", rate_kernel<6>, T);
    run("HADD2", rate_kernel<7>, T);
    run("PRMT", rate_kernel<8>, T);
    run("IDP.2A", rate_kernel<9>, T);
    run("SHF", rate_kernel<10>, T);
    run("HSET2 (set.ge.u32.f16x2)", rate_kernel<11>, T);
    run("VIADDMNMX + IMAD", pair_kernel<0, 3>, T);
    run("VIADDMNMX + IADD3", pair_kernel<0, 4>, T);
    run("VIADDMNMX + HFMA2.SAT", pair_kernel<0, 6>, T);
    run("HFMA2.SAT + IADD3", pair_kernel<6, 4>, T);
    run("HFMA2.SAT + IMAD", pair_kernel<6, 3>, T);
    run("VIMNMX + IMAD", pair_kernel<1, 3>, T);
    run("VIMNMX + VIADDMNMX", pair_kernel<1, 0>, T);
    run("VIADDMNMX + HSET2", pair_kernel<0, 11>, T);
    run("HSET2 + IADD3", pair_kernel<11, 4>, T);
    run("HSET2 + HFMA2.SAT", pair_kernel<11, 6>, T);
    run("VIADDMNMX + LOP3", pair_kernel<0, 5>, T);
    run("VIMNMX + LOP3", pair_kernel<1, 5>, T);
    run("VIMNMX3 + PRMT", pair_kernel<2, 8>, T);
    run("VIMNMX3 + IMAD", pair_kernel<2, 3>, T);
    run("PRMT + IMAD", pair_kernel<8, 3>, T);
    run("LOP3 + IMAD", pair_kernel<5, 3>, T);
    run("IADD3 + LOP3", pair_kernel<4, 5>, T);
    return 0;
}
