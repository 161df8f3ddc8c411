// Micro-benchmark (developer tool): scalar FFMA / FADD vs packed fma.rn.f32x2 / add.rn.f32x2 on sm_100a.
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o fp2_peak fp2_peak.cu && ./fp2_peak
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
    uint64_t d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float seed) {
    if (MODE == 0) {            // scalar FFMA, 8 chains
        float a[8]; for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x + i;
        const float m = 0.999f, c = 0.001f;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int r = 0; r < 16; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], m, c);
        float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
        if (s == 123.456f) out[0] = s;
    } else if (MODE == 1) {     // packed FFMA2, 8 chains of 2
        uint64_t a[8]; for (int i = 0; i < 8; ++i) a[i] = (uint64_t(__float_as_uint(seed + i)) << 32) | __float_as_uint(seed + threadIdx.x);
        const uint64_t m = (uint64_t(__float_as_uint(0.999f)) << 32) | __float_as_uint(0.999f);
        const uint64_t c = (uint64_t(__float_as_uint(0.001f)) << 32) | __float_as_uint(0.001f);
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int r = 0; r < 16; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = ffma2(a[i], m, c);
        uint64_t s = 0; for (int i = 0; i < 8; ++i) s ^= a[i];
        if (s == 123456ull) out[0] = 1.f;
    } else if (MODE == 2) {     // scalar FADD, 8 chains, register operands
        float a[8]; for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x + i;
        float c = seed * 1e-3f;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int r = 0; r < 16; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = a[i] + c;
        float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
        if (s == 123.456f) out[0] = s;
    } else {                    // packed FADD2
        uint64_t a[8]; for (int i = 0; i < 8; ++i) a[i] = (uint64_t(__float_as_uint(seed + i)) << 32) | __float_as_uint(seed + threadIdx.x);
        const uint64_t c = (uint64_t(__float_as_uint(seed * 1e-3f)) << 32) | __float_as_uint(seed * 1e-3f);
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int r = 0; r < 16; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = fadd2(a[i], c);
        uint64_t s = 0; for (int i = 0; i < 8; ++i) s ^= a[i];
        if (s == 123456ull) out[0] = 1.f;
    }
}
template <int MODE> double run(const char* name, int lanes_per_instr, int flops_per_lane) {
    float* d; cudaMalloc(&d, 4);
    const int blocks = 148 * 8, threads = 256, inner = 4096, iters = 10;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(d, inner, 1.0f);
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) k<MODE><<<blocks, threads>>>(d, inner, 1.0f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double instr = 8.0 * 16 * inner * threads * blocks * iters;      // thread-instructions
    const double tf = instr * lanes_per_instr * flops_per_lane / (ms * 1e-3) / 1e12;
    const double ipc = instr / 32 / (ms * 1e-3) / (148.0 * 1.965e9);       // warp-instr per clk per SM at 1.965 GHz
    printf("%-12s %8.3f ms  %7.2f TFLOP/s  %5.2f warp-instr/clk/SM (at 1.965 GHz)\n", name, ms, tf, ipc);
    cudaFree(d); return tf;
}
int main() {
    run<0>("FFMA", 1, 2); run<1>("FFMA2", 2, 2); run<2>("FADD", 1, 1); run<3>("FADD2", 2, 1);
    return 0;
}
