// Micro-benchmark (developer tool): issue rate of LONG STRAIGHT-LINE code when many warps of an SM run through it
// out of phase -- the execution pattern of the fused PIV kernels (a few thousand unrolled instructions per job,
// every instruction executed once per job per warp).  Body = U x 16 instructions (8 independent chains, each one
// FFMA + one IADD3 per step: issue-bound mix, limit ~3.8 warp-instr/clk/SM), repeated in a loop; the warps start
// with a skew so that they sit at different places of the body.
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o icache_stream icache_stream.cu && ./icache_stream
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int U>
__global__ void __launch_bounds__(1024, 1) k(float* out, int iters, int skew_ns) {
    float a[8]; int n[8];
    for (int i = 0; i < 8; ++i) { a[i] = 0.5f + threadIdx.x + i; n[i] = threadIdx.x + i; }
    if (skew_ns) __nanosleep(skew_ns * (threadIdx.x >> 5));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                a[i] = fmaf(a[i], 0.999f, a[(i + 3) & 7]);
                n[i] = n[i] + n[(i + 5) & 7] + u;           // distinct immediates: the unrolled body cannot be re-rolled
            }
        }
    }
    float s = 0; int z = 0;
    for (int i = 0; i < 8; ++i) { s += a[i]; z ^= n[i]; }
    if (s == 123.456f || z == 0x7fffffff) out[0] = s;
}
template <int U> void run(int warps, int skew_ns) {
    float* d; cudaMalloc(&d, 4);
    const int total = 1 << 21;                  // instructions per warp
    const int iters = total / (16 * U);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<U><<<148, warps * 32>>>(d, 4, skew_ns);
    cudaEventRecord(e0);
    k<U><<<148, warps * 32>>>(d, iters, skew_ns);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double instr = 16.0 * U * iters * warps;      // warp-instructions per SM
    printf("body %5d instr (%4d KB)  warps/SM %2d  skew %5d ns  %8.3f ms  %5.2f warp-instr/clk/SM\n", 16 * U, 16 * U * 16 / 1024,
           warps, skew_ns, ms, instr / (ms * 1e-3) / 1.965e9);
    cudaFree(d);
}
int main() {
    for (int warps : {12, 20}) for (int skew : {0, 3000}) {
        run<8>(warps, skew); run<128>(warps, skew); run<144>(warps, skew); run<160>(warps, skew); run<176>(warps, skew);
        run<192>(warps, skew); run<224>(warps, skew); run<256>(warps, skew); run<512>(warps, skew);
    }
    return 0;
}
