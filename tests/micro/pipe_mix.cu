// Micro-benchmark (developer tool): issue / pipe throughput of the instruction kinds the fused PIV kernel is
// made of, alone and mixed, at the kernel's occupancy (2-4 warps per scheduler).
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_mix pipe_mix.cu && ./pipe_mix
// Prints warp-instructions per clock per SM (4 schedulers -> 4.0 is the issue limit).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float fadd(float a, float b) { float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ int iadd(int a, int b) { int d; asm volatile("add.s32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ int lop(int a, int b) { int d; asm volatile("xor.b32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ float i2f(int a) { float d; asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(d) : "r"(a)); return d; }
__device__ __forceinline__ float2 lds64(const void* p) { float2 v; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"((unsigned)__cvta_generic_to_shared(p))); return v; }
__device__ __forceinline__ void sts64(void* p, float x, float y) { asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"((unsigned)__cvta_generic_to_shared(p)), "f"(x), "f"(y) : "memory"); }
__device__ __forceinline__ u64 pk(float x, float y) { return (u64(__float_as_uint(y)) << 32) | __float_as_uint(x); }

enum { FFMA, FADD, FFMA2, FADD2, FMUL2, MIX_FADD2_FADD, MIX_FADD2_IADD, MIX_FADD2_2IADD, MIX_FADD_IADD, MIX_FFT, I2F, LDS64, STS64, MIX_FADD2_LDS, NMODES };
static const char* NAMES[] = {"FFMA", "FADD", "FFMA2", "FADD2", "FMUL2", "FADD2+FADD 1:1", "FADD2+IADD 1:1", "FADD2+2 IADD", "FADD+IADD 1:1",
                              "fft mix (2 FADD2,1 FFMA2,2 FADD,1 FMUL,1 FFMA)", "I2F", "LDS.64", "STS.64", "FADD2+LDS.64 4:1"};
static const int PER_ITER[] = {8, 8, 8, 8, 8, 16, 16, 24, 16, 56, 8, 8, 8, 40};

template <int MODE>
__global__ void k(float* out, int iters, float seed) {
    extern __shared__ float2 sm[];
    float a[8]; u64 p[8]; int n[8];
    for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x + i; p[i] = pk(seed + i, seed + threadIdx.x); n[i] = threadIdx.x + i; }
    const float c = seed * 1e-3f, m = 0.999f;
    const u64 c2 = pk(c, c), m2 = pk(m, m);
    float2* my = sm + threadIdx.x;                 // 8-byte stride: conflict free
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == FFMA) a[i] = ffma(a[i], m, c);
                if (MODE == FADD) a[i] = fadd(a[i], c);
                if (MODE == FFMA2) p[i] = ffma2(p[i], m2, c2);
                if (MODE == FADD2) p[i] = fadd2(p[i], c2);
                if (MODE == FMUL2) p[i] = fmul2(p[i], m2);
                if (MODE == MIX_FADD2_FADD) { p[i] = fadd2(p[i], c2); a[i] = fadd(a[i], c); }
                if (MODE == MIX_FADD2_IADD) { p[i] = fadd2(p[i], c2); n[i] = iadd(n[i], it); }
                if (MODE == MIX_FADD2_2IADD) { p[i] = fadd2(p[i], c2); n[i] = iadd(n[i], it); n[i] = lop(n[i], r); }
                if (MODE == MIX_FADD_IADD) { a[i] = fadd(a[i], c); n[i] = iadd(n[i], it); }
                if (MODE == MIX_FFT) {
                    p[i] = fadd2(p[i], c2); a[i] = fadd(a[i], c); p[i] = ffma2(p[i], m2, c2); a[i] = fadd(a[i], m);
                    p[i] = fadd2(p[i], m2); a[i] = a[i] * m; a[i] = ffma(a[i], m, c);
                }
                if (MODE == I2F) a[i] = i2f(__float_as_int(a[i]) & 0xff);
                if (MODE == LDS64) { float2 v = lds64(my + 256 * ((i + n[0]) & 7)); a[i] += v.x; }
                if (MODE == STS64) { sts64(my + 256 * i, a[i], c); }
                if (MODE == MIX_FADD2_LDS) {
                    p[i] = fadd2(p[i], c2); p[i] = fadd2(p[i], m2); p[i] = fadd2(p[i], c2); p[i] = fadd2(p[i], m2);
                    float2 v = lds64(my + 256 * ((i + n[0]) & 7)); n[i] ^= __float_as_int(v.x);
                }
            }
        }
    }
    float s = 0; u64 q = 0; int z = 0;
    for (int i = 0; i < 8; ++i) { s += a[i]; q ^= p[i]; z ^= n[i]; }
    if (s == 123.456f || q == 123456ull || z == 0x7fffffff) out[0] = s;
}
template <int MODE> void run(int warps_per_sm) {
    float* d; cudaMalloc(&d, 4);
    const int threads = 256, blocks = 148 * (warps_per_sm / 8), inner = 2048, iters = 5;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 256 * 8 + 64);
    // 100 KB of dynamic smem per block caps residency at the number of blocks we want per SM
    const int smem = (warps_per_sm / 8 >= 2) ? 100 * 1024 : 200 * 1024;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads, smem>>>(d, inner, 1.0f);
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) k<MODE><<<blocks, threads, smem>>>(d, inner, 1.0f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double instr = double(PER_ITER[MODE]) * 8 * inner * threads * blocks * iters;
    printf("%-52s warps/SM %2d  %8.3f ms  %5.2f warp-instr/clk/SM (at 1.965 GHz)\n", NAMES[MODE], warps_per_sm, ms,
           instr / 32 / (ms * 1e-3) / (148.0 * 1.965e9));
    cudaFree(d);
}
template <int MODE> void sweep() { run<MODE>(8); run<MODE>(16); if (MODE + 1 < NMODES) sweep<(MODE + 1 < NMODES) ? MODE + 1 : MODE>(); }
int main() { sweep<0>(); return 0; }
