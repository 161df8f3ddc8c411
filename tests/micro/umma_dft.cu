// Micro-benchmark (not product code): the FIRST stage of the window transform -- the 64-point real DFT of
// every row of a uint8 interrogation window -- on the 5th-generation tensor cores (tcgen05.mma, kind::f16).
//
//   D[128 x 64] (FP32, TMEM)  =  A[128 x 64] (two windows' rows, uint8 -> fp16: exact)  x  T[64 x 64]
//   T = [cos(2 pi k x / 64), k = 0..32 | -sin(2 pi k x / 64), k = 1..31], split hi + lo in fp16 (two MMAs
//   accumulate into the same TMEM tile), so the result has ~FP32 accuracy.
//
// It answers the question DESIGN.md section 8 leaves for round 2: what would the tensor pipe take for the
// two row-FFT steps of a pass (plus the u8 -> float conversion), which today run on the FP32 pipe?
// Operands are staged in the canonical K-major SWIZZLE_128B shared-memory layout, one elected thread
// issues the MMAs, completion comes through tcgen05.commit on an mbarrier, the accumulator is read back
// with tcgen05.ld (lane = window row, columns = frequency bins).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_dft umma_dft.cu && ./umma_dft
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int W = 64;            // window size = K = N
constexpr int TILE_M = 128;      // two windows per MMA tile
constexpr int THREADS = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// K-major SWIZZLE_128B: rows of 128 bytes (64 fp16), atoms of 8 rows (1024 B), 16-byte chunk c of row r at chunk c ^ (r & 7)
__host__ __device__ inline int sw128_offset(int row, int chunk) { return (row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    // start address >> 4 [0,14), LBO [16,30) (unused for swizzled K-major), SBO = 1024 B >> 4 [32,46), version 1 [46,48),
    // layout SWIZZLE_128B = 2 [61,64)   (cute::UMMA::SmemDescriptor)
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr >> 4) & 0x3fff);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
// kind::f16 instruction descriptor: C = F32 (1 << 4), A = B = F16 (0), K-major both, N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t kIdesc = (1u << 4) | (static_cast<uint32_t>(W >> 3) << 17) | (static_cast<uint32_t>(TILE_M >> 4) << 24);

__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}

// grid-stride over tiles of two windows; mode 0: write the spectra (validation), 1: fold them into a checksum (timing)
__global__ void __launch_bounds__(THREADS) umma_dft_kernel(const uint8_t* __restrict__ win, int n_tiles,
                                                           const __half* __restrict__ t_hi, const __half* __restrict__ t_lo,
                                                           float* __restrict__ out, int mode) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sA = smem;                       // 16 KB
    unsigned char* sBhi = smem + 16384;             // 8 KB
    unsigned char* sBlo = smem + 16384 + 8192;      // 8 KB
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;

    // twiddle tiles: B[n][k] = T[k][n], K-major SW128 (global layout [n][k])
    for (int e = tid; e < W * 8; e += THREADS) {
        const int n = e >> 3, c = e & 7;
        *reinterpret_cast<uint4*>(sBhi + sw128_offset(n, c)) = *reinterpret_cast<const uint4*>(t_hi + n * W + c * 8);
        *reinterpret_cast<uint4*>(sBlo + sw128_offset(n, c)) = *reinterpret_cast<const uint4*>(t_lo + n * W + c * 8);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base;
    uint32_t parity = 0;
    float checksum = 0.f;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // ---- thread r converts row r of the tile (64 uint8 -> 64 fp16, exact) into the swizzled A tile ----
        const uint4* src = reinterpret_cast<const uint4*>(win + (static_cast<size_t>(tile) * TILE_M + tid) * W);
#pragma unroll
        for (int q = 0; q < 4; ++q) {               // 16 bytes -> two 16-byte chunks of fp16
            const uint4 v = __ldg(src + q);
            const uint32_t wds[4] = {v.x, v.y, v.z, v.w};
            uint32_t h[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                // bytes -> 0x64bb = 1024 + b in fp16, minus 1024
                const uint32_t lo = __byte_perm(wds[i], 0x64646464u, 0x4140), hi = __byte_perm(wds[i], 0x64646464u, 0x4342);
                const __half2 k1024 = __floats2half2_rn(1024.f, 1024.f);
                __half2 a = __hsub2(*reinterpret_cast<const __half2*>(&lo), k1024);
                __half2 b = __hsub2(*reinterpret_cast<const __half2*>(&hi), k1024);
                h[2 * i] = *reinterpret_cast<uint32_t*>(&a);
                h[2 * i + 1] = *reinterpret_cast<uint32_t*>(&b);
            }
            *reinterpret_cast<uint4*>(sA + sw128_offset(tid, 2 * q)) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(sA + sw128_offset(tid, 2 * q + 1)) = make_uint4(h[4], h[5], h[6], h[7]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t a0 = make_desc(smem_u32(sA)), bh = make_desc(smem_u32(sBhi)), bl = make_desc(smem_u32(sBlo));
#pragma unroll
            for (int k = 0; k < W / 16; ++k) {       // UMMA_K = 16 fp16 = 32 bytes = +2 in the (>> 4) start address
                mma_f16(tmem_d, a0 + 2 * k, bh + 2 * k, kIdesc, k > 0);
            }
#pragma unroll
            for (int k = 0; k < W / 16; ++k) mma_f16(tmem_d, a0 + 2 * k, bl + 2 * k, kIdesc, 1);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        mbar_wait(smem_u32(&bar), parity);
        parity ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- accumulator: lane = tile row (this thread's row), 64 columns = the row's half spectrum ----
        float d[64];
        const uint32_t taddr = tmem_d + (static_cast<uint32_t>(warp * 32) << 16);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=f"(d[32 * c + 0]), "=f"(d[32 * c + 1]), "=f"(d[32 * c + 2]), "=f"(d[32 * c + 3]), "=f"(d[32 * c + 4]),
                  "=f"(d[32 * c + 5]), "=f"(d[32 * c + 6]), "=f"(d[32 * c + 7]), "=f"(d[32 * c + 8]), "=f"(d[32 * c + 9]),
                  "=f"(d[32 * c + 10]), "=f"(d[32 * c + 11]), "=f"(d[32 * c + 12]), "=f"(d[32 * c + 13]), "=f"(d[32 * c + 14]),
                  "=f"(d[32 * c + 15]), "=f"(d[32 * c + 16]), "=f"(d[32 * c + 17]), "=f"(d[32 * c + 18]), "=f"(d[32 * c + 19]),
                  "=f"(d[32 * c + 20]), "=f"(d[32 * c + 21]), "=f"(d[32 * c + 22]), "=f"(d[32 * c + 23]), "=f"(d[32 * c + 24]),
                  "=f"(d[32 * c + 25]), "=f"(d[32 * c + 26]), "=f"(d[32 * c + 27]), "=f"(d[32 * c + 28]), "=f"(d[32 * c + 29]),
                  "=f"(d[32 * c + 30]), "=f"(d[32 * c + 31])
                : "r"(taddr + 32 * c)
                : "memory");
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (mode == 0) {
            float* o = out + (static_cast<size_t>(tile) * TILE_M + tid) * W;
#pragma unroll
            for (int c = 0; c < 64; ++c) o[c] = d[c];
        } else {
#pragma unroll
            for (int c = 0; c < 64; ++c) checksum += d[c];
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                              // TMEM tile and A tile are free again
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (mode == 1) out[blockIdx.x * THREADS + tid] = checksum;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem_base) : "memory");
}

int main(int argc, char** argv) {
    const int pairs = argc > 1 ? atoi(argv[1]) : 8;
    const int n_win = pairs * 2 * 3969;                 // both frames of `pairs` 4 MP pairs, 64 px / 50 %
    const int n_tiles = n_win / 2;
    // twiddles: column n of T: n = 0..32 -> cos(2 pi n x / 64); n = 33..63 -> -sin(2 pi (n - 32) x / 64)
    std::vector<__half> thi(W * W), tlo(W * W);
    std::vector<double> tref(W * W);
    for (int n = 0; n < W; ++n)
        for (int x = 0; x < W; ++x) {
            const int k = n <= 32 ? n : n - 32;
            const double ang = 2.0 * M_PI * ((k * x) % W) / W;
            const double v = n <= 32 ? cos(ang) : -sin(ang);
            const __half h = __float2half_rn(static_cast<float>(v));
            thi[n * W + x] = h;
            tlo[n * W + x] = __float2half_rn(static_cast<float>(v - static_cast<double>(__half2float(h))));
            tref[n * W + x] = v;
        }
    std::vector<uint8_t> hwin(static_cast<size_t>(n_win) * W * W);
    srand(1);
    for (auto& b : hwin) b = static_cast<uint8_t>(rand() & 0xff);
    uint8_t* dwin; __half *dhi, *dlo; float* dout;
    CK(cudaMalloc(&dwin, hwin.size()));
    CK(cudaMalloc(&dhi, W * W * 2)); CK(cudaMalloc(&dlo, W * W * 2));
    const int check_tiles = 64;
    CK(cudaMalloc(&dout, static_cast<size_t>(check_tiles) * TILE_M * W * 4 + (1 << 20)));
    CK(cudaMemcpy(dwin, hwin.data(), hwin.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dhi, thi.data(), W * W * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dlo, tlo.data(), W * W * 2, cudaMemcpyHostToDevice));
    const int smem = 32768 + 1024;
    CK(cudaFuncSetAttribute(umma_dft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));

    // ---- validation of the first tiles against a double-precision DFT ----
    umma_dft_kernel<<<check_tiles, THREADS, smem>>>(dwin, check_tiles, dhi, dlo, dout, 0);
    CK(cudaDeviceSynchronize());
    std::vector<float> hout(static_cast<size_t>(check_tiles) * TILE_M * W);
    CK(cudaMemcpy(hout.data(), dout, hout.size() * 4, cudaMemcpyDeviceToHost));
    double max_err = 0, max_val = 0;
    for (int r = 0; r < check_tiles * TILE_M; ++r)
        for (int n = 0; n < W; ++n) {
            double acc = 0;
            for (int x = 0; x < W; ++x) acc += hwin[static_cast<size_t>(r) * W + x] * tref[n * W + x];
            max_err = fmax(max_err, fabs(acc - hout[static_cast<size_t>(r) * W + n]));
            max_val = fmax(max_val, fabs(acc));
        }
    printf("validation: %d rows x 64 bins, max |error| %.3e (max |value| %.1f, relative %.2e)\n", check_tiles * TILE_M, max_err,
           max_val, max_err / max_val);

    // ---- timing: all tiles, several CTAs per SM so conversion, MMA and read-back of different CTAs overlap ----
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int per_sm : {1, 2, 4, 6}) {
        const int grid = sms * per_sm;
        umma_dft_kernel<<<grid, THREADS, smem>>>(dwin, n_tiles, dhi, dlo, dout, 1);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        const int reps = 5;
        for (int i = 0; i < reps; ++i) umma_dft_kernel<<<grid, THREADS, smem>>>(dwin, n_tiles, dhi, dlo, dout, 1);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= reps;
        const double flops = static_cast<double>(n_tiles) * 2.0 * TILE_M * W * W * 2;       // hi + lo
        printf("CTAs/SM %d: %.3f ms for %d pairs (both frames) = %.1f us per pair, %.1f dense TFLOP/s on the tensor pipe, %.0f GB/s of windows\n",
               per_sm, ms, pairs, ms * 1e3 / pairs, flops / ms / 1e9, hwin.size() / ms / 1e6);
    }
    printf("reference: the two row-FFT steps + u8->f32 conversion take ~1/3 of the fused FP32 pass (54 us per pair) today\n");
    return 0;
}
