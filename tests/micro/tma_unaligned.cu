// Micro-test: does a tiled TMA load of uint8 data accept an x coordinate that is not a multiple of 16?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_unaligned tma_unaligned.cu && ./tma_unaligned <x> <bx> <swizzle 0|1|2>
// Prints "OK" when the box {bx, 8} fetched at (x, 3) holds img[3 + r][x + c] (after un-swizzling).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap tm, int x, int y, int bytes, unsigned char* out) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(sm);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(d), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(b), "r"(x), "r"(y), "r"(0) : "memory");
    }
    __syncthreads();
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(b) : "memory");
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = sm[i];
}
int main(int argc, char** argv) {
    const int x = atoi(argv[1]), bx = atoi(argv[2]), swz = atoi(argv[3]);
    const int H = 64, W = 256, by = 8, y = 3;
    std::vector<unsigned char> img(H * W);
    for (int i = 0; i < H * W; ++i) img[i] = (unsigned char)((i * 2654435761u) >> 13);
    unsigned char *dimg, *dout;
    cudaMalloc(&dimg, H * W); cudaMalloc(&dout, bx * by);
    cudaMemcpy(dimg, img.data(), H * W, cudaMemcpyHostToDevice);
    void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
    CUtensorMap tm;
    cuuint64_t dims[3] = {W, H, 1}; cuuint64_t str[2] = {W, (cuuint64_t)H * W};
    cuuint32_t box[3] = {(cuuint32_t)bx, by, 1}; cuuint32_t es[3] = {1, 1, 1};
    const CUtensorMapSwizzle s = swz == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE : (swz == 1 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B);
    CUresult r = ((EncodeTiledFn)ptr)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, dimg, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, s,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("x=%d bx=%d swz=%d: encode failed %d\n", x, bx, swz, (int)r); return 0; }
    k<<<1, 128, bx * by>>>(tm, x, y, bx * by, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("x=%d bx=%d swz=%d: FAULT %s\n", x, bx, swz, cudaGetErrorString(e)); return 0; }
    std::vector<unsigned char> out(bx * by);
    cudaMemcpy(out.data(), dout, bx * by, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int rr = 0; rr < by; ++rr)
        for (int c = 0; c < bx; ++c) {
            int chunk = c >> 4;
            if (swz == 1) chunk ^= (rr >> 2) & 1;          // 32B swizzle: 256-byte pattern, bit 4 ^= bit 7
            if (swz == 2) chunk ^= (rr >> 1) & 3;          // 64B swizzle: 512-byte pattern, bits 4-5 ^= bits 7-8
            const unsigned char got = out[rr * bx + chunk * 16 + (c & 15)];
            const int gx = x + c;
            const unsigned char want = (gx < W) ? img[(y + rr) * W + gx] : 0;
            bad += (got != want);
        }
    printf("x=%d bx=%d swz=%d: %s (%d mismatches)\n", x, bx, swz, bad ? "MISMATCH" : "OK", bad);
    return 0;
}
