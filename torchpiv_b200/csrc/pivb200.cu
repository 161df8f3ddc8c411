// C ABI of libpivb200.so (see include/pivb200.h) + the small non-fused kernels:
// predictor resampling (spline operator as two tiny FP64 matrix products), the function-level
// correlation_to_displacement / window-shift kernels and the FFMA peak micro-benchmark.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cfloat>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>

#include "../../include/pivb200.h"
#include "fused_launch.cuh"
#include "soa_launch.cuh"
#include "piv_params.h"

namespace pivb200 {

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ----------------------------------------------------------------------------------------
// TMA tensor maps (driver entry point resolved at run time: no link against libcuda)
// ----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    });
    return fn;
}

// frames [n_pairs][H][pitch] uint8 -> 3-D map {W, H, n_pairs}, box {bx, by, 1}
static int make_frame_map(CUtensorMap* map, const uint8_t* base, int n_pairs, long long pair_stride,
                          int H, int W, int pitch, int bx, int by) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return PIVB200_E_DRIVER;
    const long long ps = (n_pairs > 1) ? pair_stride : static_cast<long long>(H) * pitch;
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                          static_cast<cuuint64_t>(n_pairs)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(pitch), static_cast<cuuint64_t>(ps)};
    cuuint32_t box[3] = {static_cast<cuuint32_t>(bx), static_cast<cuuint32_t>(by), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    // = Tile<W, LOADER>::SWZ
    const CUtensorMapSwizzle swz = (bx == 32) ? CU_TENSOR_MAP_SWIZZLE_32B
                                              : ((bx == 64) ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE);
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(base), dims, strides,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : PIVB200_E_DRIVER;
}

static bool tma_ok(const uint8_t* a, const uint8_t* b, int n_pairs, long long pair_stride, int pitch) {
    // PIVB200_DISABLE_TMA=1 routes every window through the flat-index gather path (test knob)
    static const bool disabled = [] { const char* e = getenv("PIVB200_DISABLE_TMA"); return e && e[0] == '1'; }();
    if (disabled) return false;
    if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) return false;
    if (pitch % 16) return false;
    if (n_pairs > 1 && (pair_stride % 16)) return false;
    return get_encode_fn() != nullptr;
}

static int check_geometry(int H, int W, int pitch, int wind, int overlap, int n_pairs,
                          int* n_rows, int* n_cols) {
    if (wind != 16 && wind != 32 && wind != 64) return PIVB200_E_WINDOW;
    if (overlap >= wind || overlap < 0) return PIVB200_E_OVERLAP;
    if (wind > H || wind > W || pitch < W || n_pairs < 1) return PIVB200_E_FRAME;
    *n_rows = (H - wind) / (wind - overlap) + 1;
    *n_cols = (W - wind) / (wind - overlap) + 1;
    if (static_cast<long long>(*n_rows) * *n_cols * n_pairs >= (1ll << 30)) return PIVB200_E_SIZE;
    if (static_cast<long long>(H) * W >= (1ll << 31)) return PIVB200_E_SIZE;
    return 0;
}

static int launch_fused(int wind, int loader, int sink, const CUtensorMap& ta, const CUtensorMap& tb,
                        const PassParams& p, cudaStream_t s) {
    // displacement passes run the pair-packed kernels (piv_soa.cuh); PIVB200_SOA=0 (all sizes) / PIVB200_SOA64=0
    // (64 px only) keep the one-transform-per-lane kernels (piv_fused.cuh) for A/B measurements
    static const bool soa = [] { const char* e = getenv("PIVB200_SOA"); return !(e && e[0] == '0'); }();
    if (soa && sink == SK_DISP && (loader == LD_FRAME_INT || loader == LD_FRAME_ALN || loader == LD_FRAME_CWS)) {
        static const bool soa64 = [] { const char* e = getenv("PIVB200_SOA64"); return !(e && e[0] == '0'); }();
        if (wind == 64 && soa64) return launch_soa_w64(loader, ta, tb, p, s);
        if (wind == 32) return launch_soa_w32(loader, ta, tb, p, s);
        if (wind == 16) return launch_soa_w16(loader, ta, tb, p, s);
    }
    switch (wind) {
        case 64: return launch_fused_w64(loader, sink, ta, tb, p, s);
        case 32: return launch_fused_w32(loader, sink, ta, tb, p, s);
        case 16: return launch_fused_w16(loader, sink, ta, tb, p, s);
    }
    return PIVB200_E_WINDOW;
}

// general window sizes (generic_pass.cuh, included further down)
static int run_generic(const uint8_t* fa, const uint8_t* fb, int n_pairs, long long pair_stride, int H, int W,
                       int pitch, int wind, int overlap, int loader, int sink, const PassParams& p,
                       cudaStream_t stream);
static inline bool fused_window(int wind) { return wind == 16 || wind == 32 || wind == 64; }

static int run_frame_pass(const uint8_t* fa, const uint8_t* fb, int n_pairs, long long pair_stride,
                          int H, int W, int pitch, int wind, int overlap, int loader, int sink,
                          PassParams& p, cudaStream_t stream) {
    if (!fused_window(wind))
        return run_generic(fa, fb, n_pairs, pair_stride, H, W, pitch, wind, overlap, loader, sink, p, stream);
    int n_rows, n_cols;
    int rc = check_geometry(H, W, pitch, wind, overlap, n_pairs, &n_rows, &n_cols);
    if (rc) return rc;
    if (!fa || !fb) return PIVB200_E_ARG;
    p.fa = fa;
    p.fb = fb;
    p.pair_stride = (n_pairs > 1) ? pair_stride : static_cast<long long>(H) * pitch;
    p.H = H;
    p.Wf = W;
    p.pitch = pitch;
    p.n_rows = n_rows;
    p.n_cols = n_cols;
    p.step = wind - overlap;
    p.n_total = static_cast<long long>(n_rows) * n_cols * n_pairs;
    p.div_n = make_fastdiv(static_cast<uint32_t>(n_rows) * static_cast<uint32_t>(n_cols));
    p.div_c = make_fastdiv(static_cast<uint32_t>(n_cols));
    // unshifted windows on a grid whose step is a multiple of 16 px (the usual first pass) start 16-byte
    // aligned in every row: the aligned loader skips the realignment network (PIVB200_NO_ALIGNED=1: A/B knob)
    static const bool no_aligned = [] { const char* e = getenv("PIVB200_NO_ALIGNED"); return e && e[0] == '1'; }();
    if (loader == LD_FRAME_INT && p.sxi == nullptr && p.step % 16 == 0 && !no_aligned) loader = LD_FRAME_ALN;
    // PIVB200_TC=1 (with PIVB200_SOA64=0): experimental 64 px first pass whose row transform runs on the tensor cores
    static const bool use_tc = [] { const char* e = getenv("PIVB200_TC"); return e != nullptr && e[0] == '1'; }();
    if (loader == LD_FRAME_ALN && wind == 64 && sink == SK_DISP && use_tc) loader = LD_FRAME_TC;
    CUtensorMap ta, tb;
    memset(&ta, 0, sizeof(ta));
    memset(&tb, 0, sizeof(tb));
    p.use_tma = tma_ok(fa, fb, n_pairs, pair_stride, pitch) ? 1 : 0;
    if (p.use_tma) {
        const int bx = (loader == LD_FRAME_ALN || loader == LD_FRAME_TC) ? wind : wind + 16;   // = Tile<W, LOADER>::BX
        const int by = (loader == LD_FRAME_CWS) ? wind + 1 : wind;
        if (make_frame_map(&ta, fa, n_pairs, pair_stride, H, W, pitch, bx, by) ||
            make_frame_map(&tb, fb, n_pairs, pair_stride, H, W, pitch, bx, by))
            p.use_tma = 0;      // still exact: every window takes the flat-index gather path
    }
    return launch_fused(wind, loader, sink, ta, tb, p, stream);
}

// ----------------------------------------------------------------------------------------
// predictor resampling: T = U * Ax^T, S = Ay * T for the fields u, v, mask (FP64)
// ----------------------------------------------------------------------------------------
// Both kernels give a thread RB = 4 outputs that share one operand (register blocking: 4 independent
// FP64 accumulation chains per field and 5 instead of 8 loads per 4 FMAs); every output is still summed
// in ascending q, so the result does not depend on the blocking.
constexpr int kPredRB = 4;

__global__ void predictor_rows_kernel(const double* __restrict__ u, const double* __restrict__ v,
                                      const uint8_t* __restrict__ mask, int n_pairs, int n0, int m0,
                                      int m1, const double* __restrict__ Ax, double* __restrict__ tmp) {
    // tmp[f][pair][i][j] = sum_q F[pair][i][q] * Ax[j][q]; a thread owns rows i..i+3 of one (f, pair, j)
    constexpr int RB = kPredRB;
    const int ib = (n0 + RB - 1) / RB;
    const long long per_field = static_cast<long long>(n_pairs) * n0 * m1;
    const long long per_field_t = static_cast<long long>(n_pairs) * ib * m1;
    const int nf = mask ? 3 : 2;
    for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < per_field_t * nf;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int f = static_cast<int>(e / per_field_t);
        long long r = e - f * per_field_t;
        const int j = static_cast<int>(r % m1);
        r /= m1;
        const int i0 = static_cast<int>(r % ib) * RB;
        const long long pair = r / ib;
        const double* ax = Ax + static_cast<long long>(j) * m0;
        double acc[RB] = {0.0, 0.0, 0.0, 0.0};
        int row[RB];
#pragma unroll
        for (int k = 0; k < RB; ++k) row[k] = min(i0 + k, n0 - 1);       // clamped rows are computed but not stored
        if (f < 2) {
            const double* src = (f == 0 ? u : v) + pair * n0 * m0;
            for (int q = 0; q < m0; ++q) {
                const double a = ax[q];
#pragma unroll
                for (int k = 0; k < RB; ++k) acc[k] = fma(src[static_cast<long long>(row[k]) * m0 + q], a, acc[k]);
            }
        } else {
            const uint8_t* src = mask + pair * n0 * m0;
            for (int q = 0; q < m0; ++q) {
                const double a = ax[q];
#pragma unroll
                for (int k = 0; k < RB; ++k)
                    acc[k] = fma(src[static_cast<long long>(row[k]) * m0 + q] ? 1.0 : 0.0, a, acc[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < RB; ++k)
            if (i0 + k < n0) tmp[f * per_field + (pair * n0 + i0 + k) * m1 + j] = acc[k];
    }
}

template <int MODE>
__global__ void predictor_cols_kernel(const double* __restrict__ tmp, int has_mask, int n_pairs, int n0,
                                      int n1, int m1, const double* __restrict__ Ay,
                                      void* __restrict__ shift_x, void* __restrict__ shift_y,
                                      double* __restrict__ base_u, double* __restrict__ base_v,
                                      double* __restrict__ pred_u, double* __restrict__ pred_v) {
    // a thread owns output rows i..i+3 of one (pair, j): every tmp element it loads feeds 4 FMAs per field
    constexpr int RB = kPredRB;
    const int ib = (n1 + RB - 1) / RB;
    const long long per_field = static_cast<long long>(n_pairs) * n0 * m1;
    const long long n_thr = static_cast<long long>(n_pairs) * ib * m1;
    for (long long t_id = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t_id < n_thr;
         t_id += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int j = static_cast<int>(t_id % m1);
        const long long r = t_id / m1;
        const int i0 = static_cast<int>(r % ib) * RB;
        const long long pair = r / ib;
        const double* t = tmp + (pair * n0) * m1 + j;
        const double* ay[RB];
#pragma unroll
        for (int k = 0; k < RB; ++k) ay[k] = Ay + static_cast<long long>(min(i0 + k, n1 - 1)) * n0;
        double su[RB] = {0.0, 0.0, 0.0, 0.0}, sv[RB] = {0.0, 0.0, 0.0, 0.0}, sm[RB] = {0.0, 0.0, 0.0, 0.0};
        for (int q = 0; q < n0; ++q) {
            const double tu = t[static_cast<long long>(q) * m1];
            const double tv = t[per_field + static_cast<long long>(q) * m1];
            const double tm = has_mask ? t[2 * per_field + static_cast<long long>(q) * m1] : 0.0;
#pragma unroll
            for (int k = 0; k < RB; ++k) {
                const double a = ay[k][q];
                su[k] = fma(a, tu, su[k]);
                sv[k] = fma(a, tv, sv[k]);
                if (has_mask) sm[k] = fma(a, tm, sm[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < RB; ++k) {
            if (i0 + k >= n1) break;
            const long long e = (pair * n1 + i0 + k) * m1 + j;
            const bool inval = has_mask && (sm[k] >= 0.5);           // PB:711 / 778
            const double pu = inval ? 0.0 : su[k], pv = inval ? 0.0 : sv[k];
            pred_u[e] = pu;
            pred_v[e] = pv;
            if (MODE == PIVB200_MODE_CWS) {
                // PB:705-706: halves taken BEFORE the invalid zeroing; shift = float32(u0 / 2)
                const double hu = su[k] / 2, hv = sv[k] / 2;
                static_cast<float*>(shift_x)[e] = static_cast<float>(hu);
                static_cast<float*>(shift_y)[e] = static_cast<float>(hv);
                base_u[e] = 2 * hu;
                base_v[e] = 2 * hv;
            } else {
                // PB:782-790: zeroing first, round-half-even
                const double ru = rint(pu / 2), rv = rint(pv / 2);
                const double lim = 1048576.0;
                static_cast<int*>(shift_x)[e] = static_cast<int>(fmin(fmax(ru, -lim), lim));
                static_cast<int*>(shift_y)[e] = static_cast<int>(fmin(fmax(rv, -lim), lim));
                base_u[e] = 2 * ru;
                base_v[e] = 2 * rv;
            }
        }
    }
}

// ----------------------------------------------------------------------------------------
// correlation_to_displacement on arbitrary maps: one warp per map (PB:360-422)
// ----------------------------------------------------------------------------------------
template <typename T>
__global__ void corr_to_disp_kernel(T* __restrict__ corr, long long n, int d, int k, int validate,
                                    double val_ratio, int vw, double* __restrict__ u,
                                    double* __restrict__ v, uint8_t* __restrict__ mask) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
    const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const int n2 = d * k;
    for (long long c = warp; c < n; c += nwarps) {
        T* map = corr + c * n2;
        // corr += eps (in corr's dtype), argmax = first maximum; NaN counts as maximal (torch)
        T best = 0;
        int bi = n2;
        bool bnan = false;
        for (int e = lane; e < n2; e += 32) {
            const T val = map[e] + static_cast<T>(1e-7);
            map[e] = val;
            const bool isn = (val != val);
            if (bi == n2 || (isn && !bnan) || (!bnan && !isn && val > best)) {
                if (!(bnan && !isn)) { best = val; bi = e; bnan = isn; }
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const T ob = __shfl_xor_sync(FULL, best, o);
            const int oi = __shfl_xor_sync(FULL, bi, o);
            const bool on = __shfl_xor_sync(FULL, bnan ? 1 : 0, o) != 0;
            bool take;
            if (oi == n2) take = false;
            else if (bi == n2) take = true;
            else if (on != bnan) take = on;
            else if (on) take = oi < bi;
            else take = (ob > best) || (ob == best && oi < bi);
            if (take) { best = ob; bi = oi; bnan = on; }
        }
        __syncwarp();
        const int m = bi;
        int left = m + 1, right = m - 1, top = m + k, bot = m - k;
        if (left >= n2 - 1) left = m;
        if (right <= 0) right = m;
        if (top >= n2 - 1) top = m;
        if (bot <= 0) bot = m;
        const double cm = static_cast<double>(map[m]), cl = static_cast<double>(map[left]),
                     cr = static_cast<double>(map[right]), ct = static_cast<double>(map[top]),
                     cb = static_cast<double>(map[bot]);
        const double lm = log(cm), ll = log(cl), lr = log(cr), lt = log(ct), lb = log(cb);
        double uu = static_cast<double>(m % k) + (lr - ll) / (2.0 * (ll + lr) - 4.0 * lm);
        double vv = static_cast<double>(m / d) + (lb - lt) / (2.0 * (lb + lt) - 4.0 * lm);
        bool invalid = false;
        if (validate) {
            // zero the clamped flat-index patch in place, remembering what was there
            const int side = 2 * vw + 1, np = side * side;
            T saved[4];
            int sid[4];
#pragma unroll
            for (int h = 0; h < 4; ++h) { saved[h] = 0; sid[h] = -1; }
            __syncwarp();
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const int q = lane + 32 * h;
                if (q < np && q < 128) {
                    const int i = q / side - vw, j = q % side - vw;      // reference loop order: i outer
                    int id = m + i + k * j;
                    id = id < 0 ? 0 : (id > n2 - 1 ? n2 - 1 : id);
                    sid[h] = id;
                    saved[h] = map[id];
                }
            }
            __syncwarp();
#pragma unroll
            for (int h = 0; h < 4; ++h)
                if (sid[h] >= 0) map[sid[h]] = 0;
            __syncwarp();
            T b2 = 0;
            int i2 = n2;
            bool n2nan = false;
            for (int e = lane; e < n2; e += 32) {
                const T val = map[e];
                const bool isn = (val != val);
                if (i2 == n2 || (isn && !n2nan) || (!n2nan && !isn && val > b2)) {
                    if (!(n2nan && !isn)) { b2 = val; i2 = e; n2nan = isn; }
                }
            }
            for (int o = 16; o > 0; o >>= 1) {
                const T ob = __shfl_xor_sync(FULL, b2, o);
                const int oi = __shfl_xor_sync(FULL, i2, o);
                const bool on = __shfl_xor_sync(FULL, n2nan ? 1 : 0, o) != 0;
                bool take;
                if (oi == n2) take = false;
                else if (i2 == n2) take = true;
                else if (on != n2nan) take = on;
                else if (on) take = oi < i2;
                else take = (ob > b2) || (ob == b2 && oi < i2);
                if (take) { b2 = ob; i2 = oi; n2nan = on; }
            }
            double c2 = static_cast<double>(b2);
            if (sizeof(T) == 4) {
                // float32: the reference's float64 copy was taken BEFORE the zeroing (PB:382), so a
                // second "peak" inside the patch reads its original value
                for (int h = 0; h < 4; ++h) {
                    const unsigned hit = __ballot_sync(FULL, sid[h] == i2);
                    if (hit) c2 = static_cast<double>(__shfl_sync(FULL, saved[h], __ffs(hit) - 1));
                }
            }
            invalid = (cm / c2) < val_ratio;
        }
        vv -= static_cast<double>(d / 2);
        uu -= static_cast<double>(k / 2);
        uu = isnan(uu) ? 0.0 : (isinf(uu) ? copysign(DBL_MAX, uu) : uu);
        vv = isnan(vv) ? 0.0 : (isinf(vv) ? copysign(DBL_MAX, vv) : vv);
        if (lane == 0) {
            u[c] = uu;
            v[c] = vv;
            if (mask) mask[c] = invalid ? 1 : 0;
        }
        __syncwarp();
    }
}

// ----------------------------------------------------------------------------------------
// reference-layout window shifts (PB:147-216), one thread per element, bit-exact arithmetic
// ----------------------------------------------------------------------------------------
__global__ void bilinear_cws_kernel(const uint8_t* __restrict__ frame, int H, int W,
                                    const int64_t* __restrict__ grid, long long n_elem, int epw,
                                    const float* __restrict__ vel_x, const float* __restrict__ vel_y,
                                    float* __restrict__ out) {
    const long long last = static_cast<long long>(H) * W - 1;
    for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < n_elem;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long idx = grid[e];
        const long long win = e / epw;
        const long long gy = idx / W, gx = idx - gy * W;     // PB:163 (floor division; idx >= 0)
        const float ny = __fadd_rn(static_cast<float>(gy), vel_y[win]);
        const float nx = __fadd_rn(static_cast<float>(gx), vel_x[win]);
        const float uxf = ceilf(nx), uyf = ceilf(ny), dxf = floorf(nx), dyf = floorf(ny);
        const long long ux = static_cast<long long>(uxf), uy = static_cast<long long>(uyf);
        const long long dx = static_cast<long long>(dxf), dy = static_cast<long long>(dyf);
        auto tap = [&](long long yy, long long xx) {
            long long q = yy * W + xx;
            q = q < 0 ? 0 : (q > last ? last : q);
            return static_cast<float>(frame[q]);
        };
        const float q11 = tap(dy, dx), q12 = tap(uy, dx), q21 = tap(dy, ux), q22 = tap(uy, ux);
        const float wx1 = __fsub_rn(static_cast<float>(ux), nx), wx0 = __fsub_rn(nx, static_cast<float>(dx));
        const float wy1 = __fsub_rn(static_cast<float>(uy), ny), wy0 = __fsub_rn(ny, static_cast<float>(dy));
        float acc = __fmul_rn(__fmul_rn(q11, wx1), wy1);
        acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(q21, wx0), wy1));
        acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(q12, wx1), wy0));
        acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(q22, wx0), wy0));
        out[e] = ((ux - dx) * (uy - dy) == 0) ? q11 : acc;
    }
}

__global__ void shift_dws_kernel(const uint8_t* __restrict__ frame, int H, int W,
                                 const int64_t* __restrict__ grid, long long n_elem, int epw,
                                 const int64_t* __restrict__ vel_x, const int64_t* __restrict__ vel_y,
                                 uint8_t* __restrict__ out) {
    const long long last = static_cast<long long>(H) * W - 1;
    for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < n_elem;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long win = e / epw;
        long long q = grid[e] + vel_y[win] * W + vel_x[win];
        q = q < 0 ? 0 : (q > last ? last : q);
        out[e] = frame[q];
    }
}

// ----------------------------------------------------------------------------------------
// FFMA peak micro-benchmark (roofline denominator for the FP32-bound fused kernel)
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, int iters, float seed) {
    float a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
          a6 = a0 + 6, a7 = a0 + 7;
    const float m = 0.999f, c = 0.001f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
        }
    }
    const float s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456f) out[0] = s;      // never true; keeps the chain alive
}

static int grid_for(long long n, int block) {
    long long g = (n + block - 1) / block;
    return static_cast<int>(g > 148 * 32 ? 148 * 32 : (g < 1 ? 1 : g));
}

}  // namespace pivb200
#include "generic_pass.cuh"
namespace pivb200 {

static int run_generic(const uint8_t* fa, const uint8_t* fb, int n_pairs, long long pair_stride, int H, int W,
                       int pitch, int wind, int overlap, int loader, int sink, const PassParams& p,
                       cudaStream_t stream) {
    const int mode = (loader == LD_FRAME_CWS) ? PIVB200_MODE_CWS : PIVB200_MODE_DWS;
    if (sink == SK_DISP)
        return run_generic_pass(fa, fb, n_pairs, pair_stride, H, W, pitch, wind, overlap, p, mode, stream);
    if (sink != SK_WIN) return PIVB200_E_ARG;
    // the (shifted) windows only
    if (!generic_window_ok(wind)) return PIVB200_E_WINDOW;
    if (overlap >= wind || overlap < 0) return PIVB200_E_OVERLAP;
    if (wind > H || wind > W || pitch < W) return PIVB200_E_FRAME;
    if (!fa || !fb || n_pairs < 1) return PIVB200_E_ARG;
    GenericParams gp;
    memset(&gp, 0, sizeof(gp));
    gp.fa = fa; gp.fb = fb;
    gp.pair_stride = (n_pairs > 1) ? pair_stride : static_cast<long long>(H) * pitch;
    gp.H = H; gp.Wf = W; gp.pitch = pitch;
    gp.wind = wind;
    gp.n_rows = (H - wind) / (wind - overlap) + 1;
    gp.n_cols = (W - wind) / (wind - overlap) + 1;
    gp.step = wind - overlap;
    gp.mode = mode;
    gp.sxf = p.sxf; gp.syf = p.syf; gp.sxi = p.sxi; gp.syi = p.syi;
    gp.n_windows = static_cast<long long>(gp.n_rows) * gp.n_cols * n_pairs;
    if (gp.n_windows >= (1ll << 30) || static_cast<long long>(H) * W >= (1ll << 31)) return PIVB200_E_SIZE;
    gp.win_a_out = p.win_a_out;
    gp.win_b_out = p.win_b_out;
    return generic_launch(gp, stream);
}

}  // namespace pivb200

using namespace pivb200;

// ========================================================================================
// C ABI
// ========================================================================================
extern "C" {

int pivb200_version(void) { return 100; }

const char* pivb200_error_string(int code) {
    switch (code) {
        case PIVB200_OK: return "ok";
        case PIVB200_E_WINDOW: return "interrogation window must be an even size of 4..128 px (16/32/64 px: fused kernels)";
        case PIVB200_E_OVERLAP: return "Overlap has to be smaller than the window_size";
        case PIVB200_E_FRAME: return "window size cannot be larger than the image";
        case PIVB200_E_ARG: return "invalid argument (null or misaligned pointer, bad count)";
        case PIVB200_E_DRIVER: return "cuTensorMapEncodeTiled unavailable";
        case PIVB200_E_SIZE: return "too many windows for one call";
    }
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "unknown error";
}

int pivb200_field_shape(int H, int W, int wind, int overlap, int* n_rows, int* n_cols) {
    if (overlap >= wind) return PIVB200_E_OVERLAP;
    if (wind > H || wind > W) return PIVB200_E_FRAME;
    if (!n_rows || !n_cols) return PIVB200_E_ARG;
    *n_rows = (H - wind) / (wind - overlap) + 1;
    *n_cols = (W - wind) / (wind - overlap) + 1;
    return 0;
}

int pivb200_pass_first(const uint8_t* frames_a, const uint8_t* frames_b, int n_pairs,
                       long long pair_stride, int H, int W, int pitch, int wind, int overlap,
                       int validate, double val_ratio, double* u, double* v, uint8_t* mask,
                       float* ratio, void* stream) {
    if (!u || !v || (validate && !mask)) return PIVB200_E_ARG;
    PassParams p;
    memset(&p, 0, sizeof(p));
    p.first_pass = 1;
    p.validate = validate;
    p.val_ratio = val_ratio;
    p.u = u;
    p.v = v;
    p.mask = mask;
    p.ratio = ratio;
    return run_frame_pass(frames_a, frames_b, n_pairs, pair_stride, H, W, pitch, wind, overlap,
                          LD_FRAME_INT, SK_DISP, p, static_cast<cudaStream_t>(stream));
}

int pivb200_pass_next(const uint8_t* frames_a, const uint8_t* frames_b, int n_pairs,
                      long long pair_stride, int H, int W, int pitch, int wind, int overlap,
                      int mode, const void* shift_x, const void* shift_y, const double* base_u,
                      const double* base_v, const double* pred_u, const double* pred_v,
                      int validate, double val_ratio, double* u, double* v, uint8_t* mask,
                      float* ratio, void* stream) {
    if (!u || !v || (validate && !mask) || !shift_x || !shift_y) return PIVB200_E_ARG;
    if ((pred_u == nullptr) != (pred_v == nullptr) || (base_u == nullptr) != (base_v == nullptr))
        return PIVB200_E_ARG;
    PassParams p;
    memset(&p, 0, sizeof(p));
    p.validate = validate;
    p.val_ratio = val_ratio;
    p.u = u;
    p.v = v;
    p.mask = mask;
    p.ratio = ratio;
    p.base_u = base_u;
    p.base_v = base_v;
    p.pred_u = pred_u;
    p.pred_v = pred_v;
    int loader;
    if (mode == PIVB200_MODE_CWS) {
        p.sxf = static_cast<const float*>(shift_x);
        p.syf = static_cast<const float*>(shift_y);
        loader = LD_FRAME_CWS;
    } else if (mode == PIVB200_MODE_DWS) {
        p.sxi = static_cast<const int*>(shift_x);
        p.syi = static_cast<const int*>(shift_y);
        loader = LD_FRAME_INT;
    } else {
        return PIVB200_E_ARG;
    }
    return run_frame_pass(frames_a, frames_b, n_pairs, pair_stride, H, W, pitch, wind, overlap, loader,
                          SK_DISP, p, static_cast<cudaStream_t>(stream));
}

int pivb200_windows(const uint8_t* frames_a, const uint8_t* frames_b, int n_pairs,
                    long long pair_stride, int H, int W, int pitch, int wind, int overlap,
                    int mode, const void* shift_x, const void* shift_y, float* win_a,
                    float* win_b, void* stream) {
    if (!win_a || !win_b) return PIVB200_E_ARG;
    if ((reinterpret_cast<uintptr_t>(win_a) | reinterpret_cast<uintptr_t>(win_b)) & 15) return PIVB200_E_ARG;
    PassParams p;
    memset(&p, 0, sizeof(p));
    p.win_a_out = win_a;
    p.win_b_out = win_b;
    int loader;
    if (mode == PIVB200_MODE_CWS) {
        if (!shift_x || !shift_y) return PIVB200_E_ARG;
        p.sxf = static_cast<const float*>(shift_x);
        p.syf = static_cast<const float*>(shift_y);
        loader = LD_FRAME_CWS;
    } else if (mode == PIVB200_MODE_DWS) {
        if ((shift_x == nullptr) != (shift_y == nullptr)) return PIVB200_E_ARG;
        p.sxi = static_cast<const int*>(shift_x);
        p.syi = static_cast<const int*>(shift_y);
        loader = LD_FRAME_INT;
    } else {
        return PIVB200_E_ARG;
    }
    return run_frame_pass(frames_a, frames_b, n_pairs, pair_stride, H, W, pitch, wind, overlap, loader,
                          SK_WIN, p, static_cast<cudaStream_t>(stream));
}

int pivb200_correlate(const void* windows_a, const void* windows_b, int dtype, long long n,
                      int wind, float* corr, void* stream) {
    if (!fused_window(wind) && !generic_window_ok(wind)) return PIVB200_E_WINDOW;
    if (!windows_a || !windows_b || !corr || n < 1 || (dtype != 0 && dtype != 1)) return PIVB200_E_ARG;
    if ((reinterpret_cast<uintptr_t>(windows_a) | reinterpret_cast<uintptr_t>(windows_b)) & 15)
        return PIVB200_E_ARG;
    if (n >= (1ll << 30)) return PIVB200_E_SIZE;
    if (!fused_window(wind)) {
        GenericParams gp;
        memset(&gp, 0, sizeof(gp));
        gp.wind = wind;
        gp.n_windows = n;
        gp.wa = windows_a;
        gp.wb = windows_b;
        gp.explicit_dtype = dtype;
        gp.corr_out = corr;
        return generic_launch(gp, static_cast<cudaStream_t>(stream));
    }
    PassParams p;
    memset(&p, 0, sizeof(p));
    p.wa = windows_a;
    p.wb = windows_b;
    p.corr_out = corr;
    p.n_total = n;
    p.n_rows = 1;
    p.n_cols = 1;
    p.step = wind;
    CUtensorMap ta, tb;
    memset(&ta, 0, sizeof(ta));
    memset(&tb, 0, sizeof(tb));
    return launch_fused(wind, dtype == 0 ? LD_EXPL_F32 : LD_EXPL_U8, SK_CORR, ta, tb, p,
                        static_cast<cudaStream_t>(stream));
}

int pivb200_predictor(const double* u_prev, const double* v_prev, const uint8_t* mask_prev,
                      int n_pairs, int n0, int m0, int n1, int m1, const double* Ay,
                      const double* Ax, int mode, double* tmp, void* shift_x, void* shift_y,
                      double* base_u, double* base_v, double* pred_u, double* pred_v,
                      void* stream) {
    if (!u_prev || !v_prev || !Ay || !Ax || !tmp || !shift_x || !shift_y || !base_u || !base_v ||
        !pred_u || !pred_v)
        return PIVB200_E_ARG;
    if (n_pairs < 1 || n0 < 1 || m0 < 1 || n1 < 1 || m1 < 1) return PIVB200_E_ARG;
    if (mode != PIVB200_MODE_CWS && mode != PIVB200_MODE_DWS) return PIVB200_E_ARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int nf = mask_prev ? 3 : 2;
    const long long n_tmp = static_cast<long long>(n_pairs) * ((n0 + kPredRB - 1) / kPredRB) * m1 * nf;
    predictor_rows_kernel<<<grid_for(n_tmp, 128), 128, 0, s>>>(u_prev, v_prev, mask_prev, n_pairs, n0, m0,
                                                               m1, Ax, tmp);
    count_launch();
    const long long n_out = static_cast<long long>(n_pairs) * ((n1 + kPredRB - 1) / kPredRB) * m1;
    if (mode == PIVB200_MODE_CWS)
        predictor_cols_kernel<PIVB200_MODE_CWS><<<grid_for(n_out, 128), 128, 0, s>>>(
            tmp, mask_prev != nullptr, n_pairs, n0, n1, m1, Ay, shift_x, shift_y, base_u, base_v, pred_u, pred_v);
    else
        predictor_cols_kernel<PIVB200_MODE_DWS><<<grid_for(n_out, 128), 128, 0, s>>>(
            tmp, mask_prev != nullptr, n_pairs, n0, n1, m1, Ay, shift_x, shift_y, base_u, base_v, pred_u, pred_v);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

int pivb200_corr_to_disp(void* corr, int dtype, long long n, int d, int k, int validate,
                         double val_ratio, int val_window, double* u, double* v, uint8_t* mask,
                         void* stream) {
    if (!corr || !u || !v || (validate && !mask) || n < 1 || d < 2 || k < 2) return PIVB200_E_ARG;
    if (val_window < 0 || (2 * val_window + 1) * (2 * val_window + 1) > 128) return PIVB200_E_ARG;
    if (static_cast<long long>(d) * k >= (1ll << 30)) return PIVB200_E_SIZE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int block = 128;
    const int grid = grid_for(n * 32, block);
    if (dtype == 0)
        corr_to_disp_kernel<float><<<grid, block, 0, s>>>(static_cast<float*>(corr), n, d, k, validate,
                                                          val_ratio, val_window, u, v, mask);
    else if (dtype == 1)
        corr_to_disp_kernel<double><<<grid, block, 0, s>>>(static_cast<double*>(corr), n, d, k, validate,
                                                           val_ratio, val_window, u, v, mask);
    else
        return PIVB200_E_ARG;
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

int pivb200_bilinear_cws(const uint8_t* frame, int H, int W, const int64_t* grid, long long n_elem,
                         int elem_per_window, const float* vel_x, const float* vel_y, float* out,
                         void* stream) {
    if (!frame || !grid || !vel_x || !vel_y || !out || n_elem < 1 || elem_per_window < 1) return PIVB200_E_ARG;
    bilinear_cws_kernel<<<grid_for(n_elem, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        frame, H, W, grid, n_elem, elem_per_window, vel_x, vel_y, out);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

int pivb200_shift_dws(const uint8_t* frame, int H, int W, const int64_t* grid, long long n_elem,
                      int elem_per_window, const int64_t* vel_x, const int64_t* vel_y,
                      uint8_t* out, void* stream) {
    if (!frame || !grid || !vel_x || !vel_y || !out || n_elem < 1 || elem_per_window < 1) return PIVB200_E_ARG;
    shift_dws_kernel<<<grid_for(n_elem, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        frame, H, W, grid, n_elem, elem_per_window, vel_x, vel_y, out);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

int pivb200_measure_fp32_peak(int iters, double* tflops_host, void* stream) {
    if (!tflops_host || iters < 1) return PIVB200_E_ARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    float* d_out = nullptr;
    cudaError_t err = cudaMalloc(&d_out, sizeof(float));
    if (err != cudaSuccess) return static_cast<int>(err);
    const int blocks = sms * 8, threads = 256, inner = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    ffma_peak_kernel<<<blocks, threads, 0, s>>>(d_out, inner, 1.0f);   // warm-up
    cudaEventRecord(e0, s);
    for (int i = 0; i < iters; ++i) {
        ffma_peak_kernel<<<blocks, threads, 0, s>>>(d_out, inner, 1.0f);
        count_launch();
    }
    cudaEventRecord(e1, s);
    err = cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    if (err != cudaSuccess) return static_cast<int>(err);
    const double flops = 2.0 * 8 * 16 * static_cast<double>(inner) * threads * blocks * iters;
    *tflops_host = flops / (ms * 1e-3) / 1e12;
    return 0;
}

long long pivb200_launch_count(void) { return g_launches.load(); }

}  // extern "C"
