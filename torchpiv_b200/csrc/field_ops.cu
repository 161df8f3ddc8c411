// Vector-field kernels that follow the fused correlation passes on the device:
//   * normalised median test (outlier detection on the 3x3 neighbourhood)
//   * 3x3 stencil replacement of invalid vectors (Jacobi sweeps, holes fill from their rim inwards)
//   * streaming statistics of a sequence of fields (running mean / Reynolds-stress moments)
//
// The reference has NO on-device counterpart: it validates by the peak ratio only and fills holes on
// the host with SciPy's Delaunay interpolation (PB:266-344, 884-892); the running statistics live in
// its Qt worker (workers.py:79-119).  These kernels are the additive "stencil" post-processing named by
// BASELINE.json's north_star; OfflinePIV's default stays the reference-exact host path
// (torchpiv_b200/postprocess.py).  One thread per vector, HBM-bound, a few hundred KB per pair.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>

#include "../../include/pivb200.h"

namespace pivb200 {
void count_launch();

namespace {

constexpr int kBlock = 256;
inline int grid_of(long long n) {
    long long g = (n + kBlock - 1) / kBlock;
    return static_cast<int>(g < 1 ? 1 : (g > 148 * 32 ? 148 * 32 : g));
}

// sorted insert-free median of up to 8 values (selection by insertion sort; n is tiny)
__device__ __forceinline__ double median_of(double* a, int n) {
    for (int i = 1; i < n; ++i) {
        const double key = a[i];
        int j = i - 1;
        while (j >= 0 && a[j] > key) { a[j + 1] = a[j]; --j; }
        a[j + 1] = key;
    }
    return (n & 1) ? a[n >> 1] : 0.5 * (a[(n >> 1) - 1] + a[n >> 1]);      // numpy.median convention
}

// neighbours of (r, c) inside the field that are usable (not flagged), excluding the centre
__device__ __forceinline__ int gather_ring(const double* __restrict__ f, const uint8_t* __restrict__ bad,
                                           int n_rows, int n_cols, int r, int c, double* out) {
    int n = 0;
#pragma unroll
    for (int dr = -1; dr <= 1; ++dr)
#pragma unroll
        for (int dc = -1; dc <= 1; ++dc) {
            if (dr == 0 && dc == 0) continue;
            const int rr = r + dr, cc = c + dc;
            if (rr < 0 || rr >= n_rows || cc < 0 || cc >= n_cols) continue;
            const int e = rr * n_cols + cc;
            if (bad != nullptr && bad[e]) continue;
            out[n++] = f[e];
        }
    return n;
}

__global__ void nmt_kernel(const double* __restrict__ u, const double* __restrict__ v,
                           const uint8_t* __restrict__ mask, long long n_total, int n_rows, int n_cols,
                           double threshold, double eps, uint8_t* __restrict__ outlier) {
    const int per = n_rows * n_cols;
    for (long long g = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; g < n_total;
         g += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long pair = g / per;
        const int e = static_cast<int>(g - pair * per);
        const int r = e / n_cols, c = e - r * n_cols;
        const double* fu = u + pair * per;
        const double* fv = v + pair * per;
        const uint8_t* bad = mask ? mask + pair * per : nullptr;
        uint8_t flag = bad ? (bad[e] ? 1 : 0) : 0;
        if (!flag) {
            double a[8], b[8];
            const int n = gather_ring(fu, bad, n_rows, n_cols, r, c, a);
            gather_ring(fv, bad, n_rows, n_cols, r, c, b);
            if (n >= 2) {          // a median test needs at least two neighbours
                const double mu = median_of(a, n), mv = median_of(b, n);
                for (int i = 0; i < n; ++i) { a[i] = fabs(a[i] - mu); b[i] = fabs(b[i] - mv); }
                const double ru = median_of(a, n), rv = median_of(b, n);
                const double tu = fabs(fu[e] - mu) / (ru + eps), tv = fabs(fv[e] - mv) / (rv + eps);
                flag = (tu > threshold || tv > threshold) ? 1 : 0;
            }
        }
        outlier[g] = flag;
    }
}

// One Jacobi sweep: every flagged vector with at least one usable neighbour becomes the median of its
// usable neighbours (values of the PREVIOUS sweep) and is unflagged in the output mask.
__global__ void replace_sweep_kernel(const double* __restrict__ u_in, const double* __restrict__ v_in,
                                     const uint8_t* __restrict__ bad_in, double* __restrict__ u_out,
                                     double* __restrict__ v_out, uint8_t* __restrict__ bad_out,
                                     long long n_total, int n_rows, int n_cols) {
    const int per = n_rows * n_cols;
    for (long long g = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; g < n_total;
         g += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long pair = g / per;
        const int e = static_cast<int>(g - pair * per);
        const int r = e / n_cols, c = e - r * n_cols;
        double nu = u_in[g], nv = v_in[g];
        uint8_t flag = bad_in[g];
        if (flag) {
            double a[8], b[8];
            const int n = gather_ring(u_in + pair * per, bad_in + pair * per, n_rows, n_cols, r, c, a);
            gather_ring(v_in + pair * per, bad_in + pair * per, n_rows, n_cols, r, c, b);
            if (n > 0) {
                nu = median_of(a, n);
                nv = median_of(b, n);
                flag = 0;
            }
        }
        u_out[g] = nu;
        v_out[g] = nv;
        bad_out[g] = flag;
    }
}

__global__ void zero_flagged_kernel(double* __restrict__ u, double* __restrict__ v,
                                    const uint8_t* __restrict__ bad, long long n_total) {
    for (long long g = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; g < n_total;
         g += static_cast<long long>(gridDim.x) * blockDim.x)
        if (bad[g]) { u[g] = 0.0; v[g] = 0.0; }
}

// Running moments of a sequence of fields, merged batch by batch (Chan et al.'s pairwise update: no
// cancellation, unlike raw power sums).  mom[0..1] = mean u, v; mom[2..4] = sums of (u-mean_u)^2,
// (v-mean_v)^2, (u-mean_u)(v-mean_v) over the `n_before` fields seen so far.
__global__ void stats_accumulate_kernel(const double* __restrict__ u, const double* __restrict__ v,
                                        int n_pairs, int per, double n_before, double* __restrict__ mom) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < per; e += gridDim.x * blockDim.x) {
        double mu = 0, mv = 0;
        for (int p = 0; p < n_pairs; ++p) {
            mu += u[static_cast<long long>(p) * per + e];
            mv += v[static_cast<long long>(p) * per + e];
        }
        const double nb = static_cast<double>(n_pairs);
        mu /= nb;
        mv /= nb;
        double suu = 0, svv = 0, suv = 0;
        for (int p = 0; p < n_pairs; ++p) {        // second read of the batch: L2 hits
            const double a = u[static_cast<long long>(p) * per + e] - mu, b = v[static_cast<long long>(p) * per + e] - mv;
            suu = fma(a, a, suu); svv = fma(b, b, svv); suv = fma(a, b, suv);
        }
        const double n = n_before + nb, w = n_before * nb / n;
        const double du = mu - mom[e], dv = mv - mom[per + e];
        mom[e] += du * nb / n;
        mom[per + e] += dv * nb / n;
        mom[2 * per + e] += suu + du * du * w;
        mom[3 * per + e] += svv + dv * dv * w;
        mom[4 * per + e] += suv + du * dv * w;
    }
}

}  // namespace
}  // namespace pivb200

using namespace pivb200;

extern "C" {

int pivb200_nmt(const double* u, const double* v, const uint8_t* mask, int n_pairs, int n_rows, int n_cols,
                double threshold, double eps, uint8_t* outlier, void* stream) {
    if (!u || !v || !outlier || n_pairs < 1 || n_rows < 1 || n_cols < 1) return PIVB200_E_ARG;
    if (!(threshold > 0.0) || !(eps >= 0.0)) return PIVB200_E_ARG;
    if (outlier == mask) return PIVB200_E_ARG;          // neighbours read the input mask
    const long long n = static_cast<long long>(n_pairs) * n_rows * n_cols;
    if (static_cast<long long>(n_rows) * n_cols >= (1ll << 31)) return PIVB200_E_SIZE;
    nmt_kernel<<<grid_of(n), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(u, v, mask, n, n_rows, n_cols,
                                                                            threshold, eps, outlier);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

long long pivb200_replace_workspace_bytes(int n_pairs, int n_rows, int n_cols) {
    if (n_pairs < 1 || n_rows < 1 || n_cols < 1) return 0;
    const long long n = static_cast<long long>(n_pairs) * n_rows * n_cols;
    return 2 * ((n * 8 + 255) / 256 * 256) + (n + 255) / 256 * 256;
}

int pivb200_replace(double* u, double* v, uint8_t* invalid, int n_pairs, int n_rows, int n_cols,
                    int max_sweeps, void* workspace, void* stream) {
    if (!u || !v || !invalid || !workspace || n_pairs < 1 || n_rows < 1 || n_cols < 1 || max_sweeps < 1)
        return PIVB200_E_ARG;
    if (reinterpret_cast<uintptr_t>(workspace) & 7) return PIVB200_E_ARG;
    const long long n = static_cast<long long>(n_pairs) * n_rows * n_cols;
    if (static_cast<long long>(n_rows) * n_cols >= (1ll << 31)) return PIVB200_E_SIZE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const long long fbytes = (n * 8 + 255) / 256 * 256;
    double* u2 = static_cast<double*>(workspace);
    double* v2 = reinterpret_cast<double*>(static_cast<char*>(workspace) + fbytes);
    uint8_t* b2 = reinterpret_cast<uint8_t*>(static_cast<char*>(workspace) + 2 * fbytes);
    // an even number of sweeps leaves the result in the caller's buffers
    const int sweeps = max_sweeps + (max_sweeps & 1);
    const int grid = grid_of(n);
    for (int it = 0; it < sweeps; ++it) {
        if ((it & 1) == 0) replace_sweep_kernel<<<grid, kBlock, 0, s>>>(u, v, invalid, u2, v2, b2, n, n_rows, n_cols);
        else replace_sweep_kernel<<<grid, kBlock, 0, s>>>(u2, v2, b2, u, v, invalid, n, n_rows, n_cols);
        count_launch();
    }
    // vectors no sweep could reach (whole field invalid, or max_sweeps too small) become 0 and stay flagged
    zero_flagged_kernel<<<grid, kBlock, 0, s>>>(u, v, invalid, n);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

int pivb200_stats_accumulate(const double* u, const double* v, int n_pairs, int n_rows, int n_cols,
                             long long n_before, double* moments, void* stream) {
    if (!u || !v || !moments || n_pairs < 1 || n_rows < 1 || n_cols < 1 || n_before < 0) return PIVB200_E_ARG;
    const long long per = static_cast<long long>(n_rows) * n_cols;
    if (per >= (1ll << 31)) return PIVB200_E_SIZE;
    stats_accumulate_kernel<<<grid_of(per), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
        u, v, n_pairs, static_cast<int>(per), static_cast<double>(n_before), moments);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
