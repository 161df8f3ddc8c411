// General-size pass: interrogation windows of ANY even size up to 128 px (the reference accepts any
// window, e.g. 48 px or the 42 / 28 px that multipass_scale = 1.5 produces; PB:453-456, 855-857).
//
// The fused in-register FFT kernels (piv_fused.cuh) exist for 16 / 32 / 64 px only.  Everything else
// takes this path: one CTA per window reads the (shifted) window straight from the frames with the
// reference's flat-index addressing (PB:147-216), evaluates the circular cross-correlation with a
// direct DFT in shared memory (one radix-2/4 split, O(w^3 / r)) (both frames packed into one complex transform, FP32, twiddles
// from a table computed in FP64), subtracts the minimum (PB:518/724/796) and writes the fft-shifted
// map to a scratch buffer; correlation_to_displacement (corr_to_disp_kernel, PB:346-422) and the
// predictor glue (PB:728-738 / 800-810) follow as two small kernels.  Correct for every geometry,
// roughly 15-30x slower per window than the fused kernels -- a completeness path, not the headline one.
//
// Included by pivb200.cu (needs corr_to_disp_kernel and grid_for).
#pragma once

namespace pivb200 {

constexpr int kGenericMaxWindow = 128;
constexpr int kGenericThreads = 256;
constexpr int kGenericShiftClamp = 1 << 20;          // same documented clamp as the fused kernels

struct GenericParams {
    const unsigned char* fa;
    const unsigned char* fb;
    long long pair_stride;
    int H, Wf, pitch;
    int wind, n_rows, n_cols, step;
    long long first_window, n_windows;     // this launch handles windows [first, first + n)
    int mode;                              // PIVB200_MODE_*; shifts may be null (pass 1)
    const float* sxf;
    const float* syf;
    const int* sxi;
    const int* syi;
    int normalize;                         // pass 1: windows divided by their mean (PB:513-514)
    int subtract_min;
    const void* wa;                        // explicit windows instead of frames (correalte_fft API)
    const void* wb;
    int explicit_dtype;                    // 0 float32, 1 uint8
    float* corr_out;                       // [n_windows][w][w], fft-shifted
    float* win_a_out;                      // optional: the (shifted) windows themselves
    float* win_b_out;
};

// one pixel of a shifted window, exactly like bilinear_cws_kernel / shift_dws_kernel above
__device__ __forceinline__ float generic_fetch(const GenericParams& p, const unsigned char* frame, int gy, int gx,
                                               bool cws, float vxf, float vyf, int vxi, int vyi) {
    const long long last = static_cast<long long>(p.H) * p.Wf - 1;
    auto at = [&](long long q) {
        q = q < 0 ? 0 : (q > last ? last : q);
        const long long y = q / p.Wf, x = q - y * p.Wf;
        return static_cast<float>(frame[y * p.pitch + x]);
    };
    if (!cws) return at(static_cast<long long>(gy + vyi) * p.Wf + (gx + vxi));
    const float ny = __fadd_rn(static_cast<float>(gy), vyf);
    const float nx = __fadd_rn(static_cast<float>(gx), vxf);
    const float uxf = ceilf(nx), uyf = ceilf(ny), dxf = floorf(nx), dyf = floorf(ny);
    const long long ux = static_cast<long long>(uxf), uy = static_cast<long long>(uyf);
    const long long dx = static_cast<long long>(dxf), dy = static_cast<long long>(dyf);
    const float q11 = at(dy * p.Wf + dx), q12 = at(uy * p.Wf + dx), q21 = at(dy * p.Wf + ux), q22 = at(uy * p.Wf + ux);
    const float wx1 = __fsub_rn(uxf, nx), wx0 = __fsub_rn(nx, dxf);
    const float wy1 = __fsub_rn(uyf, ny), wy0 = __fsub_rn(ny, dyf);
    float acc = __fmul_rn(__fmul_rn(q11, wx1), wy1);
    acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(q21, wx0), wy1));
    acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(q12, wx1), wy0));
    acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(q22, wx0), wy0));
    return ((ux - dx) * (uy - dy) == 0) ? q11 : acc;
}

__device__ __forceinline__ float block_reduce(float v, float* red, int op) {   // op 0: sum, 1: min
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float t = __shfl_xor_sync(0xffffffffu, v, o);
        v = op ? fminf(v, t) : v + t;
    }
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = red[0];
    for (int i = 1; i < kGenericThreads / 32; ++i) r = op ? fminf(r, red[i]) : r + red[i];
    return r;
}

// In-place DFT of every line of Z (rows: elements Z[L][x]; columns: Z[x][L]); a warp owns a line.
// One decimation-in-time split by r = 4 (w >= 96, 4 | w), 2, or 1 keeps the lanes busy and divides the
// O(w^2) work per line by r:  S_q[k'] = sum_x' z[r x' + q] W_m^{x' k'}  (m = w / r, direct sums, lane l owns
// k' = l, l + 32, ...), then X[k' + m t] = sum_q W_r^{q t} (W_w^{q k'} S_q[k']).  conj = inverse transform
// (unnormalised).  Four accumulators per lane in every case: (r, outputs per lane) = (1, 4), (2, 2), (4, 1).
__device__ __forceinline__ float2 gmul(float2 a, float2 t) {
    return make_float2(fmaf(a.x, t.x, -a.y * t.y), fmaf(a.x, t.y, a.y * t.x));
}
template <int R>
__device__ __forceinline__ void generic_lines_r(float2* Z, const float2* tw, int w, int pitch, bool columns, bool conj) {
    constexpr int J = 4 / R;                     // outputs k' per lane
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int es = columns ? pitch : 1, ls = columns ? 1 : pitch;
    const int m = w / R;
    const float sgn = conj ? -1.f : 1.f;
    for (int L = warp; L < w; L += kGenericThreads / 32) {
        float2* line = Z + L * ls;
        float2 acc[R][J];
        int idx[J], stepk[J];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            idx[j] = 0;
            stepk[j] = (R * (lane + 32 * j)) % w;                          // W_m^{k'} = W_w^{R k'}
#pragma unroll
            for (int q = 0; q < R; ++q) acc[q][j] = make_float2(0.f, 0.f);
        }
        for (int x = 0; x < m; ++x) {
            float2 v[R];
#pragma unroll
            for (int q = 0; q < R; ++q) v[q] = line[(R * x + q) * es];       // broadcast reads
#pragma unroll
            for (int j = 0; j < J; ++j) {
                if (lane + 32 * j < m) {
                    float2 t = tw[idx[j]];
                    t.y *= sgn;
#pragma unroll
                    for (int q = 0; q < R; ++q) {
                        acc[q][j].x = fmaf(v[q].x, t.x, fmaf(-v[q].y, t.y, acc[q][j].x));
                        acc[q][j].y = fmaf(v[q].x, t.y, fmaf(v[q].y, t.x, acc[q][j].y));
                    }
                    idx[j] += stepk[j];
                    if (idx[j] >= w) idx[j] -= w;
                }
            }
        }
        __syncwarp();                                                        // the whole line was read
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int k = lane + 32 * j;
            if (k >= m) continue;
            float2 T[R];
            T[0] = acc[0][j];
#pragma unroll
            for (int q = 1; q < R; ++q) {
                float2 t = tw[(q * k) % w];
                t.y *= sgn;
                T[q] = gmul(acc[q][j], t);
            }
            if constexpr (R == 1) {
                line[k * es] = T[0];
            } else if constexpr (R == 2) {
                line[k * es] = make_float2(T[0].x + T[1].x, T[0].y + T[1].y);
                line[(k + m) * es] = make_float2(T[0].x - T[1].x, T[0].y - T[1].y);
            } else {
                // radix-4 butterfly with W_4 = -i (forward) or +i (inverse)
                const float2 s02 = make_float2(T[0].x + T[2].x, T[0].y + T[2].y);
                const float2 d02 = make_float2(T[0].x - T[2].x, T[0].y - T[2].y);
                const float2 s13 = make_float2(T[1].x + T[3].x, T[1].y + T[3].y);
                const float2 d13 = make_float2(T[1].x - T[3].x, T[1].y - T[3].y);
                const float2 rot = make_float2(sgn * d13.y, -sgn * d13.x);     // -i d13 (forward), +i d13 (inverse)
                line[k * es] = make_float2(s02.x + s13.x, s02.y + s13.y);
                line[(k + m) * es] = make_float2(d02.x + rot.x, d02.y + rot.y);
                line[(k + 2 * m) * es] = make_float2(s02.x - s13.x, s02.y - s13.y);
                line[(k + 3 * m) * es] = make_float2(d02.x - rot.x, d02.y - rot.y);
            }
        }
    }
    __syncthreads();
}
__device__ __forceinline__ void generic_lines(float2* Z, const float2* tw, int w, int pitch, bool columns, bool conj) {
    if (w % 4 == 0 && w >= 96) generic_lines_r<4>(Z, tw, w, pitch, columns, conj);
    else if (w >= 16) generic_lines_r<2>(Z, tw, w, pitch, columns, conj);
    else generic_lines_r<1>(Z, tw, w, pitch, columns, conj);
}

__global__ void __launch_bounds__(kGenericThreads) generic_corr_kernel(const GenericParams p) {
    extern __shared__ __align__(16) unsigned char gsm[];
    const int w = p.wind, pitch = w + 1, half = w / 2;
    float2* Z = reinterpret_cast<float2*>(gsm);
    float2* tw = Z + w * pitch;
    float* red = reinterpret_cast<float*>(tw + w);
    for (int k = threadIdx.x; k < w; k += blockDim.x) {
        double s, c;
        sincospi(2.0 * k / w, &s, &c);
        tw[k] = make_float2(static_cast<float>(c), static_cast<float>(-s));      // e^{-2 pi i k / w}
    }
    __syncthreads();
    const int per_pair = p.n_rows * p.n_cols;
    for (long long wq = blockIdx.x; wq < p.n_windows; wq += gridDim.x) {
        const long long g = p.first_window + wq;
        // ---- the two windows, packed as z = a + i b -------------------------------------------
        float sum_a = 0.f, sum_b = 0.f;
        if (p.wa != nullptr) {
            for (int e = threadIdx.x; e < w * w; e += blockDim.x) {
                const long long q = g * w * w + e;
                const float a = p.explicit_dtype ? static_cast<float>(static_cast<const unsigned char*>(p.wa)[q])
                                                 : static_cast<const float*>(p.wa)[q];
                const float b = p.explicit_dtype ? static_cast<float>(static_cast<const unsigned char*>(p.wb)[q])
                                                 : static_cast<const float*>(p.wb)[q];
                Z[(e / w) * pitch + e % w] = make_float2(a, b);
            }
        } else {
            const long long pair = g / per_pair;
            const int loc = static_cast<int>(g - pair * per_pair);
            const int r0 = (loc / p.n_cols) * p.step, c0 = (loc % p.n_cols) * p.step;
            const unsigned char* fa = p.fa + pair * p.pair_stride;
            const unsigned char* fb = p.fb + pair * p.pair_stride;
            const bool cws = (p.mode == PIVB200_MODE_CWS) && p.sxf != nullptr;
            float vxf = 0.f, vyf = 0.f;
            int vxi = 0, vyi = 0;
            if (cws) {
                const float lim = static_cast<float>(kGenericShiftClamp);
                vxf = fminf(fmaxf(p.sxf[g], -lim), lim);
                vyf = fminf(fmaxf(p.syf[g], -lim), lim);
            } else if (p.sxi != nullptr) {
                vxi = max(-kGenericShiftClamp, min(kGenericShiftClamp, p.sxi[g]));
                vyi = max(-kGenericShiftClamp, min(kGenericShiftClamp, p.syi[g]));
            }
            for (int e = threadIdx.x; e < w * w; e += blockDim.x) {
                const int i = e / w, j = e - i * w;
                // frame a is shifted by -s, frame b by +s (PB:720-723 / 792-795)
                const float a = generic_fetch(p, fa, r0 + i, c0 + j, cws, -vxf, -vyf, -vxi, -vyi);
                const float b = generic_fetch(p, fb, r0 + i, c0 + j, cws, vxf, vyf, vxi, vyi);
                Z[i * pitch + j] = make_float2(a, b);
                sum_a += a;
                sum_b += b;
                if (p.win_a_out) {
                    p.win_a_out[g * w * w + e] = a;
                    p.win_b_out[g * w * w + e] = b;
                }
            }
        }
        if (p.corr_out == nullptr) { __syncthreads(); continue; }         // windows only
        if (p.subtract_min && p.wa == nullptr) {
            // Pass mode.  The window means are removed before the transform: subtracting a mean shifts every
            // correlation value by the same constant, which `- amin` removes anyway, and the FP32 map
            // keeps ~4 more significant bits (same reason the fused kernels drop the DC bin).  Pass 1
            // additionally divides by the mean (PB:513-514).
            const float ma = block_reduce(sum_a, red, 0) / static_cast<float>(w * w);
            const float mb = block_reduce(sum_b, red, 0) / static_cast<float>(w * w);
            for (int e = threadIdx.x; e < w * w; e += blockDim.x) {
                float2& z = Z[(e / w) * pitch + e % w];
                z = make_float2(z.x - ma, z.y - mb);
                if (p.normalize) z = make_float2(z.x / ma, z.y / mb);     // 0 / 0 = NaN like the reference
            }
        }
        __syncthreads();
        // ---- Z^ = DFT2(a + i b) ----------------------------------------------------------------
        generic_lines(Z, tw, w, pitch, false, false);
        generic_lines(Z, tw, w, pitch, true, false);
        // ---- P = conj(A^) B^ with A^ = (Z^[k] + conj Z^[-k]) / 2, B^ = (Z^[k] - conj Z^[-k]) / 2i -----
        for (int e = threadIdx.x; e < w * w; e += blockDim.x) {
            const int ky = e / w, kx = e - ky * w;
            const int ny = ky ? w - ky : 0, nx = kx ? w - kx : 0;
            const int en = ny * w + nx;
            if (e > en) continue;                                          // the partner's thread does both
            const float2 zk = Z[ky * pitch + kx], zn = Z[ny * pitch + nx];
            const float2 A = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
            const float2 B = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));
            const float2 P = make_float2(fmaf(A.x, B.x, A.y * B.y), fmaf(A.x, B.y, -A.y * B.x));
            Z[ky * pitch + kx] = P;
            if (en != e) Z[ny * pitch + nx] = make_float2(P.x, -P.y);
        }
        __syncthreads();
        // ---- inverse transform (real result) ---------------------------------------------------
        generic_lines(Z, tw, w, pitch, true, true);
        generic_lines(Z, tw, w, pitch, false, true);
        const float scale = 1.0f / (static_cast<float>(w) * static_cast<float>(w));
        float mn = 0.f;
        if (p.subtract_min) {
            float m = FLT_MAX;
            for (int e = threadIdx.x; e < w * w; e += blockDim.x) m = fminf(m, Z[(e / w) * pitch + e % w].x * scale);
            mn = block_reduce(m, red, 1);
        }
        float* out = p.corr_out + wq * w * w;
        for (int e = threadIdx.x; e < w * w; e += blockDim.x) {
            const int i = e / w, j = e - i * w;                             // position in the fft-shifted map
            const int si = i >= half ? i - half : i + half, sj = j >= half ? j - half : j + half;
            float val = Z[si * pitch + sj].x * scale;
            // NaN maps (black window divided by its zero mean) stay NaN: fminf drops NaNs, the subtraction keeps them
            out[e] = p.subtract_min ? val - mn : val;
        }
        __syncthreads();
    }
}

// u = base + du, predictor replacement (PB:731-738 / 803-810)
__global__ void generic_glue_kernel(const double* __restrict__ du, const double* __restrict__ dv,
                                    const unsigned char* __restrict__ invalid, const double* __restrict__ base_u,
                                    const double* __restrict__ base_v, const double* __restrict__ pred_u,
                                    const double* __restrict__ pred_v, long long n, double* __restrict__ u,
                                    double* __restrict__ v) {
    for (long long g = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; g < n;
         g += static_cast<long long>(gridDim.x) * blockDim.x) {
        const double a = du[g], b = dv[g];
        double uo = a + (base_u ? base_u[g] : 0.0), vo = b + (base_v ? base_v[g] : 0.0);
        if (pred_u) {
            const bool bad = invalid && invalid[g];
            const double pu = pred_u[g], pv = pred_v[g];
            if ((a > pu && rint(pu) > 0.0) || bad) uo = pu;
            if ((b > pv && rint(pv) > 0.0) || bad) vo = pv;
        }
        u[g] = uo;
        v[g] = vo;
    }
}

inline bool generic_window_ok(int wind) { return wind >= 4 && wind <= kGenericMaxWindow && wind % 2 == 0; }

inline size_t generic_smem_bytes(int w) {
    return static_cast<size_t>(w) * (w + 1) * sizeof(float2) + static_cast<size_t>(w) * sizeof(float2) + 64 * sizeof(float);
}

inline int generic_launch(const GenericParams& gp, cudaStream_t s) {
    const size_t smem = generic_smem_bytes(gp.wind);
    cudaError_t err = cudaFuncSetAttribute(generic_corr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem));
    if (err != cudaSuccess) return static_cast<int>(err);
    int dev = 0, sms = 0, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, generic_corr_kernel, kGenericThreads, smem);
    long long grid = static_cast<long long>(sms) * (per_sm > 0 ? per_sm : 1);
    if (grid > gp.n_windows) grid = gp.n_windows;
    generic_corr_kernel<<<static_cast<unsigned>(grid), kGenericThreads, smem, s>>>(gp);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

// A whole pass for a general window size: correlation maps in chunks through a stream-ordered scratch
// buffer, then correlation_to_displacement and the predictor glue.
inline int run_generic_pass(const unsigned char* fa, const unsigned char* fb, int n_pairs, long long pair_stride,
                            int H, int W, int pitch, int wind, int overlap, const PassParams& p, int mode,
                            cudaStream_t s) {
    if (!generic_window_ok(wind)) return PIVB200_E_WINDOW;
    if (overlap >= wind || overlap < 0) return PIVB200_E_OVERLAP;
    if (wind > H || wind > W || pitch < W) return PIVB200_E_FRAME;
    if (!fa || !fb || n_pairs < 1) return PIVB200_E_ARG;
    const int n_rows = (H - wind) / (wind - overlap) + 1, n_cols = (W - wind) / (wind - overlap) + 1;
    const long long n_total = static_cast<long long>(n_rows) * n_cols * n_pairs;
    if (n_total >= (1ll << 30) || static_cast<long long>(H) * W >= (1ll << 31)) return PIVB200_E_SIZE;
    GenericParams gp;
    memset(&gp, 0, sizeof(gp));
    gp.fa = fa; gp.fb = fb;
    gp.pair_stride = (n_pairs > 1) ? pair_stride : static_cast<long long>(H) * pitch;
    gp.H = H; gp.Wf = W; gp.pitch = pitch;
    gp.wind = wind; gp.n_rows = n_rows; gp.n_cols = n_cols; gp.step = wind - overlap;
    gp.mode = mode;
    gp.sxf = p.sxf; gp.syf = p.syf; gp.sxi = p.sxi; gp.syi = p.syi;
    gp.normalize = p.first_pass;
    gp.subtract_min = 1;
    // scratch: maps of one chunk + du, dv of the whole pass
    const long long map_elems = static_cast<long long>(wind) * wind;
    long long chunk = (256ll << 20) / (map_elems * 4);                    // <= 256 MB of maps at a time
    if (chunk < 1) chunk = 1;
    if (chunk > n_total) chunk = n_total;
    float* maps = nullptr;
    double* dd = nullptr;
    {
        // keep the scratch memory in the device's default pool between passes (the default release
        // threshold of 0 hands it back to the driver at every synchronisation: a cudaMalloc per pass)
        static std::once_flag pool_once;
        std::call_once(pool_once, [] {
            int dev = 0;
            cudaMemPool_t pool;
            unsigned long long keep = 1ull << 30;
            if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess)
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        });
    }
    cudaError_t err = cudaMallocAsync(reinterpret_cast<void**>(&maps), chunk * map_elems * 4, s);
    if (err != cudaSuccess) return static_cast<int>(err);
    err = cudaMallocAsync(reinterpret_cast<void**>(&dd), n_total * 16, s);
    if (err != cudaSuccess) { cudaFreeAsync(maps, s); return static_cast<int>(err); }
    double* du = dd;
    double* dv = dd + n_total;
    int rc = 0;
    for (long long first = 0; first < n_total && rc == 0; first += chunk) {
        const long long n = (n_total - first < chunk) ? n_total - first : chunk;
        gp.first_window = first;
        gp.n_windows = n;
        gp.corr_out = maps;
        rc = generic_launch(gp, s);
        if (rc) break;
        corr_to_disp_kernel<float><<<grid_for(n * 32, 128), 128, 0, s>>>(
            maps, n, wind, wind, p.validate, p.val_ratio, 3, du + first, dv + first, p.mask ? p.mask + first : nullptr);
        count_launch();
        rc = static_cast<int>(cudaGetLastError());
    }
    if (rc == 0) {
        generic_glue_kernel<<<grid_for(n_total, 256), 256, 0, s>>>(du, dv, p.validate ? p.mask : nullptr, p.base_u,
                                                                   p.base_v, p.pred_u, p.pred_v, n_total, p.u, p.v);
        count_launch();
        rc = static_cast<int>(cudaGetLastError());
        if (rc == 0 && p.ratio) rc = static_cast<int>(cudaMemsetAsync(p.ratio, 0, n_total * 4, s));
    }
    cudaFreeAsync(maps, s);
    cudaFreeAsync(dd, s);
    return rc;
}

}  // namespace pivb200
