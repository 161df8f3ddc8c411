// General-size pass: interrogation windows of ANY even size up to 256 px (the reference accepts any
// window, e.g. 48 px or the 42 / 28 px that multipass_scale = 1.5 produces; PB:453-456, 855-857).
//
// The fused in-register FFT kernels (piv_fused.cuh) exist for 16 / 32 / 64 px only.  Everything else
// takes this path: one CTA per window reads the (shifted) window straight from the frames with the
// reference's flat-index addressing (PB:147-216), evaluates the circular cross-correlation in shared
// memory (both frames packed into ONE complex transform z = a + i b, FP32, twiddles from a table computed
// in FP64), subtracts the minimum (PB:518/724/796) and writes the fft-shifted map to a scratch buffer;
// correlation_to_displacement (corr_to_disp_kernel, PB:346-422) and the predictor glue (PB:728-738 /
// 800-810) follow as two small kernels.
//
// The transform is a mixed-radix FFT (radices 16, 8, 4, 2, 3, 5, 7, 11, 13) done IN PLACE without any reordering
// pass: decimation in frequency forward (natural order in, digit-reversed order out), the spectrum product
// P = conj(A^) B^ in the digit-reversed index space (partner bin -k through a small position table), and a
// decimation-in-time transform back (digit-reversed in, natural out).  A warp takes one butterfly index, its
// lanes take the lines: every shared-memory access of a pass is then either contiguous or strided by the odd
// row pitch, i.e. bank-conflict free for any window size.  Sizes with a prime factor above 13 (22 = 2 x 11 is
// fine, 34 = 2 x 17 is not) fall back to direct sums (O(w^3 / r)).  Windows of 162-256 px keep their complex
// array in a global scratch slab (one per CTA, L2-resident) instead of shared memory; same code otherwise.
// Three instantiations: <false> (radices up to 8, 64 registers, up to 256 threads: small windows are bound by
// the latency of their ~25 block-wide phases and want many resident blocks), <true> (all radices, 512 threads),
// <true, true> (global slab).
//
// Included by pivb200.cu (needs corr_to_disp_kernel and grid_for).
#pragma once
#include "fft_regs.cuh"
#include "fft_soa.cuh"

namespace pivb200 {

// cos / sin of 2 pi x for any x, at compile time (ct_cos2pi of fft_regs.cuh is for power-of-two denominators)
__host__ __device__ constexpr double generic_ct_cos(double x) {
    x = x - static_cast<double>(static_cast<long long>(x));
    if (x < 0.0) x += 1.0;
    if (x > 0.5) x = 1.0 - x;
    bool neg = false;
    if (x > 0.25) { x = 0.5 - x; neg = true; }
    const double r = (x <= 0.125) ? ct_cos_small(2.0 * kPi * x) : ct_sin_small(2.0 * kPi * (0.25 - x));
    return neg ? -r : r;
}
__host__ __device__ constexpr double generic_ct_sin(double x) { return generic_ct_cos(x - 0.25); }

constexpr int kGenericMaxWindow = 256;           // the reference's GUI limit (ControlsWidgets.py:91)
constexpr int kGenericMaxSharedWindow = 160;     // up to here the w (w + 1) complex values of a window fit shared memory;
                                                 // larger windows keep them in a global scratch slab (L2-resident)
constexpr int kGenericThreads = 512;             // upper bound; generic_threads(w) picks the block size
constexpr int kGenericMaxStages = 8;
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
    float2* zglobal;                       // windows above kGenericMaxSharedWindow: one w (w + 1) slab per CTA, else null
    int n_stages;                          // mixed-radix plan of the window size (0 = direct sums); generic_set_plan
    int radix[kGenericMaxStages];
    int blk[kGenericMaxStages];            // block size M of stage s (decimation in frequency order): w, w / R0, ...
    FastDiv div_m[kGenericMaxStages];      // division by M / R
};

// One pixel of a shifted window, exactly like bilinear_cws_kernel / shift_dws_kernel above.  The per-axis part
// (PB:163-170: new = coord + v rounded in float32, floor / ceil taps, weights) is computed once per window column
// and once per pixel row by the caller.
struct GenericAxis {
    int up, dn;         // ceil / floor tap (absolute pixel coordinate; |shift| is clamped to 2^20)
    float w1, w0;       // (up - new), (new - down)
};
__device__ __forceinline__ GenericAxis generic_axis(int coord, float v) {
    const float nw = __fadd_rn(static_cast<float>(coord), v);
    const float upf = ceilf(nw), dnf = floorf(nw);
    GenericAxis t;
    t.up = static_cast<int>(upf);
    t.dn = static_cast<int>(dnf);
    t.w1 = __fsub_rn(upf, nw);
    t.w0 = __fsub_rn(nw, dnf);
    return t;
}
// flat index y * Wf + x clamped to the array ends (columns outside the frame wrap into the neighbouring rows,
// PB:172-177 / 207-211); pixels inside the frame skip the 64-bit division
__device__ __forceinline__ float generic_at(const GenericParams& p, const unsigned char* frame, int yy, int xx) {
    if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.Wf) return static_cast<float>(frame[yy * p.pitch + xx]);
    const long long last = static_cast<long long>(p.H) * p.Wf - 1;
    long long q = static_cast<long long>(yy) * p.Wf + xx;
    q = q < 0 ? 0 : (q > last ? last : q);
    const long long y = q / p.Wf, x = q - y * p.Wf;
    return static_cast<float>(frame[y * p.pitch + x]);
}
__device__ __forceinline__ float generic_bilinear(const GenericParams& p, const unsigned char* frame, const GenericAxis& ay,
                                                  const GenericAxis& ax) {
    const float q11 = generic_at(p, frame, ay.dn, ax.dn), q12 = generic_at(p, frame, ay.up, ax.dn);
    const float q21 = generic_at(p, frame, ay.dn, ax.up), q22 = generic_at(p, frame, ay.up, ax.up);
    float acc = __fmul_rn(__fmul_rn(q11, ax.w1), ay.w1);
    acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(q21, ax.w0), ay.w1));
    acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(q12, ax.w1), ay.w0));
    acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(q22, ax.w0), ay.w0));
    // an exact-integer coordinate on either axis: all weights vanish, the value is the (floor y, floor x) tap (PB:170, 193)
    return (ax.up == ax.dn || ay.up == ay.dn) ? q11 : acc;
}

__device__ __forceinline__ int generic_div(const FastDiv& d, int n) {
    const uint32_t un = static_cast<uint32_t>(n), t = __umulhi(d.M, un);
    return static_cast<int>((t + ((un - t) >> d.s1)) >> d.s2);
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
    for (int i = 1; i < static_cast<int>(blockDim.x >> 5); ++i) r = op ? fminf(r, red[i]) : r + red[i];
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
template <int R, int J = 4 / R>                  // J = outputs k' per lane
__device__ __forceinline__ void generic_lines_r(float2* Z, const float2* tw, int w, int pitch, bool columns, bool conj) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int es = columns ? pitch : 1, ls = columns ? 1 : pitch;
    const int m = w / R;
    const float sgn = conj ? -1.f : 1.f;
    for (int L = warp; L < w; L += static_cast<int>(blockDim.x >> 5)) {
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
// ----------------------------------------------------------------------------------------
// Mixed-radix in-place FFT of every line of Z
// ----------------------------------------------------------------------------------------
// R-point DFT, forward sign (e^{-2 pi i t q / R}), natural order in and out
template <int R>
__device__ __forceinline__ void generic_small_dft(float2 (&v)[R]) {
    // complex additions are packed FADD2 (one instruction for re and im)
    auto add = [](float2 a, float2 b) { return padd(a, b); };
    auto sub = [](float2 a, float2 b) { return psub(a, b); };
    auto mul_mi = [](float2 a) { return make_float2(a.y, -a.x); };                  // * -i
    if constexpr (R == 2) {
        const float2 t = sub(v[0], v[1]);
        v[0] = add(v[0], v[1]);
        v[1] = t;
    } else if constexpr (R == 4) {
        const float2 t0 = add(v[0], v[2]), t1 = sub(v[0], v[2]), t2 = add(v[1], v[3]), t3 = mul_mi(sub(v[1], v[3]));
        v[0] = add(t0, t2); v[2] = sub(t0, t2);
        v[1] = add(t1, t3); v[3] = sub(t1, t3);
    } else if constexpr (R == 8) {
        float2 e[4] = {v[0], v[2], v[4], v[6]}, o[4] = {v[1], v[3], v[5], v[7]};
        generic_small_dft<4>(e);
        generic_small_dft<4>(o);
        constexpr float h = 0.70710678118654752440f;
        o[1] = make_float2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));           // * (1 - i) / sqrt 2
        o[2] = mul_mi(o[2]);
        o[3] = make_float2(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y));          // * (-1 - i) / sqrt 2
#pragma unroll
        for (int q = 0; q < 4; ++q) { v[q] = add(e[q], o[q]); v[q + 4] = sub(e[q], o[q]); }
    } else if constexpr (R == 16) {
        float2 e[8], o[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) { e[q] = v[2 * q]; o[q] = v[2 * q + 1]; }
        generic_small_dft<8>(e);
        generic_small_dft<8>(o);
        static_for<1, 8>([&](auto qc) {
            constexpr int q = decltype(qc)::value;
            constexpr float c = float(generic_ct_cos(q / 16.0)), sn = float(generic_ct_sin(q / 16.0));
            o[q] = make_float2(fmaf(o[q].x, c, o[q].y * sn), fmaf(o[q].y, c, -o[q].x * sn));      // * e^{-2 pi i q / 16}
        });
#pragma unroll
        for (int q = 0; q < 8; ++q) { v[q] = add(e[q], o[q]); v[q + 8] = sub(e[q], o[q]); }
    } else {
        // odd prime: out[q] = v0 + sum_t (c a_t -+ i s b_t) with a_t = v[t] + v[R-t], b_t = v[t] - v[R-t]
        constexpr int Hh = (R - 1) / 2;
        float2 a[Hh], b[Hh];
#pragma unroll
        for (int t = 1; t <= Hh; ++t) { a[t - 1] = add(v[t], v[R - t]); b[t - 1] = sub(v[t], v[R - t]); }
        const float2 v0 = v[0];
        float2 s0 = v0;
#pragma unroll
        for (int t = 0; t < Hh; ++t) s0 = add(s0, a[t]);
        v[0] = s0;
        static_for<1, Hh + 1>([&](auto qc) {
            constexpr int q = decltype(qc)::value;
            float re = v0.x, im = v0.y, xr = 0.f, xi = 0.f;
            static_for<1, Hh + 1>([&](auto tc) {
                constexpr int t = decltype(tc)::value;
                constexpr float c = float(generic_ct_cos(double((t * q) % R) / R)), sn = float(generic_ct_sin(double((t * q) % R) / R));
                re = fmaf(c, a[t - 1].x, re);
                im = fmaf(c, a[t - 1].y, im);
                xr = fmaf(sn, b[t - 1].y, xr);          // -i s b = (s b.y, -s b.x)
                xi = fmaf(-sn, b[t - 1].x, xi);
            });
            v[q] = make_float2(re + xr, im + xi);
            v[R - q] = make_float2(re - xr, im - xi);
        });
    }
}

// One in-place radix-R pass over all w lines (rows: elements Z[L][x]; columns: Z[x][L]).  Blocks of M elements,
// butterflies over the R elements k + t M/R of a block; twiddles W_M^{t k} AFTER the butterfly (decimation in
// frequency) or BEFORE it (decimation in time).  A warp takes butterfly j, its lanes take the lines L, L + 32, ...
// `sub` is subtracted from every element as it is read (window means, first pass only).
template <int R, bool DIT>
__device__ __forceinline__ void generic_fft_pass(float2* Z, const float2* tw, int w, int pitch, bool columns, int M,
                                                 const FastDiv& div_m, float2 sub) {
    const int m = M / R, nb = w / R, tstep = (M == w) ? 1 : w / M;           // R is a compile-time constant: cheap divisions
    const int es = columns ? pitch : 1, ls = columns ? 1 : pitch;
    const int lane = threadIdx.x & 31, nwy = blockDim.x >> 5;
    const int stride = m * es;
    for (int j = threadIdx.x >> 5; j < nb; j += nwy) {
        const int b = generic_div(div_m, j), k = j - b * m;
        const int e0 = (b * M + k) * es;
        for (int L = lane; L < w; L += 32) {
            float2* base = Z + L * ls + e0;
            float2 v[R];
#pragma unroll
            for (int q = 0; q < R; ++q) {
                v[q] = base[q * stride];
                v[q].x -= sub.x;
                v[q].y -= sub.y;
            }
            if constexpr (DIT) {
                if (k != 0) {
#pragma unroll
                    for (int q = 1; q < R; ++q) v[q] = gmul(v[q], tw[q * k * tstep]);
                }
            }
            generic_small_dft<R>(v);
            if constexpr (!DIT) {
                if (k != 0) {
#pragma unroll
                    for (int q = 1; q < R; ++q) v[q] = gmul(v[q], tw[q * k * tstep]);
                }
            }
#pragma unroll
            for (int q = 0; q < R; ++q) base[q * stride] = v[q];
        }
    }
    __syncthreads();
}
// BIG = false leaves out the radices 11, 13 and 16: the kernel for small windows then fits 64 registers
template <bool DIT, bool BIG>
__device__ __forceinline__ void generic_fft_pass_r(int R, float2* Z, const float2* tw, int w, int pitch, bool columns, int M,
                                                   const FastDiv& div_m, float2 sub) {
    if constexpr (BIG) {
        if (R == 11) { generic_fft_pass<11, DIT>(Z, tw, w, pitch, columns, M, div_m, sub); return; }
        if (R == 13) { generic_fft_pass<13, DIT>(Z, tw, w, pitch, columns, M, div_m, sub); return; }
        if (R == 16) { generic_fft_pass<16, DIT>(Z, tw, w, pitch, columns, M, div_m, sub); return; }
    }
    switch (R) {
        case 2: generic_fft_pass<2, DIT>(Z, tw, w, pitch, columns, M, div_m, sub); break;
        case 3: generic_fft_pass<3, DIT>(Z, tw, w, pitch, columns, M, div_m, sub); break;
        case 4: generic_fft_pass<4, DIT>(Z, tw, w, pitch, columns, M, div_m, sub); break;
        case 5: generic_fft_pass<5, DIT>(Z, tw, w, pitch, columns, M, div_m, sub); break;
        case 7: generic_fft_pass<7, DIT>(Z, tw, w, pitch, columns, M, div_m, sub); break;
        default: generic_fft_pass<8, DIT>(Z, tw, w, pitch, columns, M, div_m, sub); break;
    }
}
// forward transform of every line: natural order in, digit-reversed order out (bin k at generic_pos(k))
template <bool BIG>
__device__ __forceinline__ void generic_fft_dif(const GenericParams& p, float2* Z, const float2* tw, int w, int pitch, bool columns,
                                                float2 sub) {
    for (int s = 0; s < p.n_stages; ++s) {
        generic_fft_pass_r<false, BIG>(p.radix[s], Z, tw, w, pitch, columns, p.blk[s], p.div_m[s], sub);
        sub = make_float2(0.f, 0.f);
    }
}
// forward transform of every line: digit-reversed order in, natural order out
template <bool BIG>
__device__ __forceinline__ void generic_fft_dit(const GenericParams& p, float2* Z, const float2* tw, int w, int pitch, bool columns) {
    for (int s = p.n_stages - 1; s >= 0; --s)
        generic_fft_pass_r<true, BIG>(p.radix[s], Z, tw, w, pitch, columns, p.blk[s], p.div_m[s], make_float2(0.f, 0.f));
}
// where generic_fft_dif leaves bin k
__device__ __forceinline__ int generic_pos(const GenericParams& p, int k, int w) {
    int pos = 0, rem = k, M = w;
    for (int s = 0; s < p.n_stages; ++s) {
        const int R = p.radix[s];
        const int nxt = rem / R;
        const int q = rem - nxt * R;
        rem = nxt;
        M /= R;
        pos += q * M;
    }
    return pos;
}

__device__ __forceinline__ void generic_lines(float2* Z, const float2* tw, int w, int pitch, bool columns, bool conj) {
    if (w % 4 == 0 && w >= 96) generic_lines_r<4>(Z, tw, w, pitch, columns, conj);
    else if (w % 2 == 0 && w >= 16) generic_lines_r<2>(Z, tw, w, pitch, columns, conj);
    else if (w <= 128) generic_lines_r<1>(Z, tw, w, pitch, columns, conj);
    else generic_lines_r<1, 8>(Z, tw, w, pitch, columns, conj);          // odd sizes above 128 px
}

// GZ: the window's complex array lives in global memory (p.zglobal) instead of shared memory
template <bool BIG, bool GZ = false>
__global__ void __launch_bounds__(BIG ? kGenericThreads : 256, BIG ? 1 : 4) generic_corr_kernel(const GenericParams p) {
    extern __shared__ __align__(16) unsigned char gsm[];
    // odd windows: the reference's irfft2 returns a [w, w - 1] map (PB:255; torch.fft.irfft2 without `s`): the inverse
    // transform along x is a length-(w - 1) c2r over the bins 0 .. (w - 1) / 2 of the length-w spectrum
    const int w = p.wind, pitch = w | 1, half = w / 2;
    const bool odd = (w & 1) != 0;
    const int nx = odd ? w - 1 : w;                                       // columns of the map
    float2* Z = GZ ? p.zglobal + static_cast<size_t>(blockIdx.x) * w * pitch : reinterpret_cast<float2*>(gsm);
    float2* tw = GZ ? reinterpret_cast<float2*>(gsm) : Z + w * pitch;
    float2* twn = tw + w;                                                 // e^{+2 pi i j / nx}, odd windows only
    float* red = reinterpret_cast<float*>(twn + (odd ? w : 0));
    unsigned short* pos_of = reinterpret_cast<unsigned short*>(red + 64);     // bin k -> position (digit reversal)
    unsigned short* bin_at = pos_of + w;                                      // position -> bin
    const bool fft = p.n_stages > 0;
    for (int k = threadIdx.x; k < w; k += blockDim.x) {
        double s, c;
        sincospi(2.0 * k / w, &s, &c);
        tw[k] = make_float2(static_cast<float>(c), static_cast<float>(-s));      // e^{-2 pi i k / w}
        const int q = fft ? generic_pos(p, k, w) : k;
        pos_of[k] = static_cast<unsigned short>(q);
        bin_at[q] = static_cast<unsigned short>(k);
        if (odd && k < nx) {
            sincospi(2.0 * k / nx, &s, &c);
            twn[k] = make_float2(static_cast<float>(c), static_cast<float>(s));
        }
    }
    __syncthreads();
    const int per_pair = p.n_rows * p.n_cols;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, nwy = blockDim.x >> 5;
    for (long long wq = blockIdx.x; wq < p.n_windows; wq += gridDim.x) {
        const long long g = p.first_window + wq;
        // ---- the two windows, packed as z = a + i b (a warp takes rows, its lanes take columns) ----
        float sum_a = 0.f, sum_b = 0.f;
        if (p.wa != nullptr) {
            for (int i = ty; i < w; i += nwy)
                for (int j = tx; j < w; j += 32) {
                    const long long q = g * w * w + i * w + j;
                    const float a = p.explicit_dtype ? static_cast<float>(static_cast<const unsigned char*>(p.wa)[q])
                                                     : static_cast<const float*>(p.wa)[q];
                    const float b = p.explicit_dtype ? static_cast<float>(static_cast<const unsigned char*>(p.wb)[q])
                                                     : static_cast<const float*>(p.wb)[q];
                    Z[i * pitch + j] = make_float2(a, b);
                }
        } else {
            const long long pair = g / per_pair;
            const int loc = static_cast<int>(g - pair * per_pair);
            const int wr = loc / p.n_cols;
            const int r0 = wr * p.step, c0 = (loc - wr * p.n_cols) * p.step;
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
            // a lane keeps its column(s): the x-axis taps are computed once per column, the y-axis taps once per pixel row;
            // frame a is shifted by -s, frame b by +s (PB:720-723 / 792-795)
            auto put = [&](int i, int j, float a, float b) {
                Z[i * pitch + j] = make_float2(a, b);
                sum_a += a;
                sum_b += b;
                if (p.win_a_out) {
                    p.win_a_out[g * w * w + i * w + j] = a;
                    p.win_b_out[g * w * w + i * w + j] = b;
                }
            };
            // unshifted / integer-shifted windows that lie inside the frame (the common case): plain byte loads, four rows
            // in flight per lane
            const bool inside = !cws && r0 - abs(vyi) >= 0 && r0 + w + abs(vyi) <= p.H && c0 - abs(vxi) >= 0 &&
                                c0 + w + abs(vxi) <= p.Wf;
            if (inside) {
                const unsigned char* pa = fa + static_cast<long long>(r0 - vyi) * p.pitch + (c0 - vxi);
                const unsigned char* pb = fb + static_cast<long long>(r0 + vyi) * p.pitch + (c0 + vxi);
                for (int j = tx; j < w; j += 32) {
                    int i = ty;
                    for (; i + 3 * nwy < w; i += 4 * nwy) {
                        unsigned char va[4], vb[4];
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            va[t] = pa[static_cast<long long>(i + t * nwy) * p.pitch + j];
                            vb[t] = pb[static_cast<long long>(i + t * nwy) * p.pitch + j];
                        }
#pragma unroll
                        for (int t = 0; t < 4; ++t) put(i + t * nwy, j, static_cast<float>(va[t]), static_cast<float>(vb[t]));
                    }
                    for (; i < w; i += nwy)
                        put(i, j, static_cast<float>(pa[static_cast<long long>(i) * p.pitch + j]),
                            static_cast<float>(pb[static_cast<long long>(i) * p.pitch + j]));
                }
            } else {
                for (int j = tx; j < w; j += 32) {
                    GenericAxis xa, xb;
                    if (cws) { xa = generic_axis(c0 + j, -vxf); xb = generic_axis(c0 + j, vxf); }
                    for (int i = ty; i < w; i += nwy) {
                        float a, b;
                        if (cws) {
                            a = generic_bilinear(p, fa, generic_axis(r0 + i, -vyf), xa);
                            b = generic_bilinear(p, fb, generic_axis(r0 + i, vyf), xb);
                        } else {
                            a = generic_at(p, fa, r0 + i - vyi, c0 + j - vxi);
                            b = generic_at(p, fb, r0 + i + vyi, c0 + j + vxi);
                        }
                        put(i, j, a, b);
                    }
                }
            }
        }
        if (p.corr_out == nullptr) { __syncthreads(); continue; }         // windows only
        // Pass mode.  The window means are removed before the transform: subtracting a mean shifts every
        // correlation value by the same constant, which `- amin` removes anyway, and the FP32 map keeps ~4 more
        // significant bits (same reason the fused kernels drop the DC bin).  Pass 1 additionally divides by the
        // means (PB:513-514): a factor 1 / (mean a * mean b) on the whole map, applied when the map is written
        // (0 / 0 = NaN for a black window, like the reference).
        float2 means = make_float2(0.f, 0.f);
        float norm = 1.0f;
        if (p.subtract_min && p.wa == nullptr) {
            means.x = block_reduce(sum_a, red, 0) / static_cast<float>(w * w);
            means.y = block_reduce(sum_b, red, 0) / static_cast<float>(w * w);
            if (p.normalize) norm = (1.0f / means.x) * (1.0f / means.y);
            if (!fft) {
                for (int i = ty; i < w; i += nwy)
                    for (int j = tx; j < w; j += 32) {
                        float2& z = Z[i * pitch + j];
                        z = make_float2(z.x - means.x, z.y - means.y);
                    }
            }
        }
        __syncthreads();
        // ---- Z^ = DFT2(a + i b) ----------------------------------------------------------------
        if (fft) {
            generic_fft_dif<BIG>(p, Z, tw, w, pitch, false, means);      // the means are subtracted as the first pass reads
            generic_fft_dif<BIG>(p, Z, tw, w, pitch, true, make_float2(0.f, 0.f));
        } else {
            generic_lines(Z, tw, w, pitch, false, false);
            generic_lines(Z, tw, w, pitch, true, false);
        }
        // ---- P = conj(A^) B^ with A^ = (Z^[k] + conj Z^[-k]) / 2, B^ = (Z^[k] - conj Z^[-k]) / 2i -----
        // (py, px) = position of bin (ky, kx), (qy, qx) = position of its partner (-ky, -kx).  The FFT path
        // stores conj(P): the map is real, so the inverse transform is the FORWARD transform of conj(P).
        for (int py = ty; py < w; py += nwy) {
            const int ky = bin_at[py];
            const int qy = pos_of[ky ? w - ky : 0];
            for (int px = tx; px < w; px += 32) {
                const int kx = bin_at[px];
                const int qx = pos_of[kx ? w - kx : 0];
                const int e = py * w + px, en = qy * w + qx;
                if (e > en) continue;                                      // the partner's thread does both
                const float2 zk = Z[py * pitch + px], zn = Z[qy * pitch + qx];
                const float2 A = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
                const float2 B = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));
                const float2 P = make_float2(fmaf(A.x, B.x, A.y * B.y), fmaf(A.x, B.y, -A.y * B.x));
                Z[py * pitch + px] = fft ? make_float2(P.x, -P.y) : P;
                if (en != e) Z[qy * pitch + qx] = fft ? P : make_float2(P.x, -P.y);
            }
        }
        __syncthreads();
        // ---- inverse transform (real result) ---------------------------------------------------
        if (fft) generic_fft_dit<BIG>(p, Z, tw, w, pitch, true);
        else generic_lines(Z, tw, w, pitch, true, true);
        if (!odd) {
            if (fft) generic_fft_dit<BIG>(p, Z, tw, w, pitch, false);
            else generic_lines(Z, tw, w, pitch, false, true);
        } else {
            // Column y now holds w Q[y][kx] (direct sums) or its conjugate (FFT path), kx at position pos_of[kx].  Row y of
            // the map: out[s] = Re Q0 + (-1)^s Re Q_{n/2} + 2 sum_{0 < k < n/2} Re(Q_k e^{2 pi i k s / n}), n = w - 1 (what a
            // c2r transform computes: the imaginary parts of bins 0 and n/2 do not enter).  A warp takes a row, its lanes
            // the samples s, s + 32, ...; the results replace the row.
            const float sgn = fft ? -1.f : 1.f;
            const int hn = nx / 2;
            for (int y = ty; y < w; y += nwy) {
                float2* row = Z + y * pitch;
                float acc[8];
                int idx[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int sidx = tx + 32 * t;
                    acc[t] = row[pos_of[0]].x + ((sidx & 1) ? -1.f : 1.f) * row[pos_of[hn]].x;
                    idx[t] = sidx % nx;                       // (k s) mod n for k = 1
                }
                for (int k = 1; k < hn; ++k) {
                    const float2 q = row[pos_of[k]];
                    const float qr = 2.f * q.x, qi = 2.f * sgn * q.y;
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        const int sidx = tx + 32 * t;
                        if (sidx < nx) {
                            const float2 e = twn[idx[t]];
                            acc[t] = fmaf(qr, e.x, fmaf(-qi, e.y, acc[t]));
                            idx[t] += sidx;
                            if (idx[t] >= nx) idx[t] -= nx;
                        }
                    }
                }
                __syncwarp();
#pragma unroll
                for (int t = 0; t < 8; ++t)
                    if (tx + 32 * t < nx) row[tx + 32 * t] = make_float2(acc[t], 0.f);
            }
            __syncthreads();
        }
        const float scale = norm / (static_cast<float>(w) * static_cast<float>(nx));
        float mn = 0.f;
        if (p.subtract_min) {
            float m = FLT_MAX;
            for (int i = ty; i < w; i += nwy)
                for (int j = tx; j < nx; j += 32) m = fminf(m, Z[i * pitch + j].x * scale);
            mn = block_reduce(m, red, 1);
        }
        // fft-shifted map [w][nx] (torch.fft.fftshift: element i comes from (i - N / 2) mod N on either axis)
        float* out = p.corr_out + wq * w * nx;
        const int hx = nx / 2, hy = w - half;
        for (int i = ty; i < w; i += nwy) {
            const int si = i + hy >= w ? i + hy - w : i + hy;
            for (int j = tx; j < nx; j += 32) {
                const int sj = j >= hx ? j - hx : j + hx;
                const float val = Z[si * pitch + sj].x * scale;
                // NaN maps (black window divided by its zero mean) stay NaN: fminf drops NaNs, the subtraction keeps them
                out[i * nx + j] = p.subtract_min ? val - mn : val;
            }
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

// any size 4..256: odd windows give the reference's [w, w - 1] maps
inline bool generic_window_ok(int wind) { return wind >= 4 && wind <= kGenericMaxWindow; }
inline int generic_map_cols(int wind) { return (wind & 1) ? wind - 1 : wind; }

inline size_t generic_smem_bytes(int w) {
    const size_t z = (w <= kGenericMaxSharedWindow) ? static_cast<size_t>(w) * (w | 1) * sizeof(float2) : 0;
    return z + static_cast<size_t>(w) * sizeof(float2) * ((w & 1) ? 2 : 1) + 64 * sizeof(float) +
           2 * static_cast<size_t>(w) * sizeof(unsigned short);
}

// threads per window: small windows are latency-bound by their ~25 block-wide phases, so they get small blocks
// and many resident blocks per SM; a 128 px window has 16 K pixels and one block per SM (shared memory)
inline int generic_threads(int w) {
    if (const char* e = getenv("PIVB200_GENERIC_THREADS")) { const int t = atoi(e); if (t >= 32 && t <= kGenericThreads && t % 32 == 0) return t; }
    const int px = w * w;
    return px <= 1024 ? 64 : (px <= 4096 ? 128 : (px <= 9216 ? 256 : kGenericThreads));
}

// radices of the window size (8, 4, 2 first, then the odd primes up to 13); n_stages = 0 -> direct sums
inline void generic_set_plan(GenericParams& gp) {
    static const int kRadices[] = {16, 8, 4, 2, 3, 5, 7, 11, 13};
    int n = gp.wind, ns = 0;
    const bool small = generic_threads(gp.wind) <= 256;       // small windows: no radix 16 (64-register kernel)
    for (int r : kRadices)
        while (n % r == 0 && ns < kGenericMaxStages && !(small && r == 16)) {
            gp.radix[ns] = r;
            gp.blk[ns] = n;
            gp.div_m[ns] = make_fastdiv(static_cast<uint32_t>(n / r));
            ++ns;
            n /= r;
        }
    gp.n_stages = (n == 1) ? ns : 0;
    if (const char* e = getenv("PIVB200_GENERIC_DIRECT")) if (atoi(e)) gp.n_stages = 0;      // A/B switch: direct sums
}

inline int generic_launch(const GenericParams& gp_in, cudaStream_t s) {
    GenericParams gp = gp_in;
    generic_set_plan(gp);
    const size_t smem = generic_smem_bytes(gp.wind);
    const int threads = generic_threads(gp.wind);
    bool big = threads > 256;
    for (int i = 0; i < gp.n_stages; ++i) big = big || gp.radix[i] > 8;
    const bool gz = gp.wind > kGenericMaxSharedWindow;
    auto kern = gz ? generic_corr_kernel<true, true> : (big ? generic_corr_kernel<true> : generic_corr_kernel<false>);
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (err != cudaSuccess) return static_cast<int>(err);
    int dev = 0, sms = 0, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
    long long grid = static_cast<long long>(sms) * (per_sm > 0 ? per_sm : 1);
    if (grid > gp.n_windows) grid = gp.n_windows;
    gp.zglobal = nullptr;
    if (gz) {           // stream-ordered scratch: grid slabs of w (w + 1) complex values (78 MB for 256 px on 148 SMs)
        err = cudaMallocAsync(reinterpret_cast<void**>(&gp.zglobal),
                              static_cast<size_t>(grid) * gp.wind * (gp.wind | 1) * sizeof(float2), s);
        if (err != cudaSuccess) return static_cast<int>(err);
    }
    kern<<<static_cast<unsigned>(grid), threads, smem, s>>>(gp);
    count_launch();
    err = cudaGetLastError();
    if (gz) cudaFreeAsync(gp.zglobal, s);
    return static_cast<int>(err);
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
    const int map_cols = generic_map_cols(wind);
    const long long map_elems = static_cast<long long>(wind) * map_cols;
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
            maps, n, wind, map_cols, p.validate, p.val_ratio, 3, du + first, dv + first, p.mask ? p.mask + first : nullptr);
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
