// In-register complex FFTs (N = 2..64) with compile-time twiddles for sm_100a.
//
// One thread owns a whole N-point transform in `float2 x[N]`; every index below is a
// compile-time constant, so the array lives in registers and every twiddle is an FFMA/FMUL
// immediate.  Decimation in frequency, radix 8/4/2, no data reordering: the transform leaves
// bin k at x[Dif<N>::pos(k)] (digit-reversed), which costs nothing because consumers also
// address the array statically.  Only the FORWARD transform (e^{-2 pi i jk/N}) is generated;
// an inverse transform is the same code run on (im, re)-swapped data (free register renaming).
//
// This is what replaces torch.fft.rfft2 / irfft2 of the reference (PIVbackend.py:255-256).
#pragma once
#include <cuda_runtime.h>
#include <type_traits>

namespace pivb200 {

// ----------------------------------------------------------------------------------------
// compile-time loop
// ----------------------------------------------------------------------------------------
template <int I, int N, class F>
__host__ __device__ __forceinline__ void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(static_cast<F&&>(f));
    }
}

// ----------------------------------------------------------------------------------------
// compile-time trigonometry (double Taylor series on [0, pi/4], octant symmetry)
// ----------------------------------------------------------------------------------------
constexpr double kPi = 3.14159265358979323846264338327950288;

__host__ __device__ constexpr double ct_sin_small(double x) {
    double x2 = x * x, term = x, sum = x;
    for (int i = 1; i < 14; ++i) {
        term *= -x2 / double((2 * i) * (2 * i + 1));
        sum += term;
    }
    return sum;
}
__host__ __device__ constexpr double ct_cos_small(double x) {
    double x2 = x * x, term = 1.0, sum = 1.0;
    for (int i = 1; i < 14; ++i) {
        term *= -x2 / double((2 * i - 1) * (2 * i));
        sum += term;
    }
    return sum;
}
// cos(2 pi k / n), n a power of two >= 4
__host__ __device__ constexpr double ct_cos2pi(int k, int n) {
    k = ((k % n) + n) % n;
    if (k > n / 2) k = n - k;               // cos(2pi - t) = cos t
    bool neg = false;
    if (k > n / 4) { k = n / 2 - k; neg = true; }   // cos(pi - t) = -cos t
    double r = 0.0;
    if (k == 0) r = 1.0;
    else if (4 * k == n) r = 0.0;
    else if (8 * k <= n) r = ct_cos_small(2.0 * kPi * double(k) / double(n));
    else r = ct_sin_small(2.0 * kPi * (double(n) / 4.0 - double(k)) / double(n));
    return neg ? -r : r;
}
__host__ __device__ constexpr double ct_sin2pi(int k, int n) { return ct_cos2pi(k - n / 4, n); }

constexpr float kSqrtHalf = 0.70710678118654752440f;

// Complex add / subtract as ONE packed FP32 instruction (sm_100a FADD2 / FFMA2: two lanes of a 64-bit
// register pair per issue slot -- same FLOP rate as the scalar forms, half the issue slots; the fused
// kernels are issue bound).  Host builds (unit tests of the index logic) use plain arithmetic.
__host__ __device__ __forceinline__ float2 cadd(float2 a, float2 b) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
    return __fadd2_rn(a, b);
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}
__host__ __device__ __forceinline__ float2 csub(float2 a, float2 b) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
    return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a);
#else
    return make_float2(a.x - b.x, a.y - b.y);
#endif
}
// a * s (real s)
__host__ __device__ __forceinline__ float2 cscale(float2 a, float s) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
    return __fmul2_rn(a, make_float2(s, s));
#else
    return make_float2(a.x * s, a.y * s);
#endif
}
// element-wise a * b and a * b + c on pairs (FMUL2 / FFMA2); NOT complex products
__host__ __device__ __forceinline__ float2 cmul2(float2 a, float2 b) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
    return __fmul2_rn(a, b);
#else
    return make_float2(a.x * b.x, a.y * b.y);
#endif
}
__host__ __device__ __forceinline__ float2 cfma(float2 a, float2 b, float2 c) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
    return __ffma2_rn(a, b, c);
#else
    return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
// -i (a - b), computed directly as the rotated pair (two scalar subtractions)
__host__ __device__ __forceinline__ float2 csub_mi(float2 a, float2 b) { return make_float2(a.y - b.y, b.x - a.x); }

// a * e^{-2 pi i K / N}
template <int K, int N>
__host__ __device__ __forceinline__ float2 mul_tw(float2 a) {
    constexpr int k = ((K % N) + N) % N;
    if constexpr (k == 0) {
        return a;
    } else if constexpr (4 * k == N) {            // -i
        return make_float2(a.y, -a.x);
    } else if constexpr (2 * k == N) {            // -1
        return make_float2(-a.x, -a.y);
    } else if constexpr (4 * k == 3 * N) {        // +i
        return make_float2(-a.y, a.x);
    } else if constexpr (8 * k == N) {            // (1 - i)/sqrt2
        return cscale(make_float2(a.x + a.y, a.y - a.x), kSqrtHalf);
    } else if constexpr (8 * k == 3 * N) {        // (-1 - i)/sqrt2
        return cscale(make_float2(a.y - a.x, -a.x - a.y), kSqrtHalf);
    } else if constexpr (8 * k == 5 * N) {        // (-1 + i)/sqrt2
        return cscale(make_float2(-a.x - a.y, a.x - a.y), kSqrtHalf);
    } else if constexpr (8 * k == 7 * N) {        // (1 + i)/sqrt2
        return cscale(make_float2(a.x - a.y, a.x + a.y), kSqrtHalf);
    } else {
        constexpr float c = float(ct_cos2pi(k, N));
        constexpr float s = float(-ct_sin2pi(k, N));
        return make_float2(fmaf(a.x, c, -a.y * s), fmaf(a.x, s, a.y * c));
    }
}

// ----------------------------------------------------------------------------------------
// small DFTs, natural order in and out, forward sign
// ----------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void dft2(float2& a0, float2& a1) {
    float2 t = csub(a0, a1);
    a0 = cadd(a0, a1);
    a1 = t;
}
__host__ __device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
    const float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3);
    const float2 t3 = csub_mi(a1, a3);              // -i (a1 - a3)
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    a1 = cadd(t1, t3);                              // t1 - i (a1 - a3)
    a3 = csub(t1, t3);                              // t1 + i (a1 - a3)
}
template <int R>
__host__ __device__ __forceinline__ void dft_small(float2 (&a)[R]) {
    if constexpr (R == 2) {
        dft2(a[0], a[1]);
    } else if constexpr (R == 4) {
        dft4(a[0], a[1], a[2], a[3]);
    } else {
        static_assert(R == 8, "radix");
        float2 b0 = cadd(a[0], a[4]), d0 = csub(a[0], a[4]);
        float2 b1 = cadd(a[1], a[5]);
        float2 b2 = cadd(a[2], a[6]), d2 = csub_mi(a[2], a[6]);                 // * e^{-2 pi i 2/8} = -i
        float2 b3 = cadd(a[3], a[7]);
        const float2 u = csub(a[1], a[5]), v = csub(a[3], a[7]);
        float2 d1 = cscale(make_float2(u.x + u.y, u.y - u.x), kSqrtHalf);        // * (1 - i)/sqrt2
        float2 d3 = cscale(make_float2(v.y - v.x, -v.x - v.y), kSqrtHalf);       // * (-1 - i)/sqrt2
        dft4(b0, b1, b2, b3);       // X0 X2 X4 X6
        dft4(d0, d1, d2, d3);       // X1 X3 X5 X7
        a[0] = b0; a[2] = b1; a[4] = b2; a[6] = b3;
        a[1] = d0; a[3] = d1; a[5] = d2; a[7] = d3;
    }
}

template <int N>
__host__ __device__ constexpr int pick_radix() {
    return (N == 64 || N == 32 || N == 8) ? 8 : ((N % 4 == 0) ? 4 : 2);
}

// ----------------------------------------------------------------------------------------
// Register views: the transform addresses logical element i at physical slot V::at(i).  Since every
// index is a compile-time constant a permuted view costs nothing, which is what lets a transform
// consume data that a previous transform left in digit-reversed order (no reordering pass).
// ----------------------------------------------------------------------------------------
struct ViewId {
    __host__ __device__ static constexpr int at(int i) { return i; }
};

// ----------------------------------------------------------------------------------------
// Dif<N, S, O, T, V>: forward FFT of the N logical elements O + j*S of a T-element register array.
// Bin k ends up at logical position O + S * pos(k), i.e. physical slot V::at(O + S * pos(k)).
// ----------------------------------------------------------------------------------------
template <int N, int S, int O, int T, class V = ViewId>
struct Dif {
    static constexpr int R = pick_radix<N>();
    static constexpr int M = N / R;

    __host__ __device__ static constexpr int pos(int k) {
        if constexpr (M == 1) return k;
        else return M * (k % R) + Dif<M, S, O, T, V>::pos(k / R);
    }
    // physical slot of bin k of a whole-array transform (S = 1, O = 0)
    __host__ __device__ static constexpr int slot(int k) { return V::at(pos(k)); }

    __host__ __device__ static __forceinline__ void run(float2 (&x)[T]) {
        static_for<0, M>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            float2 a[R];
            static_for<0, R>([&](auto qc) {
                constexpr int q = decltype(qc)::value;
                a[q] = x[V::at(O + (j + M * q) * S)];
            });
            dft_small<R>(a);
            static_for<0, R>([&](auto pc) {
                constexpr int p = decltype(pc)::value;
                if constexpr (M > 1) x[V::at(O + (j + M * p) * S)] = mul_tw<j * p, N>(a[p]);
                else x[V::at(O + (j + M * p) * S)] = a[p];
            });
        });
        if constexpr (M > 1) {
            static_for<0, R>([&](auto pc) {
                constexpr int p = decltype(pc)::value;
                Dif<M, S, O + M * p * S, T, V>::run(x);
            });
        }
    }
};

template <int N>
using Fft = Dif<N, 1, 0, N>;

// View that reads logical element j from the slot where Fft<N> left bin j: FftRev<N> transforms
// the OUTPUT of an Fft<N> in place (bin k of the second transform lands at FftRev<N>::slot(k)).
template <int N>
struct ViewRev {
    __host__ __device__ static constexpr int at(int i) { return Fft<N>::pos(i); }
};
template <int N>
using FftRev = Dif<N, 1, 0, N, ViewRev<N>>;

}  // namespace pivb200
