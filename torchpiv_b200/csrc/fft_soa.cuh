// Pair-packed in-register FFTs for sm_100a ("SoA" layout): one thread runs TWO independent N-point
// complex transforms at once.  A `float2` holds the same real quantity of the two transforms (.x =
// transform 0, .y = transform 1); the register array `float2 x[2N]` keeps the real parts at even and
// the imaginary parts at odd indices (element i = (x[2i], x[2i+1])).  Every butterfly, every twiddle
// product and every multiplication by +-i is then a packed FADD2 / FMUL2 / FFMA2 with a 32-bit
// immediate that serves both halves (sm_100a broadcasts the immediate), or a register renaming --
// half the issue slots of the (re, im)-packed transforms in fft_regs.cuh, whose rotations and
// twiddle products have to fall back to scalar instructions.
//
// Decimation in frequency, radix 8/4/2, compile-time twiddles, no data reordering: bin k of an N-point
// transform ends up at element Dif2<N>::pos(k).  Only the FORWARD transform is generated; an inverse
// transform is the same code addressed through an accessor that swaps the roles of the real and
// imaginary slots (free renaming).
//
// This is what replaces torch.fft.rfft2 / irfft2 of the reference (PIVbackend.py:255-256) in
// piv_soa.cuh (32 and 16 px windows).
#pragma once
#include <cuda_runtime.h>

#include "fft_regs.cuh"

namespace pivb200 {

// ----------------------------------------------------------------------------------------
// packed pair arithmetic (host fallbacks keep the index logic testable on the CPU)
// ----------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ float2 padd(float2 a, float2 b) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
    return __fadd2_rn(a, b);
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}
__host__ __device__ __forceinline__ float2 pneg(float2 a) { return make_float2(-a.x, -a.y); }
// a - b: the negation folds into the FADD2 source modifier
__host__ __device__ __forceinline__ float2 psub(float2 a, float2 b) { return padd(a, pneg(b)); }
// a * s, a * s + c with one scalar constant for both halves (FMUL2 / FFMA2 immediate form)
__host__ __device__ __forceinline__ float2 pmuls(float2 a, float s) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
    return __fmul2_rn(a, make_float2(s, s));
#else
    return make_float2(a.x * s, a.y * s);
#endif
}
__host__ __device__ __forceinline__ float2 pfmas(float2 a, float s, float2 c) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
    return __ffma2_rn(a, make_float2(s, s), c);
#else
    return make_float2(fmaf(a.x, s, c.x), fmaf(a.y, s, c.y));
#endif
}
__host__ __device__ __forceinline__ float2 pmul(float2 a, float2 b) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
    return __fmul2_rn(a, b);
#else
    return make_float2(a.x * b.x, a.y * b.y);
#endif
}
__host__ __device__ __forceinline__ float2 pfma(float2 a, float2 b, float2 c) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
    return __ffma2_rn(a, b, c);
#else
    return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}

// a complex PAIR: two complex numbers, real parts in re, imaginary parts in im
struct Cp {
    float2 re, im;
};
__host__ __device__ __forceinline__ Cp cp_add(Cp a, Cp b) { return Cp{padd(a.re, b.re), padd(a.im, b.im)}; }
__host__ __device__ __forceinline__ Cp cp_sub(Cp a, Cp b) { return Cp{psub(a.re, b.re), psub(a.im, b.im)}; }
// a - i b, a + i b
__host__ __device__ __forceinline__ Cp cp_sub_i(Cp a, Cp b) { return Cp{padd(a.re, b.im), psub(a.im, b.re)}; }
__host__ __device__ __forceinline__ Cp cp_add_i(Cp a, Cp b) { return Cp{psub(a.re, b.im), padd(a.im, b.re)}; }

// a * e^{-2 pi i K / N}
template <int K, int N>
__host__ __device__ __forceinline__ Cp cp_mul_tw(Cp a) {
    constexpr int k = ((K % N) + N) % N;
    if constexpr (k == 0) {
        return a;
    } else if constexpr (4 * k == N) {            // -i
        return Cp{a.im, pneg(a.re)};
    } else if constexpr (2 * k == N) {            // -1
        return Cp{pneg(a.re), pneg(a.im)};
    } else if constexpr (4 * k == 3 * N) {        // +i
        return Cp{pneg(a.im), a.re};
    } else {
        constexpr float c = float(ct_cos2pi(k, N));
        constexpr float s = float(ct_sin2pi(k, N));
        // (re + i im)(c - i s) = (re c + im s) + i (im c - re s)
        return Cp{pfmas(a.re, c, pmuls(a.im, s)), pfmas(a.im, c, pmuls(a.re, -s))};
    }
}

// ----------------------------------------------------------------------------------------
// small DFTs on complex pairs, natural order in and out, forward sign
// ----------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void cp_dft4(Cp& a0, Cp& a1, Cp& a2, Cp& a3) {
    const Cp t0 = cp_add(a0, a2), t1 = cp_sub(a0, a2), t2 = cp_add(a1, a3), t3 = cp_sub(a1, a3);
    a0 = cp_add(t0, t2);
    a2 = cp_sub(t0, t2);
    a1 = cp_sub_i(t1, t3);
    a3 = cp_add_i(t1, t3);
}
template <int R>
__host__ __device__ __forceinline__ void cp_dft_small(Cp (&a)[R]) {
    if constexpr (R == 2) {
        const Cp t = cp_sub(a[0], a[1]);
        a[0] = cp_add(a[0], a[1]);
        a[1] = t;
    } else if constexpr (R == 4) {
        cp_dft4(a[0], a[1], a[2], a[3]);
    } else {
        static_assert(R == 8, "radix");
        Cp b0 = cp_add(a[0], a[4]), d0 = cp_sub(a[0], a[4]);
        Cp b1 = cp_add(a[1], a[5]);
        Cp b2 = cp_add(a[2], a[6]);
        Cp b3 = cp_add(a[3], a[7]);
        const Cp u = cp_sub(a[1], a[5]), w = cp_sub(a[2], a[6]), v = cp_sub(a[3], a[7]);
        Cp d1 = Cp{pmuls(padd(u.re, u.im), kSqrtHalf), pmuls(psub(u.im, u.re), kSqrtHalf)};     // * (1 - i)/sqrt2
        Cp d2 = Cp{w.im, pneg(w.re)};                                                          // * -i
        Cp d3 = Cp{pmuls(psub(v.im, v.re), kSqrtHalf), pmuls(padd(v.re, v.im), -kSqrtHalf)};    // * (-1 - i)/sqrt2
        cp_dft4(b0, b1, b2, b3);       // X0 X2 X4 X6
        cp_dft4(d0, d1, d2, d3);       // X1 X3 X5 X7
        a[0] = b0; a[2] = b1; a[4] = b2; a[6] = b3;
        a[1] = d0; a[3] = d1; a[5] = d2; a[7] = d3;
    }
}

// ----------------------------------------------------------------------------------------
// Accessors: logical element i of the transform lives at x[A::re(i)] / x[A::im(i)]
// ----------------------------------------------------------------------------------------
struct AccFwd {          // forward transform of elements in natural slots
    __host__ __device__ static constexpr int re(int i) { return 2 * i; }
    __host__ __device__ static constexpr int im(int i) { return 2 * i + 1; }
};
struct AccInv {          // inverse transform (roles of re / im swapped), natural slots
    __host__ __device__ static constexpr int re(int i) { return 2 * i + 1; }
    __host__ __device__ static constexpr int im(int i) { return 2 * i; }
};

// Dif2<N, S, O, A>: forward FFT of the N logical elements O + j*S.  Bin k ends up at logical element
// O + S * pos(k).
template <int N, int S, int O, class A>
struct Dif2 {
    static constexpr int R = pick_radix<N>();
    static constexpr int M = N / R;

    __host__ __device__ static constexpr int pos(int k) {
        if constexpr (M == 1) return k;
        else return M * (k % R) + Dif2<M, S, O, A>::pos(k / R);
    }

    template <int T>
    __host__ __device__ static __forceinline__ void run(float2 (&x)[T]) {
        static_for<0, M>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            Cp a[R];
            static_for<0, R>([&](auto qc) {
                constexpr int q = decltype(qc)::value;
                constexpr int e = O + (j + M * q) * S;
                a[q] = Cp{x[A::re(e)], x[A::im(e)]};
            });
            cp_dft_small<R>(a);
            static_for<0, R>([&](auto pc) {
                constexpr int p = decltype(pc)::value;
                constexpr int e = O + (j + M * p) * S;
                Cp o = a[p];
                if constexpr (M > 1) o = cp_mul_tw<j * p, N>(a[p]);
                x[A::re(e)] = o.re;
                x[A::im(e)] = o.im;
            });
        });
        if constexpr (M > 1) {
            static_for<0, R>([&](auto pc) {
                constexpr int p = decltype(pc)::value;
                Dif2<M, S, O + M * p * S, A>::run(x);
            });
        }
    }
};

template <int N>
using Fft2 = Dif2<N, 1, 0, AccFwd>;          // forward, natural input slots, bin k at element pos(k)
template <int N>
using Ifft2 = Dif2<N, 1, 0, AccInv>;         // inverse (unnormalised), natural input slots

// inverse transform whose logical input element j sits where Fft2<N> left bin j (element pos(j)):
// output sample m ends up at element pos(pos(m))
template <int N>
struct AccInvRev {
    __host__ __device__ static constexpr int re(int i) { return 2 * Fft2<N>::pos(i) + 1; }
    __host__ __device__ static constexpr int im(int i) { return 2 * Fft2<N>::pos(i); }
};
template <int N>
using Ifft2Rev = Dif2<N, 1, 0, AccInvRev<N>>;
template <int N>
__host__ __device__ constexpr int pos2(int m) { return Fft2<N>::pos(Fft2<N>::pos(m)); }

}  // namespace pivb200
