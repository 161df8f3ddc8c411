// Per-thread arithmetic of the pair-packed correlation pipeline (piv_soa.cuh), kept free of any
// shared-memory / warp logic so that the same code is exercised on the host by tests/micro/soa_math_host.cpp.
//
// A window of W x W pixels is handled by H = W/2 lanes.  Every register array is `float2 x[W]`:
// element i = (x[2i], x[2i+1]) = (real pair, imaginary pair), a pair being the same quantity of the two
// transforms the lane runs side by side (fft_soa.cuh).
//
//   rows     lane l owns the ADJACENT window rows 2l (.x) and 2l+1 (.y).  A real row of W samples is
//            transformed as the H-point complex sequence z[m] = r[2m] + i r[2m+1] plus one split step;
//            the lane leaves 2 R[k], k = 0..H-1, in element pos(k), the real bins 0 and W/2 sharing
//            element 0 as (2 R[0], 2 R[W/2]).
//   columns  lane c owns spectrum column c.  The pair is now (even rows, odd rows) of that column
//            (decimation in time): one H-point transform of both halves, then one radix-2 step across
//            the pair gives (Y[q], Y[q+H]) in element pos(q).
//   product  P = conj(A^) B^ element by element (column 0 = two packed real columns is separated by
//            the caller, piv_soa.cuh).
//   columns^-1  radix-2 step across the pair first (decimation in frequency), then one inverse H-point
//            transform of both halves: element pos2(m) = rows (2m, 2m+1) of the half spectrum.
//   rows^-1  Hermitian half spectrum of rows (2l, 2l+1) -> H-point complex sequence -> inverse
//            transform: element pos(m) = samples (2m, 2m+1) of both rows.
// Nothing is normalised: the result is 4 W^2 times the circular cross-correlation (the factor 2 of each
// frame's row step times W^2 of the two unnormalised inverse transforms).
//
// Replaces torch.fft.rfft2 / irfft2 + the conjugate product of correalte_fft (PIVbackend.py:249-257).
#pragma once
#include "fft_soa.cuh"

namespace pivb200 {

template <int W>
struct SoaMath {
    static constexpr int H = W / 2;
    using F = Fft2<H>;
    __host__ __device__ static constexpr int pos(int k) { return F::pos(k); }
    __host__ __device__ static constexpr int pos2(int k) { return F::pos(F::pos(k)); }
    // the bin that Fft2 leaves in element e
    __host__ __device__ static constexpr int posinv(int e) {
        for (int k = 0; k < H; ++k)
            if (F::pos(k) == e) return k;
        return -1;
    }

    // The pipeline has ONE transform body (Fft2<H> on the natural slots, forward sign) that every phase shares
    // -- the fused kernel runs it from a single call site inside a run-time loop, which keeps its per-job code
    // inside the instruction cache (profiles/r02c_icache_stream.txt).  An inverse transform is the same body on
    // data whose real pairs sit in the ODD and imaginary pairs in the EVEN slots ("swapped" below): the result
    // comes out swapped as well.

    // ---- rows, forward, AFTER the transform.  in: element pos(k) = Z[k].  out: element pos(c) = 2 R[c]
    static __host__ __device__ __forceinline__ void row_split(float2 (&x)[W]) {
        {   // k = 0 (and W/2): both real
            const float2 zr = x[0], zi = x[1];
            x[0] = pmuls(padd(zr, zi), 2.0f);
            x[1] = pmuls(psub(zr, zi), 2.0f);
        }
        {   // k = H/2: 2 conj(Z)
            constexpr int e = pos(H / 2);
            x[2 * e] = pmuls(x[2 * e], 2.0f);
            x[2 * e + 1] = pmuls(x[2 * e + 1], -2.0f);
        }
        static_for<1, H / 2>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            constexpr int ea = pos(k), eb = pos(H - k);
            constexpr float c = float(ct_cos2pi(k, W)), s = float(ct_sin2pi(k, W));
            const float2 ar = x[2 * ea], ai = x[2 * ea + 1], br = x[2 * eb], bi = x[2 * eb + 1];
            const float2 sr = padd(ar, br), si = psub(ai, bi), dr = psub(ar, br), di = padd(ai, bi);
            const float2 tr = pfmas(di, c, pmuls(dr, -s));          // -s dr + c di
            const float2 ti = pfmas(di, -s, pmuls(dr, -c));         // -c dr - s di
            x[2 * ea] = padd(sr, tr);
            x[2 * ea + 1] = padd(si, ti);
            x[2 * eb] = psub(sr, tr);
            x[2 * eb + 1] = psub(ti, si);
        });
    }

    // ---- columns, forward, AFTER the transform.  in: element pos(q) = (E[q], O[q]) (even rows, odd rows).
    //      out: element pos(q) = (Y[q], Y[q+H])
    static __host__ __device__ __forceinline__ void col_glue_fwd(float2 (&x)[W]) {
        static_for<0, H>([&](auto qc) {
            constexpr int q = decltype(qc)::value;
            constexpr int e = pos(q);
            const float2 re = x[2 * e], im = x[2 * e + 1];
            float orr, oi;                                          // w^q O[q]
            if constexpr (q == 0) {
                orr = re.y; oi = im.y;
            } else if constexpr (2 * q == H) {                      // -i
                orr = im.y; oi = -re.y;
            } else {
                constexpr float c = float(ct_cos2pi(q, W)), s = float(ct_sin2pi(q, W));
                orr = fmaf(re.y, c, im.y * s);
                oi = fmaf(im.y, c, -(re.y * s));
            }
            x[2 * e] = make_float2(re.x + orr, re.x - orr);
            x[2 * e + 1] = make_float2(im.x + oi, im.x - oi);
        });
    }

    // ---- P = conj(A) B, element by element
    static __host__ __device__ __forceinline__ void product(const float2 (&a)[W], float2 (&b)[W]) {
        static_for<0, H>([&](auto ec) {
            constexpr int e = decltype(ec)::value;
            const float2 ar = a[2 * e], ai = a[2 * e + 1], br = b[2 * e], bi = b[2 * e + 1];
            b[2 * e] = pfma(ar, br, pmul(ai, bi));
            b[2 * e + 1] = pfma(ar, bi, pneg(pmul(ai, br)));
        });
    }

    // ---- columns, inverse, BEFORE the transform.  in: element pos(q) = (P[q], P[q+H]).
    //      out: element q = (u[q], v[q]) SWAPPED (radix-2 step across the pair: decimation in frequency)
    static __host__ __device__ __forceinline__ void col_glue_inv(float2 (&x)[W]) {
        float2 y[W];
        static_for<0, H>([&](auto qc) {
            constexpr int q = decltype(qc)::value;
            constexpr int e = pos(q);
            const float2 re = x[2 * e], im = x[2 * e + 1];
            const float ur = re.x + re.y, ui = im.x + im.y, dr = re.x - re.y, di = im.x - im.y;
            float vr, vi;                                           // conj(w^q) d
            if constexpr (q == 0) {
                vr = dr; vi = di;
            } else if constexpr (2 * q == H) {                      // +i
                vr = -di; vi = dr;
            } else {
                constexpr float c = float(ct_cos2pi(q, W)), s = float(ct_sin2pi(q, W));
                vr = fmaf(dr, c, -(di * s));
                vi = fmaf(di, c, dr * s);
            }
            y[2 * q] = make_float2(ui, vi);
            y[2 * q + 1] = make_float2(ur, vr);
        });
        static_for<0, W>([&](auto ic) { constexpr int i = decltype(ic)::value; x[i] = y[i]; });
    }

    // ---- rows, inverse, BEFORE the transform, on SWAPPED slots.  in: element c = G[c] of the lane's two rows
    //      (element 0 = (G[0], G[W/2]), both real).  out: element k = Z[k], swapped
    static __host__ __device__ __forceinline__ void row_presplit(float2 (&x)[W]) {
        {
            const float2 g0 = x[1], gh = x[0];
            x[1] = padd(g0, gh);
            x[0] = psub(g0, gh);
        }
        {
            constexpr int e = H / 2;
            x[2 * e + 1] = pmuls(x[2 * e + 1], 2.0f);
            x[2 * e] = pmuls(x[2 * e], -2.0f);
        }
        static_for<1, H / 2>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            constexpr int ea = k, eb = H - k;
            constexpr float c = float(ct_cos2pi(k, W)), s = float(ct_sin2pi(k, W));
            const float2 ar = x[2 * ea + 1], ai = x[2 * ea], br = x[2 * eb + 1], bi = x[2 * eb];
            const float2 sr = padd(ar, br), si = psub(ai, bi), dr = psub(ar, br), di = padd(ai, bi);
            const float2 tr = pfmas(di, -c, pmuls(dr, -s));         // -s dr - c di
            const float2 ti = pfmas(di, -s, pmuls(dr, c));          //  c dr - s di
            x[2 * ea + 1] = padd(sr, tr);
            x[2 * ea] = padd(si, ti);
            x[2 * eb + 1] = psub(sr, tr);
            x[2 * eb] = psub(ti, si);
        });
    }

    // ---- composites (host test, documentation of the data flow)
    // rows forward: x[j] = samples j of the lane's two rows -> element pos(c) = 2 R[c]
    static __host__ __device__ __forceinline__ void row_forward(float2 (&x)[W]) { F::run(x); row_split(x); }
    // columns forward: element t = (rows 2t, 2t+1) of the column -> element pos(q) = (Y[q], Y[q+H])
    static __host__ __device__ __forceinline__ void col_forward(float2 (&x)[W]) { F::run(x); col_glue_fwd(x); }
    // columns inverse: element pos(q) = (P[q], P[q+H]) -> element pos(m) = rows (2m, 2m+1), SWAPPED
    static __host__ __device__ __forceinline__ void col_inverse(float2 (&x)[W]) { col_glue_inv(x); F::run(x); }
    // rows inverse: element c = G[c] SWAPPED -> element pos(m) = samples (2m, 2m+1) of both rows, SWAPPED
    static __host__ __device__ __forceinline__ void row_inverse(float2 (&x)[W]) { row_presplit(x); F::run(x); }
};

}  // namespace pivb200
