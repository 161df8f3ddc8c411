// Fused PIV pass kernel for sm_100a: window extraction (TMA 3-D tile loads of the uint8 frames)
// -> CWS bilinear / DWS integer window shift -> 2-D cross-correlation by in-register FFTs ->
// fft-shifted peak search, 3-point log-Gaussian sub-pixel fit, peak-ratio validation and predictor
// glue.  Only the displacement vectors leave the SM.
//
// Replaces, for one pass, the reference's eager op chain
//   moving_window_array (PB:220-247) -> [biliniar_interpolation_CWS | interpolation_DWS]
//   (PB:147-216) -> correalte_fft (PB:249-257) -> corr - amin (PB:518/724/796) ->
//   correlation_to_displacement + peak2peak_secondpeak (PB:346-422) -> replacement logic
//   (PB:728-738 / 800-810).
//
// Execution model: persistent CTAs (one per SM) of NWARPS independent warps.  A warp owns a "job" of
// NW = 64 / W interrogation windows (W = 64/32/16); H = W/2 lanes work on one window and every lane
// runs exactly one W-point complex FFT per phase, entirely in registers:
//   Ra  rows (l, l+H) of frame a packed as z = row_l + i row_{l+H}; FFT; the two real-row spectra are
//       separated in registers and their half spectra (bins 0..H, bins 0 and H packed into one
//       complex number) go to the exchange buffer X [W][H] in shared memory
//   Ca  lane l transforms column l of X -> A^[.][l]; the spectrum is PARKED in tensor memory (TMEM is
//       lane-private scratch here: tcgen05.st / tcgen05.ld, no MMA), which frees both the registers
//       and the exchange buffer
//   Rb  same as Ra for frame b (its tile is fetched by TMA while Ca computes)
//   Cb  column FFT of b, product conj(A^) B^ in place against the parked spectrum, inverse column
//       FFT on the digit-reversed registers (FftRev: no reordering) -> Q [W][H] in the same buffer
//   I   two Hermitian rows (l, l+H) of Q packed into one inverse FFT -> two rows of the correlation
//       map in registers
//   E   min / argmax / neighbours / second peak on the fft-shifted map -> u, v, mask
// Shared memory per window is W*(H+1)*8 bytes (16.5 KB for W = 64) -- half of a complex W x W
// buffer -- and it also receives the TMA tiles, so 12 (W=64) to 24 (W=16) warps are resident per SM.
//
// Loaders (template parameter LOADER, piv_params.h): LD_FRAME_ALN (unshifted windows on a 16-px aligned
// grid: TMA box = window), LD_FRAME_INT (integer shifts: box 16 B wider, rows re-aligned in registers),
// LD_FRAME_CWS (float shifts: separable bilinear taps, packed FP32), LD_EXPL_* (materialised windows, function-
// level API) and the experimental LD_FRAME_TC (64 px first pass whose row transform Ra / Rb runs on the
// tensor cores: groups of four warps stage fp16 operand tiles, one thread issues tcgen05.mma, the spectra
// come back from TMEM already in the X layout).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cfloat>
#include <cstdint>

#include "fft_regs.cuh"
#include "piv_params.h"

namespace pivb200 {

constexpr int kShiftClamp = 1 << 20;     // |shift| beyond this is clamped (documented deviation)

template <int W>
struct Geo {
    static_assert(W == 16 || W == 32 || W == 64, "window size");
    static constexpr int NW = 64 / W;                 // windows per warp job
    static constexpr int HALF = W / 2;                // lanes per window
    static constexpr int LOGW = (W == 64) ? 6 : (W == 32 ? 5 : 4);
    static constexpr int PX = W / 2 + 1;              // pitch of X / Q rows (float2); odd -> conflict free
    static constexpr int PC = W + 1;                  // pitch of map rows (float)
    static constexpr int XB = W * PX * 8;
    static constexpr int MAPB = W * PC * 4;
    // Bank staggering: the windows of one warp keep their X / Q / map data at different offsets inside
    // their buffers (0, 64, 32, 96 bytes), so that lanes of different windows that execute the same
    // shared-memory instruction hit different banks.  The TMA tile always starts at the buffer base.
    static constexpr int STAGGER = (NW > 1) ? 96 : 0;
    __host__ __device__ static constexpr int doff(int wi) { return ((wi & 1) * 64) + ((wi >> 1) * 32); }
    static constexpr int REGION = (((XB > MAPB ? XB : MAPB) + STAGGER + 255) / 256) * 256;   // one window's buffer
    static constexpr float K = 4.0f * W * W;          // map = K * sum_x a(x) b(x+s)
    static constexpr int TCOLS = 2 * W;               // TMEM columns (32-bit) one warp parks
};

template <int W, int LOADER>
struct Tile {
    static constexpr bool kFrame = (LOADER == LD_FRAME_INT || LOADER == LD_FRAME_CWS || LOADER == LD_FRAME_ALN ||
                                    LOADER == LD_FRAME_TC);
    static constexpr bool kAligned = (LOADER == LD_FRAME_ALN || LOADER == LD_FRAME_TC);
    // TMA box.  The global start address of a box row must be 16-byte aligned (an unaligned x
    // coordinate faults as "illegal instruction" on sm_100a), so the box starts at the window's
    // x origin rounded DOWN to 16 and is 16 bytes wider than the bytes that are used; the row
    // loader re-aligns in registers by d = origin & 15.
    static constexpr int USED = (LOADER == LD_FRAME_CWS) ? W + 1 : W;    // bytes of a row that are read
    static constexpr int BX = kAligned ? W : W + 16;                     // box bytes per row
    static constexpr int BY = (LOADER == LD_FRAME_CWS) ? W + 1 : W;      // box rows
    static constexpr int SWZ = (BX == 32) ? 1 : ((BX == 64) ? 2 : 0);    // TMA swizzle: 32B / 64B / none
    static constexpr int BASE_ALIGN = (SWZ == 2) ? 512 : 256;            // the swizzle is a function of the shared-memory address
    static constexpr int TX = BX * BY;
    static_assert(!kFrame || TX <= Geo<W>::REGION, "the tile is staged inside the window's buffer");
    // byte offset of 16-byte chunk `chunk` of row `row`.  Row pitches of 80 / 48 bytes (and the 32- / 64-byte
    // swizzles for pitches 32 / 64) make 8 consecutive rows hit 8 distinct 16-byte bank groups, so the
    // per-lane LDS.128 row reads are bank-conflict free.
    __device__ static __forceinline__ int off(int row, int chunk) {
        if constexpr (SWZ == 1) return row * 32 + ((chunk ^ ((row >> 2) & 1)) << 4);
        else if constexpr (SWZ == 2) return row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4);
        else return row * BX + (chunk << 4);
    }
};

// Load one tile row (all BX bytes, LDS.128) and return the NOUT words that start at byte `d`
// (0..15) of the row: word-granular shift by a 2-level select network, then a funnel shift.
template <int W, int LOADER, int NOUT>
__device__ __forceinline__ void load_row_words(const unsigned char* tile, int row, int d,
                                               uint32_t (&out)[NOUT]) {
    using T = Tile<W, LOADER>;
    constexpr int NL = T::BX / 4;
    if constexpr (T::kAligned) {
        static_assert(NOUT == NL, "aligned tiles hold exactly the window");
        static_for<0, T::BX / 16>([&](auto cc) {
            constexpr int c = decltype(cc)::value;
            const uint4 q = *reinterpret_cast<const uint4*>(tile + T::off(row, c));
            out[4 * c] = q.x; out[4 * c + 1] = q.y; out[4 * c + 2] = q.z; out[4 * c + 3] = q.w;
        });
        return;
    }
    uint32_t L[NL + 3];
    static_for<0, T::BX / 16>([&](auto cc) {
        constexpr int c = decltype(cc)::value;
        const uint4 q = *reinterpret_cast<const uint4*>(tile + T::off(row, c));
        L[4 * c] = q.x; L[4 * c + 1] = q.y; L[4 * c + 2] = q.z; L[4 * c + 3] = q.w;
    });
    L[NL] = L[NL + 1] = L[NL + 2] = 0u;
    const bool s2 = (d & 8) != 0, s1 = (d & 4) != 0;
    const int sh = (d & 3) * 8;
    static_for<0, NL>([&](auto kc) { constexpr int k = decltype(kc)::value; L[k] = s2 ? L[k + 2] : L[k]; });
    static_for<0, NL>([&](auto kc) { constexpr int k = decltype(kc)::value; L[k] = s1 ? L[k + 1] : L[k]; });
    static_for<0, NOUT>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        out[k] = __funnelshift_r(L[k], L[k + 1], sh);
    });
}

// Shared memory of one warp (all offsets relative to the warp's slot)
template <int W, int LOADER>
struct Smem {
    using G = Geo<W>;
    using T = Tile<W, LOADER>;
    static constexpr int REG_OFF = 0;                                        // NW window buffers
    static constexpr int XW_OFF = REG_OFF + G::NW * G::REGION;               // float4 [NW][W]  (CWS column taps, each weight twice)
    static constexpr int XW = (LOADER == LD_FRAME_CWS) ? G::NW * W * 16 : 0;
    static constexpr int XF_OFF = XW_OFF + XW;                               // int    [NW][W]
    static constexpr int XF = (LOADER == LD_FRAME_CWS) ? G::NW * W * 4 : 0;
    static constexpr int TD_OFF = ((XF_OFF + XF + 15) / 16) * 16;            // TileDesc [NW][2]
    static constexpr int TD = T::kFrame ? G::NW * 2 * 32 : 0;
    static constexpr int BAR_OFF = ((TD_OFF + TD + 7) / 8) * 8;
    static constexpr int TOTAL = BAR_OFF + 8;
    static constexpr int STRIDE = ((TOTAL + T::BASE_ALIGN - 1) / T::BASE_ALIGN) * T::BASE_ALIGN;
    static_assert(G::NW == 1 || T::BASE_ALIGN <= 256, "window buffers are 256-byte aligned inside a warp's slot");
    static constexpr int SMEM_MAX = 232448 - 1024;    // 227 KB opt-in limit per CTA minus the static allocation
    // tensor-core row transform (LD_FRAME_TC): CTA-wide area in front of the warp slots -- the DFT matrix
    // (hi and lo halves, 64 x 64 fp16 each) and, per group of four warps, two 128 x 64 fp16 operand tiles
    static constexpr bool kTC = (LOADER == LD_FRAME_TC);
    static constexpr int TC_GROUPS = 2;               // 2 x (128 accumulator + 128 parking) TMEM columns = 512
    static constexpr int TC_TW_OFF = 0;               // [2][8 KB]
    static constexpr int TC_A_OFF = 16384;            // [TC_GROUPS][2][16 KB]
    static constexpr int SHARED = kTC ? TC_A_OFF + TC_GROUPS * 2 * 16384 : 0;
    // warps per CTA: bounded by shared memory, by the register file (launch bounds) and by the 512
    // TMEM columns (4 lane quarters x 512 / TCOLS warps)
#ifndef PIVB200_W64_WARPS
#define PIVB200_W64_WARPS 12
#endif
#ifndef PIVB200_W32_WARPS
#define PIVB200_W32_WARPS 16
#endif
#ifndef PIVB200_W16_WARPS
#define PIVB200_W16_WARPS 24
#endif
    static constexpr int WARP_CAP = kTC ? 4 * TC_GROUPS
                                        : ((W == 64) ? PIVB200_W64_WARPS : (W == 32 ? PIVB200_W32_WARPS : PIVB200_W16_WARPS));
    static constexpr int TMEM_CAP = 4 * (512 / G::TCOLS);
    static constexpr int BY_SMEM = (SMEM_MAX - SHARED) / STRIDE;
    static constexpr int NWARPS0 = BY_SMEM < WARP_CAP ? BY_SMEM : WARP_CAP;
    static constexpr int NWARPS = NWARPS0 < TMEM_CAP ? NWARPS0 : TMEM_CAP;
    static constexpr int CTA_BYTES = SHARED + NWARPS * STRIDE;
    static_assert(!kTC || (W == 64 && NWARPS == 4 * TC_GROUPS), "the tensor-core row transform is built for 64 px windows");
};

// ----------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int x, int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(x), "r"(y), "r"(z)
        : "memory");
}

// Reductions over the LANES lanes of one window.  A whole warp (64 px windows) takes one CREDUX
// (redux.sync on f32 is new with sm_100a) instead of a five-level shuffle ladder; sub-warp groups keep
// the ladder: redux.sync with per-group member masks is serialised group by group (WARPSYNC.EXCLUSIVE),
// measured 30-60 % slower per pass on B200.
template <int LANES>
__device__ __forceinline__ float group_max(float v) {
    if constexpr (LANES == 32) {
        float r;
        asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
        return r;
    } else {
#pragma unroll
        for (int o = LANES / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
        return v;
    }
}
template <int LANES>
__device__ __forceinline__ float group_min(float v) {
    if constexpr (LANES == 32) {
        float r;
        asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
        return r;
    } else {
#pragma unroll
        for (int o = LANES / 2; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
        return v;
    }
}
template <int LANES>
__device__ __forceinline__ int group_min_int(int v) {
    if constexpr (LANES == 32) {
        return __reduce_min_sync(0xffffffffu, v);
    } else {
#pragma unroll
        for (int o = LANES / 2; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
        return v;
    }
}

// byte b of `word` as a float.  (A PRMT + FADD(-2^23) pair that avoids the quarter-rate XU pipe was
// measured: -2 % for the 64 px pass, +3 % for the 32 px passes -- not kept.)
__device__ __forceinline__ float u8f(uint32_t word, int b) {
    return static_cast<float>((word >> (8 * b)) & 0xffu);
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return max(lo, min(hi, v)); }

// ----------------------------------------------------------------------------------------
// window geometry
// ----------------------------------------------------------------------------------------
struct WinGeo {
    int pair, r0, c0;
};
__device__ __forceinline__ int fast_div(const FastDiv& d, int n) {
    const uint32_t un = static_cast<uint32_t>(n), t = __umulhi(d.M, un);
    return static_cast<int>((t + ((un - t) >> d.s1)) >> d.s2);
}
__device__ __forceinline__ WinGeo window_geo(const PassParams& p, int g) {
    const int n = p.n_rows * p.n_cols;
    WinGeo w;
    w.pair = fast_div(p.div_n, g);
    const int loc = g - w.pair * n;
    const int wr = fast_div(p.div_c, loc);
    w.r0 = wr * p.step;
    w.c0 = (loc - wr * p.n_cols) * p.step;
    return w;
}

// One axis of the reference's bilinear window shift (PB:163-170, 188-191), for one pixel coordinate:
// new = float32(coord) + v rounded in float32, taps floor/ceil, weights, exact-integer flag.
struct AxisTap {
    float w1, w0;   // (up - new), (new - down)
    int lo;         // floor(new) as an absolute pixel coordinate
    bool exact;     // up == down
};
__device__ __forceinline__ AxisTap cws_axis(int coord, float v) {
    const float nw = __fadd_rn(static_cast<float>(coord), v);
    const float dn = floorf(nw), up = ceilf(nw);
    AxisTap t;
    t.exact = (up == dn);
    t.w1 = __fsub_rn(up, nw);
    t.w0 = __fsub_rn(nw, dn);
    t.lo = static_cast<int>(dn);
    return t;
}
__device__ __forceinline__ float clamp_shift(float v) {
    return fminf(fmaxf(v, -static_cast<float>(kShiftClamp)), static_cast<float>(kShiftClamp));
}

// Everything the kernel needs to know about one (window, frame) tile; computed once per job by
// one lane (integer divisions!) and broadcast through shared memory.
struct __align__(16) TileDesc {
    int pair;       // pair index = TMA z coordinate
    int oy, ox;     // tile origin (absolute pixel) = window origin + integer part of the signed shift
    int d;          // byte offset of the origin inside the staged tile row (ox & 15 via TMA, 0 via gather)
    float vy, vx;   // signed float shift of this frame (CWS), 0 otherwise
    int r0, c0;     // window origin
};
static_assert(sizeof(TileDesc) == 32, "TileDesc layout");

// signed shift of `frame` (0 = a: minus, 1 = b: plus) and the tile origin it implies
template <int LOADER>
__device__ __forceinline__ void frame_origin(const PassParams& p, int g, const WinGeo& w, int frame,
                                             int& oy, int& ox, float& vy, float& vx) {
    if constexpr (LOADER == LD_FRAME_CWS) {
        const float sx = clamp_shift(p.sxf[g]), sy = clamp_shift(p.syf[g]);
        vx = frame ? sx : -sx;
        vy = frame ? sy : -sy;
        ox = w.c0 + static_cast<int>(floorf(vx));
        oy = w.r0 + static_cast<int>(floorf(vy));
    } else {
        int sx = 0, sy = 0;
        if (p.sxi != nullptr) {
            sx = clampi(p.sxi[g], -kShiftClamp, kShiftClamp);
            sy = clampi(p.syi[g], -kShiftClamp, kShiftClamp);
        }
        vx = vy = 0.f;
        ox = w.c0 + (frame ? sx : -sx);
        oy = w.r0 + (frame ? sy : -sy);
    }
}


// ----------------------------------------------------------------------------------------
// Tensor memory as lane-private scratch (tcgen05.alloc / st / ld; 32x32b shape: thread i of a warp
// addresses TMEM lane 32 * (warp % 4) + i, consecutive 32-bit columns)
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc_512(uint32_t smem_dst) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_dst) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_512(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 8 complex numbers = 16 columns
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float2 (&v)[8]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "f"(v[0].x), "f"(v[0].y), "f"(v[1].x), "f"(v[1].y), "f"(v[2].x), "f"(v[2].y), "f"(v[3].x),
        "f"(v[3].y), "f"(v[4].x), "f"(v[4].y), "f"(v[5].x), "f"(v[5].y), "f"(v[6].x), "f"(v[6].y), "f"(v[7].x),
        "f"(v[7].y)
        : "memory");
}
// one pair = 2 columns
__device__ __forceinline__ void tmem_st_pair(uint32_t taddr, float2 v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float2 (&v)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=f"(v[0].x), "=f"(v[0].y), "=f"(v[1].x), "=f"(v[1].y), "=f"(v[2].x), "=f"(v[2].y), "=f"(v[3].x),
          "=f"(v[3].y), "=f"(v[4].x), "=f"(v[4].y), "=f"(v[5].x), "=f"(v[5].y), "=f"(v[6].x), "=f"(v[6].y),
          "=f"(v[7].x), "=f"(v[7].y)
        : "r"(taddr)
        : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float2 (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=f"(v[0].x), "=f"(v[0].y), "=f"(v[1].x), "=f"(v[1].y), "=f"(v[2].x), "=f"(v[2].y), "=f"(v[3].x),
          "=f"(v[3].y), "=f"(v[4].x), "=f"(v[4].y), "=f"(v[5].x), "=f"(v[5].y), "=f"(v[6].x), "=f"(v[6].y),
          "=f"(v[7].x), "=f"(v[7].y), "=f"(v[8].x), "=f"(v[8].y), "=f"(v[9].x), "=f"(v[9].y), "=f"(v[10].x),
          "=f"(v[10].y), "=f"(v[11].x), "=f"(v[11].y), "=f"(v[12].x), "=f"(v[12].y), "=f"(v[13].x), "=f"(v[13].y),
          "=f"(v[14].x), "=f"(v[14].y), "=f"(v[15].x), "=f"(v[15].y)
        : "r"(taddr)
        : "memory");
}

// ----------------------------------------------------------------------------------------
// tcgen05.mma (kind::f16, single CTA): D[128 x 64, FP32, TMEM] (+)= A[128 x 64 fp16] * B[64 x 64 fp16]^T, both
// operands K-major in the canonical SWIZZLE_128B shared-memory layout (rows of 128 bytes, atoms of 8 rows,
// 16-byte chunk c of row r stored at chunk c ^ (r & 7)).  Bit layouts: cute/arch/mma_sm100_desc.hpp.
// ----------------------------------------------------------------------------------------
__host__ __device__ constexpr int sw128_offset(int row, int chunk) {
    return (row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4);
}
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return static_cast<uint64_t>((saddr >> 4) & 0x3fff)        // start address >> 4
           | (static_cast<uint64_t>(1) << 16)                  // leading byte offset: unused for swizzled K-major
           | (static_cast<uint64_t>(1024 >> 4) << 32)          // stride byte offset: 8 rows x 128 B
           | (static_cast<uint64_t>(1) << 46)                  // descriptor version (sm_100)
           | (static_cast<uint64_t>(2) << 61);                 // SWIZZLE_128B
}
// instruction descriptor: C = F32, A = B = F16, K-major, N = 64 (>> 3 at bit 17), M = 128 (>> 4 at bit 24)
constexpr uint32_t kUmmaIdesc = (1u << 4) | (static_cast<uint32_t>(64 >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(kUmmaIdesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// four uint8 of `w` -> two half2 (exact): byte b becomes the fp16 bit pattern 0x64bb = 1024 + b, minus 1024
__device__ __forceinline__ void u8x4_to_h2x2(uint32_t w, uint32_t& lo, uint32_t& hi) {
    const uint32_t a = __byte_perm(w, 0x64646464u, 0x4140), b = __byte_perm(w, 0x64646464u, 0x4342);
    const __half2 k = __floats2half2_rn(1024.f, 1024.f);
    const __half2 ha = __hsub2(*reinterpret_cast<const __half2*>(&a), k), hb = __hsub2(*reinterpret_cast<const __half2*>(&b), k);
    lo = *reinterpret_cast<const uint32_t*>(&ha);
    hi = *reinterpret_cast<const uint32_t*>(&hb);
}

// Parking order of a column spectrum.  After the forward column FFT bin q sits in register slot
// pos(q) (digit reversed); the inverse transform (same code, same register view) wants element q in
// slot q.  The product conj(A^) B^ therefore moves its result from slot pos(q) to slot q, which can be
// done in place as long as whole cycles of the permutation q -> pos(q) are handled together.  The
// parked spectrum is read back from TMEM in chunks of CH bins; ParkOrder packs whole cycles into
// chunks (first fit): 8 x 8 and 4 x 4 digit reversals are involutions (cycles of 1 or 2, CH = 8), the
// 8 x 4 reversal of the 32-point transform has two fixed points and six 5-cycles (CH = 16).
template <int W>
struct ParkOrder {
    static constexpr int CH = (W == 32) ? 16 : 8;
    static constexpr int NCH = W / CH;
    int bin[W];
    constexpr ParkOrder() : bin{} {
        bool used[W] = {};
        int fill[NCH] = {};
        for (int q = 0; q < W; ++q) {
            if (used[q]) continue;
            int len = 0;
            for (int r = q; !used[r]; r = Fft<W>::pos(r)) { used[r] = true; ++len; }
            int c = 0;
            while (fill[c] + len > CH) ++c;
            int r = q;
            for (int i = 0; i < len; ++i) { bin[c * CH + fill[c] + i] = r; r = Fft<W>::pos(r); }
            fill[c] += len;
        }
    }
};
template <int W>
__host__ __device__ constexpr int park_bin(int i) {
    constexpr ParkOrder<W> order{};
    return order.bin[i];
}

// Border window (slow path, rare): the reference addresses taps by FLAT index clamped to
// [0, H*W-1] (PB:172-180, 213-214), i.e. columns that leave the frame wrap into neighbouring rows.
// The whole warp gathers the tile byte by byte into the layout the TMA would have produced.
template <int W, int LOADER>
__device__ __noinline__ void gather_border_tile(const PassParams& p, const TileDesc dsc, unsigned char* tile,
                                                int frame, int lane) {
    using T = Tile<W, LOADER>;
    const unsigned char* f = (frame ? p.fb : p.fa) + dsc.pair * p.pair_stride;
    const int last_y = p.H - 1, last_x = p.Wf - 1;
#pragma unroll 2
    for (int e = lane; e < T::BY * T::USED; e += 32) {
        const int i = e / T::USED, jj = e - i * T::USED;
        int yy = dsc.oy + i, xx = dsc.ox + jj;
        if (xx < 0 || xx > last_x) {
            // column outside the frame: the flat index wraps into a neighbouring row
            long long flat = static_cast<long long>(yy) * p.Wf + xx;
            const long long last = static_cast<long long>(p.H) * p.Wf - 1;
            flat = flat < 0 ? 0 : (flat > last ? last : flat);
            const unsigned uf = static_cast<unsigned>(flat);      // H * W < 2^31 (checked on the host)
            yy = static_cast<int>(uf / static_cast<unsigned>(p.Wf));
            xx = static_cast<int>(uf - static_cast<unsigned>(yy) * static_cast<unsigned>(p.Wf));
        } else if (yy < 0) {
            yy = 0; xx = 0;                 // flat < 0 clamps to the first pixel
        } else if (yy > last_y) {
            yy = last_y; xx = last_x;       // flat > H*W-1 clamps to the last pixel
        }
        tile[T::off(i, jj >> 4) + (jj & 15)] = f[static_cast<long long>(yy) * p.pitch + xx];
    }
}

// Window that only PARTLY leaves the frame (the common border case: one or two columns / rows): the tile is
// fetched by TMA like an interior one -- out-of-frame elements arrive as zeros -- and only those elements are
// then patched with the reference's flat-index values.  Same element rule as gather_border_tile; the tile
// layout is the TMA one (row data starts at byte d = ox & 15).
constexpr int kPatchFlag = 256;
template <int W, int LOADER>
__device__ __noinline__ void patch_border_tile(const PassParams& p, const TileDesc dsc, unsigned char* tile,
                                               int frame, int lane) {
    using T = Tile<W, LOADER>;
    const unsigned char* f = (frame ? p.fb : p.fa) + dsc.pair * p.pair_stride;
    const int last_y = p.H - 1, last_x = p.Wf - 1, d = dsc.d & 15;
    const int ntop = clampi(-dsc.oy, 0, T::BY), nbot = clampi(dsc.oy + T::BY - p.H, 0, T::BY - ntop);
    const int nleft = clampi(-dsc.ox, 0, T::USED), nright = clampi(dsc.ox + T::USED - p.Wf, 0, T::USED - nleft);
    const int rows_oob = ntop + nbot, cols_oob = nleft + nright;
    const int n_rowpart = rows_oob * T::USED, n_all = n_rowpart + (T::BY - rows_oob) * cols_oob;
#pragma unroll 1
    for (int e = lane; e < n_all; e += 32) {
        int i, jj;
        if (e < n_rowpart) {
            const int ri = e / T::USED;
            jj = e - ri * T::USED;
            i = ri < ntop ? ri : T::BY - nbot + (ri - ntop);
        } else {
            const int e2 = e - n_rowpart, mi = e2 / cols_oob, cj = e2 - mi * cols_oob;
            i = ntop + mi;
            jj = cj < nleft ? cj : T::USED - nright + (cj - nleft);
        }
        int yy = dsc.oy + i, xx = dsc.ox + jj;
        if (xx < 0 || xx > last_x) {
            long long flat = static_cast<long long>(yy) * p.Wf + xx;
            const long long last = static_cast<long long>(p.H) * p.Wf - 1;
            flat = flat < 0 ? 0 : (flat > last ? last : flat);
            const unsigned uf = static_cast<unsigned>(flat);
            yy = static_cast<int>(uf / static_cast<unsigned>(p.Wf));
            xx = static_cast<int>(uf - static_cast<unsigned>(yy) * static_cast<unsigned>(p.Wf));
        } else if (yy < 0) {
            yy = 0; xx = 0;
        } else if (yy > last_y) {
            yy = last_y; xx = last_x;
        }
        tile[T::off(i, (d + jj) >> 4) + ((d + jj) & 15)] = f[static_cast<long long>(yy) * p.pitch + xx];
    }
}

// ----------------------------------------------------------------------------------------
// the kernel
// ----------------------------------------------------------------------------------------
template <int W, int LOADER, int SINK>
__global__ void __launch_bounds__(Smem<W, LOADER>::NWARPS * 32, 1) piv_fused_kernel(const __grid_constant__ CUtensorMap tmA,
                                                       const __grid_constant__ CUtensorMap tmB,
                                                       const __grid_constant__ PassParams p) {
    using G = Geo<W>;
    using T = Tile<W, LOADER>;
    using S = Smem<W, LOADER>;
    using F = Fft<W>;
    using PO = ParkOrder<W>;
    constexpr int NW = G::NW, HALF = G::HALF, LOGW = G::LOGW, PX = G::PX, PC = G::PC;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr bool kTmem = (SINK != SK_WIN);
    // rows of the window a lane transforms together: (2l, 2l+1) for CWS (they share the middle tile row's
    // horizontal taps), (l, l+H) otherwise (conflict-free tile reads)
    constexpr bool kAdjRows = (LOADER == LD_FRAME_CWS);

    extern __shared__ __align__(1024) unsigned char smem_cta[];
    __shared__ uint32_t tmem_base_sh;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;                      // <= S::NWARPS (launcher)
    unsigned char* smem = smem_cta + S::SHARED + warp * S::STRIDE;       // this warp's private slot
    constexpr bool kTC = S::kTC;
    __shared__ __align__(8) uint64_t tc_bar_sh[S::TC_GROUPS];                // MMA completion, one per group of four warps
    const int n_total = static_cast<int>(p.n_total);
    const int njobs = (n_total + NW - 1) / NW;
    const int job_stride = gridDim.x * nwarps;
    const uint32_t bar = smem_u32(smem + S::BAR_OFF);

    uint32_t tpark = 0;                                      // TMEM address of this warp's parking area
    if constexpr (kTmem) {
        if (warp == 0) tmem_alloc_512(smem_u32(&tmem_base_sh));
        tmem_fence_before();
        __syncthreads();
        tmem_fence_after();
        // parking area of this warp: its 32 lanes, TCOLS columns per group of four warps (with the tensor-core
        // row transform a group owns 2 TCOLS columns: accumulators first, parking after them)
        tpark = tmem_base_sh + (static_cast<uint32_t>((warp & 3) * 32) << 16) +
                static_cast<uint32_t>(kTC ? (warp >> 2) * 2 * G::TCOLS + G::TCOLS : (warp >> 2) * G::TCOLS);
    }
    if constexpr (kTC) {
        // DFT matrix B[n][x], n = output column: n = 2k -> 2 cos(2 pi k x / W) (k = 0..W/2-1), n = 2k + 1 ->
        // -2 sin(2 pi k x / W) (k >= 1), n = 1 -> 2 cos(pi x) (bin W/2): column pairs (2k, 2k+1) are the X entries
        // (2 Re R[k], 2 Im R[k]) with bins 0 and W/2 sharing slot 0, exactly what the FP32 row step stores.
        // fp16 hi + lo split (two MMAs into one accumulator) gives FP32-level accuracy.
        for (int e = threadIdx.x; e < W * W; e += blockDim.x) {
            const int n = e / W, xx = e - n * W, k = n >> 1;
            double sn, cs;
            sincospi(2.0 * ((k * xx) % W) / W, &sn, &cs);
            double v = (n & 1) ? -2.0 * sn : 2.0 * cs;
            if (n == 1) v = (xx & 1) ? -2.0 : 2.0;
            const __half hi = __double2half(v);
            const __half lo = __double2half(v - static_cast<double>(__half2float(hi)));
            const int off = sw128_offset(n, xx >> 3) + (xx & 7) * 2;
            *reinterpret_cast<__half*>(smem_cta + S::TC_TW_OFF + off) = hi;
            *reinterpret_cast<__half*>(smem_cta + S::TC_TW_OFF + 8192 + off) = lo;
        }
        if (threadIdx.x < S::TC_GROUPS) mbar_init(smem_u32(&tc_bar_sh[threadIdx.x]), 1);
        fence_proxy_async();
        __syncthreads();
    }
    if constexpr (T::kFrame) {
        if (lane == 0) mbar_init(bar, 1);
        __syncwarp();
    }

    const int wi = lane / HALF;            // window of this lane inside the job
    const int l = lane % HALF;             // rows (l, l + H) in phases R and I, column l in phase C
    unsigned char* const region = smem + S::REG_OFF + wi * G::REGION;
    float2* const Xw = reinterpret_cast<float2*>(region + G::doff(wi));      // X / Q / map of this lane's window

    // ---- descriptors of the (window, frame) tiles of one job ------------------------------
    auto make_desc = [&](int job) {
        if constexpr (T::kFrame) {
            TileDesc* desc = reinterpret_cast<TileDesc*>(smem + S::TD_OFF);
            if (lane < NW * 2) {
                const int q = lane, frame = q & 1;
                const int g = min(job * NW + (q >> 1), n_total - 1);
                const WinGeo w = window_geo(p, g);
                TileDesc dsc;
                frame_origin<LOADER>(p, g, w, frame, dsc.oy, dsc.ox, dsc.vy, dsc.vx);
                const bool interior = p.use_tma && dsc.ox >= 0 && dsc.oy >= 0 &&
                                      dsc.ox + T::USED <= p.Wf && dsc.oy + T::BY <= p.H;
                dsc.pair = w.pair;
                dsc.r0 = w.r0;
                dsc.c0 = w.c0;
                // partly outside (but the TMA box still overlaps the frame): TMA + patch of the outside elements;
                // entirely outside, or no TMA: every element is gathered
                const bool partial = p.use_tma && !interior && dsc.ox + T::USED > 0 && dsc.ox < p.Wf &&
                                     dsc.oy + T::BY > 0 && dsc.oy < p.H;
                dsc.d = interior ? (dsc.ox & 15) : (partial ? ((dsc.ox & 15) | kPatchFlag) : -1);
                desc[q] = dsc;
            }
            __syncwarp();
        }
    };

    // ---- bring the tiles of `frame` into the (currently dead) window buffers ----------------
    auto tile_ptr_plain = [&](int w2) -> unsigned char* { return smem + S::REG_OFF + w2 * G::REGION; };
    auto stage_issue = [&](int frame) {
        if constexpr (T::kFrame) {
            TileDesc* desc = reinterpret_cast<TileDesc*>(smem + S::TD_OFF);
            fence_proxy_async();            // generic accesses to the buffers are ordered before the TMA writes
            __syncwarp();
            uint32_t tx = 0;
#pragma unroll 1
            for (int w2 = 0; w2 < NW; ++w2) {
                const int q = w2 * 2 + frame;
                const TileDesc dsc = desc[q];
                if (dsc.d >= 0) {
                    tx += T::TX;
                } else {
                    gather_border_tile<W, LOADER>(p, dsc, smem + S::REG_OFF + w2 * G::REGION, frame, lane);
                }
            }
            // The barrier is armed for EVERY (job, frame), also with tx == 0 (all tiles gathered: the phase
            // completes at once), so its phase parity at the matching wait is simply `frame` -- no
            // per-warp pending / parity state to keep (it used to be spilled to local memory).
            if (lane == 0) {
                mbar_arrive_expect_tx(bar, tx);
#pragma unroll 1
                for (int w2 = 0; w2 < NW; ++w2) {
                    const TileDesc dsc = desc[w2 * 2 + frame];
                    if (dsc.d >= 0) {
                        // two call sites with the __grid_constant__ maps addressed statically: a selected
                        // pointer makes the compiler copy both maps to the (local-memory) stack
                        const uint32_t dst = smem_u32(smem + S::REG_OFF + w2 * G::REGION);
                        if (frame) tma_load_3d(dst, &tmB, bar, dsc.ox & ~15, dsc.oy, dsc.pair);
                        else tma_load_3d(dst, &tmA, bar, dsc.ox & ~15, dsc.oy, dsc.pair);
                    }
                }
            }
            __syncwarp();
            // gathered tiles are stored from byte 0
            if (lane < NW && desc[lane * 2 + frame].d < 0) desc[lane * 2 + frame].d = 0;
            __syncwarp();
        }
    };
    auto stage_wait = [&](int frame) {
        if constexpr (T::kFrame) mbar_wait(bar, static_cast<uint32_t>(frame));
    };

    // Every warp of the CTA runs the same number of iterations (optional block barriers inside keep the
    // warps in the same phase so that they share instruction fetches); a warp without work re-does
    // the last job with its output suppressed.
    if (p.sync_group && p.skew_ns > 0) __nanosleep(static_cast<unsigned>(p.skew_ns) * static_cast<unsigned>(warp >> 2));
    bool prefetched = false;                   // frame a's tiles of this iteration's job are already in flight
#pragma unroll 1
    for (int base = blockIdx.x * nwarps; base < njobs; base += job_stride) {
        const int job = min(base + warp, njobs - 1);
        const int g = min(job * NW + wi, n_total - 1);       // window of this lane (clamped: the last job may be ragged)
        const bool g_valid = (base + warp < njobs) && (job * NW + wi < n_total);
        if (!prefetched) {                     // otherwise the previous iteration's epilogue already did both
            make_desc(job);
            stage_issue(0);
        }

        float sum_a = 0.f, sum_b = 0.f;        // pixel sums of the window (valid in lanes l == 0)
        float2 x[W];                           // the FFT operand
        // ===================================================================================
        // Six FFT steps per lane around ONE shared transform body:
        //   s = 0  Ra    s = 1  Ca (+ park)    s = 2  Rb    s = 3  Cb forward (+ product)    s = 4  Cb inverse (-> Q)    s = 5  I
        // ===================================================================================
#pragma unroll 1
        for (int s = 0; s < 6; ++s) {
            if ((p.sync_mask >> s) & 1) {
                // lock step (shared instruction fetch): the whole CTA, or groups of four warps (one per
                // scheduler) so that different groups sit in different phases (FP vs shared-memory bound)
                if (p.sync_group == 1) asm volatile("bar.sync %0, 128;" ::"r"(1 + (warp >> 2)) : "memory");
                else if (p.sync_group == 2) asm volatile("bar.sync %0, %1;" ::"r"(1 + (warp & 3)), "r"((nwarps >> 2) * 32) : "memory");
                else if (p.sync_group >= 4) asm volatile("bar.sync %0, %1;" ::"r"(1 + warp / p.sync_group), "r"(p.sync_group * 32) : "memory");
                else __syncthreads();
            }
            // ---------------------------------------------------------------- load
            if (s == 0 || s == 2) {
                const int frame = s >> 1;
                stage_wait(frame);
                if constexpr (LOADER == LD_FRAME_INT || LOADER == LD_FRAME_CWS) {
                    const TileDesc* desc = reinterpret_cast<const TileDesc*>(smem + S::TD_OFF);
#pragma unroll 1
                    for (int w2 = 0; w2 < NW; ++w2) {
                        const TileDesc dsc = desc[w2 * 2 + frame];
                        if (dsc.d >= kPatchFlag) patch_border_tile<W, LOADER>(p, dsc, tile_ptr_plain(w2), frame, lane);
                    }
                    __syncwarp();
                }
                if constexpr (kTC) {
                    // rows l and l + W/2 of this warp's window -> fp16 rows 32 q + l of the group's two operand
                    // tiles (q = this warp's TMEM lane quarter), so that after the MMAs lane l of THIS warp finds
                    // the spectrum of row l in accumulator columns [0, W) and of row l + W/2 in [W, 2 W)
                    const int grp = warp >> 2, q = warp & 3;
                    unsigned char* const tA = smem_cta + S::TC_A_OFF + grp * 32768;
                    uint32_t w0[W / 4], w1[W / 4];
                    load_row_words<W, LOADER, W / 4>(region, l, 0, w0);
                    load_row_words<W, LOADER, W / 4>(region, l + HALF, 0, w1);
                    const int arow = 32 * q + l;
                    static_for<0, W / 8>([&](auto cc) {
                        constexpr int c = decltype(cc)::value;          // 8 pixels = 2 words -> one 16-byte chunk of fp16
                        uint4 h0, h1;
                        u8x4_to_h2x2(w0[2 * c], h0.x, h0.y);
                        u8x4_to_h2x2(w0[2 * c + 1], h0.z, h0.w);
                        u8x4_to_h2x2(w1[2 * c], h1.x, h1.y);
                        u8x4_to_h2x2(w1[2 * c + 1], h1.z, h1.w);
                        *reinterpret_cast<uint4*>(tA + sw128_offset(arow, c)) = h0;
                        *reinterpret_cast<uint4*>(tA + 16384 + sw128_offset(arow, c)) = h1;
                    });
                    fence_proxy_async();                    // generic-proxy stores -> visible to the tensor core
                    tmem_fence_before();
                    asm volatile("bar.sync %0, 128;" ::"r"(8 + grp) : "memory");
                    const uint32_t gbar = smem_u32(&tc_bar_sh[grp]);
                    const uint32_t tacc = tmem_base_sh + static_cast<uint32_t>(grp * 2 * G::TCOLS);
                    if (q == 0 && lane == 0) {
                        tmem_fence_after();
                        const uint64_t dA1 = umma_desc_sw128(smem_u32(tA)), dA2 = umma_desc_sw128(smem_u32(tA + 16384));
                        const uint64_t dBh = umma_desc_sw128(smem_u32(smem_cta + S::TC_TW_OFF));
                        const uint64_t dBl = umma_desc_sw128(smem_u32(smem_cta + S::TC_TW_OFF + 8192));
                        static_for<0, 2>([&](auto hc) {
                            constexpr int half_id = decltype(hc)::value;     // rows l (tile 1) / rows l + W/2 (tile 2)
                            const uint64_t dA = half_id ? dA2 : dA1;
                            static_for<0, W / 16>([&](auto kc) {             // UMMA K = 16 fp16 = 32 bytes = +2 in the address field
                                constexpr int k = decltype(kc)::value;
                                umma_f16(tacc + half_id * W, dA + 2 * k, dBh + 2 * k, k > 0);
                            });
                            static_for<0, W / 16>([&](auto kc) {
                                constexpr int k = decltype(kc)::value;
                                umma_f16(tacc + half_id * W, dA + 2 * k, dBl + 2 * k, 1);
                            });
                        });
                        umma_commit(gbar);
                    }
                    mbar_wait(gbar, static_cast<uint32_t>(frame));       // two commits per job: parity = frame
                    tmem_fence_after();
                    const uint32_t tsrc = tacc + (static_cast<uint32_t>(q * 32) << 16);
                    static_for<0, W / 16>([&](auto cc) {
                        constexpr int c = decltype(cc)::value;           // 32 columns = 16 (re, im) pairs
                        float2 v[16];
                        tmem_ld16(tsrc + 32 * c, v);
                        static_for<0, 16>([&](auto ic) { constexpr int i = decltype(ic)::value; x[16 * c + i] = v[i]; });
                    });
                    tmem_wait_ld();
                } else if constexpr (LOADER == LD_FRAME_INT || LOADER == LD_FRAME_ALN) {
                    const TileDesc* desc = reinterpret_cast<const TileDesc*>(smem + S::TD_OFF);
                    const int d = desc[wi * 2 + frame].d & 15;
                    uint32_t w0[W / 4], w1[W / 4];
                    // (a uniform switch over the word part of d with statically addressed registers instead
                    // of the select network was measured: no gain -- the loader is latency, not issue bound)
                    load_row_words<W, LOADER, W / 4>(region, l, d, w0);
                    load_row_words<W, LOADER, W / 4>(region, l + HALF, d, w1);
                    static_for<0, W>([&](auto jc) {
                        constexpr int j = decltype(jc)::value;
                        x[j] = make_float2(u8f(w0[j >> 2], j & 3), u8f(w1[j >> 2], j & 3));
                    });
                } else if constexpr (LOADER == LD_FRAME_CWS) {
                    const TileDesc* desc = reinterpret_cast<const TileDesc*>(smem + S::TD_OFF);
                    float4* xw = reinterpret_cast<float4*>(smem + S::XW_OFF);
                    int* xf = reinterpret_cast<int*>(smem + S::XF_OFF);
                    // per-column tap descriptors of this frame (shared by all rows of a window)
                    bool flag = false;
#pragma unroll
                    for (int e = lane; e < NW * W; e += 32) {
                        const int j = e & (W - 1), w2 = e >> LOGW;
                        const TileDesc dsc = desc[w2 * 2 + frame];
                        const AxisTap cx = cws_axis(dsc.c0 + j, dsc.vx);
                        xw[e] = make_float4(cx.w1, cx.w1, cx.w0, cx.w0);      // pairs: operands of the packed taps
                        xf[e] = (cx.exact ? 2 : 0) | ((cx.lo - (dsc.ox + j)) & 1);
                        flag |= cx.exact;
                    }
                    __syncwarp();
                    const TileDesc dsc = desc[wi * 2 + frame];
                    const float4* xwq = xw + wi * W;
                    const int* xfq = xf + wi * W;
                    // Window rows (2l, 2l+1) need tile rows 2l .. 2l+2.  The reference's four-term sum
                    // (PB:187-192) is evaluated in a separable form, vertical tap first: per tile column one
                    // packed pair v = (A wyA1 + B wyA0, B wyB1 + C wyB0) for the lane's two window rows (A, B, C =
                    // its three tile rows), then the horizontal tap v[j] wx1 + v[j+1] wx0, again packed -- the
                    // result IS the FFT operand (row 2l, row 2l+1).  Same value up to FP32 rounding (~1e-7
                    // relative; the function-level entry point pivb200_bilinear_cws keeps the reference's exact
                    // evaluation order).  5 packed FP32 instructions per output pair instead of 10 scalar ones.
                    const int ra = 2 * l;
                    const AxisTap cyA = cws_axis(dsc.r0 + ra, dsc.vy), cyB = cws_axis(dsc.r0 + ra + 1, dsc.vy);
                    // exact-integer coordinates anywhere in the job (columns: the table loop above; rows: every row
                    // of a window belongs to one of its lanes)
                    const bool anyflag = __any_sync(FULL, flag || cyA.exact || cyB.exact);
                    uint32_t wA[W / 4 + 1], wB[W / 4 + 1], wC[W / 4 + 1];
                    load_row_words<W, LOADER, W / 4 + 1>(region, ra, dsc.d & 15, wA);
                    load_row_words<W, LOADER, W / 4 + 1>(region, ra + 1, dsc.d & 15, wB);
                    load_row_words<W, LOADER, W / 4 + 1>(region, ra + 2, dsc.d & 15, wC);
                    float cA = u8f(wA[0], 0), cB = u8f(wB[0], 0), cC = u8f(wC[0], 0);
                    if (!anyflag) {
                        const float2 wy1 = make_float2(cyA.w1, cyB.w1), wy0 = make_float2(cyA.w0, cyB.w0);
                        float2 vc = cfma(make_float2(cA, cB), wy1, cmul2(make_float2(cB, cC), wy0));
                        static_for<0, W>([&](auto jc) {
                            constexpr int j = decltype(jc)::value;
                            const float nA = u8f(wA[(j + 1) >> 2], (j + 1) & 3);
                            const float nB = u8f(wB[(j + 1) >> 2], (j + 1) & 3);
                            const float nC = u8f(wC[(j + 1) >> 2], (j + 1) & 3);
                            const float2 vn = cfma(make_float2(nA, nB), wy1, cmul2(make_float2(nB, nC), wy0));
                            const float4 wx = xwq[j];
                            x[j] = cfma(vc, make_float2(wx.x, wx.y), cmul2(vn, make_float2(wx.z, wx.w)));
                            vc = vn;
                        });
                    } else {
                        // some coordinate of the job is an exact integer: there the reference's weights all
                        // vanish and the value is patched to the tap at (floor y, floor x) (PB:170, 193)
                        const bool jyA = (cyA.lo - (dsc.oy + ra)) & 1, jyB = (cyB.lo - (dsc.oy + ra + 1)) & 1;
                        static_for<0, W>([&](auto jc) {
                            constexpr int j = decltype(jc)::value;
                            const float nA = u8f(wA[(j + 1) >> 2], (j + 1) & 3);
                            const float nB = u8f(wB[(j + 1) >> 2], (j + 1) & 3);
                            const float nC = u8f(wC[(j + 1) >> 2], (j + 1) & 3);
                            const float4 wx4 = xwq[j];
                            const float2 wx = make_float2(wx4.x, wx4.z);
                            const int fl = xfq[j];
                            const float hA = fmaf(cA, wx.x, nA * wx.y);
                            const float hB = fmaf(cB, wx.x, nB * wx.y);
                            const float hC = fmaf(cC, wx.x, nC * wx.y);
                            const float qA = (fl & 1) ? nA : cA, qB = (fl & 1) ? nB : cB, qC = (fl & 1) ? nC : cC;
                            const float vA = ((fl & 2) || cyA.exact) ? (jyA ? qB : qA) : fmaf(hA, cyA.w1, hB * cyA.w0);
                            const float vB = ((fl & 2) || cyB.exact) ? (jyB ? qC : qB) : fmaf(hB, cyB.w1, hC * cyB.w0);
                            x[j] = make_float2(vA, vB);
                            cA = nA; cB = nB; cC = nC;
                        });
                    }
                } else if constexpr (LOADER == LD_EXPL_F32) {
                    const float* base = static_cast<const float*>(frame ? p.wb : p.wa) + static_cast<long long>(g) * W * W;
                    const float4* r0 = reinterpret_cast<const float4*>(base + l * W);
                    const float4* r1 = reinterpret_cast<const float4*>(base + (l + HALF) * W);
                    static_for<0, W / 4>([&](auto cc) {
                        constexpr int c = decltype(cc)::value;
                        const float4 va = __ldg(r0 + c), vb = __ldg(r1 + c);
                        x[4 * c] = make_float2(va.x, vb.x);
                        x[4 * c + 1] = make_float2(va.y, vb.y);
                        x[4 * c + 2] = make_float2(va.z, vb.z);
                        x[4 * c + 3] = make_float2(va.w, vb.w);
                    });
                } else {
                    const unsigned char* base = static_cast<const unsigned char*>(frame ? p.wb : p.wa) + static_cast<long long>(g) * W * W;
                    const uint4* r0 = reinterpret_cast<const uint4*>(base + l * W);
                    const uint4* r1 = reinterpret_cast<const uint4*>(base + (l + HALF) * W);
                    static_for<0, W / 16>([&](auto cc) {
                        constexpr int c = decltype(cc)::value;
                        const uint4 qa = __ldg(r0 + c), qb = __ldg(r1 + c);
                        const uint32_t wa[4] = {qa.x, qa.y, qa.z, qa.w};
                        const uint32_t wb[4] = {qb.x, qb.y, qb.z, qb.w};
                        static_for<0, 16>([&](auto bc) {
                            constexpr int b = decltype(bc)::value;
                            x[16 * c + b] = make_float2(u8f(wa[b >> 2], b & 3), u8f(wb[b >> 2], b & 3));
                        });
                    });
                }
                __syncwarp();                   // tiles fully read before X overwrites the buffer

                if constexpr (SINK == SK_WIN) {
                    if (g_valid) {
                        float* dst = (frame ? p.win_b_out : p.win_a_out) + static_cast<long long>(g) * W * W;
                        float4* o0 = reinterpret_cast<float4*>(dst + (kAdjRows ? 2 * l : l) * W);
                        float4* o1 = reinterpret_cast<float4*>(dst + (kAdjRows ? 2 * l + 1 : l + HALF) * W);
                        static_for<0, W / 4>([&](auto cc) {
                            constexpr int c = decltype(cc)::value;
                            o0[c] = make_float4(x[4 * c].x, x[4 * c + 1].x, x[4 * c + 2].x, x[4 * c + 3].x);
                            o1[c] = make_float4(x[4 * c].y, x[4 * c + 1].y, x[4 * c + 2].y, x[4 * c + 3].y);
                        });
                    }
                    if (s == 0) stage_issue(1);
                    ++s;                        // no transforms: skip the column step
                    continue;
                }
            } else if (s == 1 || s == 3) {
                // lane l' stored the spectra of its two window rows in buffer rows l' and l' + H
                static_for<0, W>([&](auto tc) {
                    constexpr int t = decltype(tc)::value;
                    constexpr int brow = kAdjRows ? (t >> 1) + (t & 1) * HALF : t;
                    x[t] = Xw[brow * PX + l];
                });
                __syncwarp();                   // X fully read: the buffer is dead
                if (s == 1) stage_issue(1);     // frame b tiles arrive while column a is transformed
            } else if (s == 5) {
                // two Hermitian rows l, l + W/2 of Q packed into one complex inverse FFT
                const float2 a0 = Xw[l * PX], b0 = Xw[(l + HALF) * PX];
                x[0] = make_float2(b0.x, a0.x);
                x[HALF] = make_float2(b0.y, a0.y);
                static_for<1, HALF>([&](auto kc_) {
                    constexpr int k = decltype(kc_)::value;
                    const float2 R1 = Xw[l * PX + k], R2 = Xw[(l + HALF) * PX + k];
                    x[k] = make_float2(R1.y + R2.x, R1.x - R2.y);
                    x[W - k] = make_float2(R2.x - R1.y, R1.x + R2.y);
                });
                __syncwarp();                   // Q fully read before the map overwrites it
            }

            if constexpr (kTC) {
                if (s != 0 && s != 2) F::run(x);            // the row transforms came from the tensor cores
            } else if constexpr (SINK != SK_WIN) {
                F::run(x);
            }

            // ---------------------------------------------------------------- store
            if (kTC && (s == 0 || s == 2)) {
                // x[0 .. W/2) = spectrum of row l, x[W/2 .. W) = of row l + W/2, already in the X layout
                float2* X1 = Xw + l * PX;
                float2* X2 = Xw + (l + HALF) * PX;
                static_for<0, HALF>([&](auto kc_) {
                    constexpr int k = decltype(kc_)::value;
                    X1[k] = x[k];
                    X2[k] = x[HALF + k];
                });
                __syncwarp();
            } else if (s == 0 || s == 2) {
                // z = r1 + i r2: R1[k] = Z[k] + conj Z[W-k], R2[k] = -i (Z[k] - conj Z[W-k])  (x2, folded into K);
                // bins 0 and W/2 are real and share one complex slot
                float2* X1 = Xw + l * PX;
                float2* X2 = Xw + (l + HALF) * PX;
                const float2 z0 = x[F::pos(0)], zh = x[F::pos(HALF)];
                X1[0] = make_float2(2.0f * z0.x, 2.0f * zh.x);
                X2[0] = make_float2(2.0f * z0.y, 2.0f * zh.y);
                static_for<1, HALF>([&](auto kc_) {
                    constexpr int k = decltype(kc_)::value;
                    const float2 zk = x[F::pos(k)], zn = x[F::pos(W - k)];
                    X1[k] = make_float2(zk.x + zn.x, zk.y - zn.y);
                    X2[k] = make_float2(zk.y + zn.y, zn.x - zk.x);
                });
                __syncwarp();
            } else if (s == 1) {
                if constexpr (kTmem) {
                    static_for<0, W / 8>([&](auto cc) {
                        constexpr int c = decltype(cc)::value;
                        float2 v[8];
                        static_for<0, 8>([&](auto ic) {
                            constexpr int i = decltype(ic)::value;
                            v[i] = x[F::pos(park_bin<W>(8 * c + i))];
                        });
                        tmem_st8(tpark + 16 * c, v);
                    });
                    tmem_wait_st();
                }
            } else if (s == 3) {
                if constexpr (kTmem) {
                    // The lane that owns the packed real columns (0, W/2), l == 0, needs bins q and W - q of
                    // both spectra at once to separate them.  It drops its two spectra into the (dead)
                    // window buffer and the H lanes of the window do the H + 1 small pair jobs in parallel.
                    float2* const sB = Xw;              // [W] b spectrum of column 0, natural order
                    float2* const sA = Xw + W;          // [W] parked a spectrum
                    float2* const sV = Xw + 2 * W;      // [W] inverse-transform input of column 0, stored (im, re)
                    if (l == 0) {
                        static_for<0, W>([&](auto qc) { constexpr int q = decltype(qc)::value; sB[q] = x[F::pos(q)]; });
                    }
                    static_for<0, PO::NCH>([&](auto cc) {
                        constexpr int c = decltype(cc)::value;
                        constexpr int CH = PO::CH;
                        float2 A[CH];
                        if constexpr (CH == 8) tmem_ld8(tpark + 2 * CH * c, A);
                        else tmem_ld16(tpark + 2 * CH * c, A);
                        tmem_wait_ld();
                        if (l == 0) {
                            static_for<0, CH>([&](auto ic) { constexpr int i = decltype(ic)::value; sA[park_bin<W>(CH * c + i)] = A[i]; });
                        }
                        // P[q] = conj(A^[q]) B^[q]: read from slot pos(q), stored (im, re) into slot q for the
                        // inverse transform (the chunk is closed under q -> pos(q), so this is in place)
                        float2 P[CH];
                        static_for<0, CH>([&](auto ic) {
                            constexpr int i = decltype(ic)::value;
                            const float2 B = x[F::pos(park_bin<W>(CH * c + i))];
                            P[i] = make_float2(fmaf(A[i].x, B.y, -A[i].y * B.x), fmaf(A[i].x, B.x, A[i].y * B.y));
                        });
                        static_for<0, CH>([&](auto ic) { constexpr int i = decltype(ic)::value; x[park_bin<W>(CH * c + i)] = P[i]; });
                    });
                    __syncwarp();
                    if (l == 0) {
                        // bins 0 and W/2 are their own partners: everything is real
                        const float2 A0 = sA[0], B0 = sB[0], Ah = sA[HALF], Bh = sB[HALF];
                        sum_a = 0.5f * A0.x;
                        sum_b = 0.5f * B0.x;
                        // drop the DC bin (mean product): a constant that `- amin` removes anyway
                        const float p00 = (SINK == SK_DISP) ? 0.f : A0.x * B0.x;
                        sV[0] = make_float2(A0.y * B0.y, p00);
                        sV[HALF] = make_float2(Ah.y * Bh.y, Ah.x * Bh.x);
                    } else {
                        // C = c0^ + i cH^ (two real columns): separate, multiply, re-pack
                        const float2 Aq = sA[l], An = sA[W - l], Bq = sB[l], Bn = sB[W - l];
                        const float2 a0 = make_float2(Aq.x + An.x, Aq.y - An.y);      // 2 c0^[q]
                        const float2 ah = make_float2(Aq.y + An.y, An.x - Aq.x);      // 2 cH^[q]
                        const float2 b0 = make_float2(0.25f * (Bq.x + Bn.x), 0.25f * (Bq.y - Bn.y));
                        const float2 bh = make_float2(0.25f * (Bq.y + Bn.y), 0.25f * (Bn.x - Bq.x));
                        const float2 P0 = make_float2(fmaf(a0.x, b0.x, a0.y * b0.y), fmaf(a0.x, b0.y, -a0.y * b0.x));
                        const float2 Ph = make_float2(fmaf(ah.x, bh.x, ah.y * bh.y), fmaf(ah.x, bh.y, -ah.y * bh.x));
                        // V[q] = P0 + i Ph, V[W-q] = conj(P0) + i conj(Ph); stored (im, re)
                        sV[l] = make_float2(P0.y + Ph.x, P0.x - Ph.y);
                        sV[W - l] = make_float2(Ph.x - P0.y, P0.x + Ph.y);
                    }
                    __syncwarp();
                    if (l == 0) {
                        static_for<0, W>([&](auto qc) { constexpr int q = decltype(qc)::value; x[q] = sV[q]; });
                    }
                    __syncwarp();                       // scratch fully read before Q overwrites the buffer
                }
            } else if (s == 4) {
                static_for<0, W>([&](auto tc) {
                    constexpr int t = decltype(tc)::value;
                    const float2 o = x[F::pos(t)];
                    Xw[t * PX + l] = make_float2(o.y, o.x);
                });
                __syncwarp();
            }
        }
        if constexpr (SINK == SK_WIN) continue;

        // raw row l -> shifted row l + W/2 (values x[].y); raw row l + W/2 -> shifted row l (x[].x)
        float* mapw = reinterpret_cast<float*>(Xw);
        float mx_hi = -FLT_MAX, mx_lo = -FLT_MAX, mn = FLT_MAX;     // hi: shifted row l + HALF
        static_for<0, W>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            constexpr int sc = (j + HALF) % W;
            const float2 o = x[F::pos(j)];
            mapw[(l + HALF) * PC + sc] = o.y;
            mapw[l * PC + sc] = o.x;
            mx_hi = fmaxf(mx_hi, o.y);
            mx_lo = fmaxf(mx_lo, o.x);
            mn = fminf(mn, fminf(o.x, o.y));
        });
        __syncwarp();

        if constexpr (SINK == SK_CORR) {
#pragma unroll 1
            for (int w2 = 0; w2 < NW; ++w2) {
                const int g2 = job * NW + w2;
                if (base + warp >= njobs || g2 >= n_total) break;
                const float* mw = reinterpret_cast<const float*>(smem + S::REG_OFF + w2 * G::REGION + G::doff(w2));
                float* out = p.corr_out + static_cast<long long>(g2) * W * W;
                for (int e = lane; e < W * W; e += 32)
                    out[e] = mw[(e >> LOGW) * PC + (e & (W - 1))] * (1.0f / G::K);
            }
            __syncwarp();
            continue;
        }

        // =============================== epilogue ==========================================
        if constexpr (SINK == SK_DISP) {
            // predictor glue operands of this window: requested now, consumed after the fit
            double base_u = 0.0, base_v = 0.0, pred_u = 0.0, pred_v = 0.0;
            if (l == 0 && g_valid) {
                if (p.base_u) { base_u = p.base_u[g]; base_v = p.base_v[g]; }
                if (p.pred_u) { pred_u = p.pred_u[g]; pred_v = p.pred_v[g]; }
            }
            const float gmax = group_max<HALF>(fmaxf(mx_hi, mx_lo)), gmin = group_min<HALF>(mn);
            // first maximum in flat (row-major) order of the shifted map (torch argmax, PB:383)
            constexpr int BIG = 1 << 20;
            int R = group_min_int<HALF>(min(mx_lo == gmax ? l : BIG, mx_hi == gmax ? l + HALF : BIG));
            R = min(R, W - 1);      // only reachable with NaN input
            int C = group_min_int<HALF>(
                (mapw[R * PC + l] == gmax) ? l : ((mapw[R * PC + l + HALF] == gmax) ? l + HALF : BIG));
            C = min(C, W - 1);
            constexpr int N2 = W * W;
            const int m = R * W + C;
            auto at = [&](int f) { return mapw[(f >> LOGW) * PC + (f & (W - 1))]; };
            // flat neighbours, guarded only at the array ends (PB:385-392)
            const int il = (m + 1 >= N2 - 1) ? m : m + 1;
            const int ir = (m - 1 <= 0) ? m : m - 1;
            const int it_ = (m + W >= N2 - 1) ? m : m + W;
            const int ib = (m - W <= 0) ? m : m - W;
            double eps = static_cast<double>(G::K) * 1e-7;
            // lanes l == 0 hold the pixel sums of their window
            const float sa = __shfl_sync(FULL, sum_a, wi * HALF), sb = __shfl_sync(FULL, sum_b, wi * HALF);
            if (p.first_pass) eps *= (static_cast<double>(sa) / N2) * (static_cast<double>(sb) / N2);
            const float f_l = at(il), f_r = at(ir), f_t = at(it_), f_b = at(ib);
            // second peak: maximum outside the 7x7 flat-index patch around m, each patch index
            // clamped to [0, N2-1] (PB:346-358).  Rows that cannot touch the patch reuse the
            // row maxima from registers; the <= 8 candidate rows are rescanned from smem.
            float sp = -FLT_MAX;
            if (p.validate) {
                const int lo_f = m - 3 - 3 * W, hi_f = m + 3 + 3 * W;
                const int ra = max(lo_f, 0) >> LOGW, rb = min(hi_f, N2 - 1) >> LOGW;
                if (l < ra || l > rb) sp = fmaxf(sp, mx_lo);
                if (l + HALF < ra || l + HALF > rb) sp = fmaxf(sp, mx_hi);
                for (int rr = ra; rr <= rb; ++rr) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int cc = l + h * HALF;
                        const int f = rr * W + cc;
                        const int e = f - lo_f;                     // (i+3) + W (j+3)
                        bool in_patch = (e >= 0) && ((e & (W - 1)) <= 6) && ((e >> LOGW) <= 6);
                        in_patch |= (f == 0 && lo_f <= 0) || (f == N2 - 1 && hi_f >= N2 - 1);
                        if (!in_patch) sp = fmaxf(sp, mapw[rr * PC + cc]);
                    }
                }
                sp = group_max<HALF>(sp);
            }
            // The map has been read for the last time: the next job's descriptors and frame-a tiles are requested
            // now, so the TMA (and the shift loads of make_desc) run underneath the FP64 fit below.
            // (Not for the CWS loader: measured 1-2 % slower there -- its descriptor set-up is heavier and the
            // kernel is register-bound.)
            if constexpr (LOADER != LD_FRAME_CWS) {
                __syncwarp();
                prefetched = false;
                if (base + job_stride < njobs) {
                    make_desc(min(base + job_stride + warp, njobs - 1));
                    stage_issue(0);
                    prefetched = true;
                }
            }
            const double dmin = static_cast<double>(gmin);
            const double cm = (static_cast<double>(gmax) - dmin) + eps;
            const double cl = (static_cast<double>(f_l) - dmin) + eps;
            const double cr = (static_cast<double>(f_r) - dmin) + eps;
            const double ct = (static_cast<double>(f_t) - dmin) + eps;
            const double cb = (static_cast<double>(f_b) - dmin) + eps;
            // the five logarithms are evaluated by five lanes of the window (one log() body, 1/5 of the FP64 work)
            const double lsel = (l == 0) ? cm : (l == 1) ? cl : (l == 2) ? cr : (l == 3) ? ct : cb;
            const double lg = log(lsel);
            const int l0 = wi * HALF;
            const double lm = __shfl_sync(FULL, lg, l0), ll = __shfl_sync(FULL, lg, l0 + 1), lr = __shfl_sync(FULL, lg, l0 + 2),
                         lt = __shfl_sync(FULL, lg, l0 + 3), lb = __shfl_sync(FULL, lg, l0 + 4);
            double du = static_cast<double>(C) + (lr - ll) / (2.0 * (ll + lr) - 4.0 * lm) - static_cast<double>(HALF);
            double dv = static_cast<double>(R) + (lb - lt) / (2.0 * (lb + lt) - 4.0 * lm) - static_cast<double>(HALF);
            // torch.nan_to_num (PB:418-419)
            du = isnan(du) ? 0.0 : (isinf(du) ? copysign(DBL_MAX, du) : du);
            dv = isnan(dv) ? 0.0 : (isinf(dv) ? copysign(DBL_MAX, dv) : dv);

            bool invalid = false;
            float ratio = 0.f;
            if (p.validate) {
                const double c2 = (static_cast<double>(sp) - dmin) + eps;
                const double rt = cm / c2;
                invalid = rt < p.val_ratio;
                ratio = static_cast<float>(rt);
            }
            if (p.first_pass && (sa == 0.f || sb == 0.f)) {
                // black window: the reference divides by a zero mean (PB:513-514), every value is NaN,
                // nan_to_num gives 0 and the NaN ratio compares False (valid)
                du = dv = 0.0;
                invalid = false;
                ratio = 0.f;
            }
            if (l == 0 && g_valid) {
                double uo = du + base_u;
                double vo = dv + base_v;
                if (p.pred_u) {
                    // PB:731-738: reject where the correction exceeds a positive predictor, or invalid
                    if ((du > pred_u && rint(pred_u) > 0.0) || invalid) uo = pred_u;
                    if ((dv > pred_v && rint(pred_v) > 0.0) || invalid) vo = pred_v;
                }
                p.u[g] = uo;
                p.v[g] = vo;
                if (p.mask) p.mask[g] = invalid ? 1 : 0;
                if (p.ratio) p.ratio[g] = ratio;
            }
            __syncwarp();
        }
    }

    if constexpr (kTmem) {
        tmem_fence_before();
        __syncthreads();
        if (warp == 0) tmem_dealloc_512(tmem_base_sh);
    }
}

}  // namespace pivb200
