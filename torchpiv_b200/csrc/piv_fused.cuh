// Fused PIV pass kernel for sm_100a: window extraction (TMA 2-D/3-D tile loads of the uint8
// frame) -> CWS bilinear / DWS integer window shift -> 2-D cross-correlation by in-register
// FFTs -> fft-shifted peak search, 3-point log-Gaussian sub-pixel fit, peak-ratio validation
// and predictor glue.  Only the displacement vectors leave the SM.
//
// Replaces, for one pass, the reference's eager op chain
//   moving_window_array (PB:220-247) -> [biliniar_interpolation_CWS | interpolation_DWS]
//   (PB:147-216) -> correalte_fft (PB:249-257) -> corr - amin (PB:518/724/796) ->
//   correlation_to_displacement + peak2peak_secondpeak (PB:346-422) -> replacement logic
//   (PB:728-738 / 800-810).
//
// Execution model: one WARP owns a "job" of NW = 64 / W interrogation windows (W = 64/32/16)
// and is persistent over jobs; a CTA is a single warp, so the only synchronisation is
// __syncwarp and one mbarrier for the TMA tiles.  Per job:
//   R  64 row FFTs (2 per lane) of z = a + i b                       -> smem M  [W][W+1] float2
//   C  per lane one column pair (k, W-k): 2 column FFTs, spectrum separation fused with the
//      conjugate product (4 conj(FA) FB), inverse column FFT        -> smem Q  [W][W/2+1] float2
//   I  per lane two Hermitian rows packed into one inverse FFT      -> 2 rows of the map in regs
//   E  min / argmax / neighbours / second peak on the shifted map   -> u, v, mask
// The next job's tiles are requested (TMA) right after phase R has consumed the current ones.
#pragma once
#include <cuda.h>
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
    static constexpr int HALF = W / 2;                // lanes per window in phases C, I, E
    static constexpr int LOGW = (W == 64) ? 6 : (W == 32 ? 5 : 4);
    static constexpr int PM = W + 1;                  // pitch of M   (float2)
    static constexpr int PQ = W / 2 + 1;              // pitch of Q   (float2)
    static constexpr int PC = W + 1;                  // pitch of map (float)
    static constexpr int MB = W * PM * 8;             // bytes of one window's exchange region
    static constexpr float K = 4.0f * W * W;          // map = K * sum_x a(x) b(x+s)
};

template <int W, int LOADER>
struct Tile {
    static constexpr bool kFrame = (LOADER == LD_FRAME_INT || LOADER == LD_FRAME_CWS);
    // TMA box.  The global start address of a box row must be 16-byte aligned (an unaligned x
    // coordinate faults as "illegal instruction" on sm_100a), so the box starts at the window's
    // x origin rounded DOWN to 16 and is 16 bytes wider than the bytes that are used; the row
    // loader re-aligns in registers by d = origin & 15.
    static constexpr int USED = (LOADER == LD_FRAME_CWS) ? W + 1 : W;    // bytes of a row that are read
    static constexpr int BX = W + 16;                                    // box bytes per row
    static constexpr int BY = (LOADER == LD_FRAME_CWS) ? W + 1 : W;      // box rows
    static constexpr int SWZ = (BX == 32) ? 1 : 0;                       // TMA swizzle: 32B / none
    static constexpr int ALIGN = (SWZ == 1) ? 256 : 128;
    static constexpr int TX = BX * BY;
    static constexpr int BYTES = kFrame ? ((TX + ALIGN - 1) / ALIGN) * ALIGN : 0;
    // byte offset of 16-byte chunk `chunk` of row `row`.  Row pitches of 80 / 48 bytes (and the
    // 32-byte swizzle for 32) make 8 consecutive rows hit 8 distinct 16-byte bank groups, so the
    // per-lane LDS.128 row reads are bank-conflict free.
    __device__ static __forceinline__ int off(int row, int chunk) {
        if constexpr (SWZ == 1) return row * 32 + ((chunk ^ ((row >> 2) & 1)) << 4);
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

template <int W, int LOADER>
struct Smem {
    using G = Geo<W>;
    using T = Tile<W, LOADER>;
    static constexpr int TILE_OFF = 0;
    static constexpr int TILES = G::NW * 2 * T::BYTES;
    static constexpr int EX_OFF = ((TILES + 127) / 128) * 128;
    static constexpr int EX = G::NW * G::MB;
    static constexpr int XW_OFF = EX_OFF + EX;                               // float2 [NW][2][W]
    static constexpr int XW = (LOADER == LD_FRAME_CWS) ? G::NW * 2 * W * 8 : 0;
    static constexpr int XF_OFF = XW_OFF + XW;                               // int    [NW][2][W]
    static constexpr int XF = (LOADER == LD_FRAME_CWS) ? G::NW * 2 * W * 4 : 0;
    static constexpr int TD_OFF = ((XF_OFF + XF + 15) / 16) * 16;            // TileDesc [2][NW][2]: per (window, frame), double buffered
    static constexpr int TD = T::kFrame ? 2 * G::NW * 2 * 32 : 0;
    static constexpr int BAR_OFF = ((TD_OFF + TD + 7) / 8) * 8;
    static constexpr int TOTAL = BAR_OFF + 8;
    // One CTA per SM made of NWARPS independent warps (each with its own STRIDE bytes of smem)
    // that walk the phases in lock step, so the SM fetches the (large, fully unrolled) instruction
    // stream once per CTA instead of once per warp.
    static constexpr int STRIDE = ((TOTAL + 255) / 256) * 256;
    static constexpr int SMEM_MAX = 232448;           // 227 KB opt-in limit per CTA
    static constexpr int NWARPS = (SMEM_MAX / STRIDE) > 16 ? 16 : (SMEM_MAX / STRIDE);
    static constexpr int CTA_BYTES = NWARPS * STRIDE;
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

__device__ __forceinline__ float u8f(uint32_t word, int b) {
    return static_cast<float>((word >> (8 * b)) & 0xffu);
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return max(lo, min(hi, v)); }

// 4 * conj(FA) * FB where FA = (X + conj(Yc)) / 2, FB = (X - conj(Yc)) / (2i): spectrum separation of
// z = a + i b fused with the conjugate product of the reference's correalte_fft (PB:255).
__device__ __forceinline__ float2 xcorr_bin(float2 X, float2 Yc) {
    float re = 2.0f * fmaf(X.x, Yc.y, X.y * Yc.x);
    float im = fmaf(Yc.x, Yc.x, Yc.y * Yc.y) - fmaf(X.x, X.x, X.y * X.y);
    return make_float2(re, im);
}

// ----------------------------------------------------------------------------------------
// window geometry
// ----------------------------------------------------------------------------------------
struct WinGeo {
    int pair, r0, c0;
};
__device__ __forceinline__ WinGeo window_geo(const PassParams& p, int g) {
    const int n = p.n_rows * p.n_cols;
    WinGeo w;
    w.pair = g / n;
    const int loc = g - w.pair * n;
    const int wr = loc / p.n_cols;
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
// the kernel
// ----------------------------------------------------------------------------------------
template <int W, int LOADER, int SINK>
__global__ void __launch_bounds__(Smem<W, LOADER>::NWARPS * 32, 1) piv_fused_kernel(const __grid_constant__ CUtensorMap tmA,
                                                       const __grid_constant__ CUtensorMap tmB,
                                                       const PassParams p) {
    using G = Geo<W>;
    using T = Tile<W, LOADER>;
    using S = Smem<W, LOADER>;
    using F = Fft<W>;
    constexpr int NW = G::NW, HALF = G::HALF, LOGW = G::LOGW, PM = G::PM, PQ = G::PQ, PC = G::PC;
    constexpr unsigned FULL = 0xffffffffu;

    extern __shared__ __align__(1024) unsigned char smem_cta[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    unsigned char* smem = smem_cta + warp * S::STRIDE;       // this warp's private region
    const int n_total = static_cast<int>(p.n_total);
    const int njobs = (n_total + NW - 1) / NW;
    const int job_stride = gridDim.x * S::NWARPS;
    const uint32_t bar = smem_u32(smem + S::BAR_OFF);
    uint32_t parity = 0;
    bool pending = false;

    if constexpr (T::kFrame) {
        if (lane == 0) mbar_init(bar, 1);
        __syncwarp();
    }

    // ---- request / gather the input tiles of one job -----------------------------------
    // buf: which half of the double-buffered descriptor table the job uses
    auto stage_tiles = [&](int job, int buf) {
        if constexpr (T::kFrame) {
            TileDesc* desc = reinterpret_cast<TileDesc*>(smem + S::TD_OFF) + buf * (NW * 2);
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
                dsc.d = interior ? (dsc.ox & 15) : -1;
                desc[q] = dsc;
            }
            __syncwarp();
            uint32_t tx = 0;
#pragma unroll 1
            for (int q = 0; q < NW * 2; ++q) {
                const TileDesc dsc = desc[q];
                if (dsc.d >= 0) {
                    tx += T::TX;
                } else {
                    // border window: the reference addresses taps by FLAT index clamped to
                    // [0, H*W-1] (PB:172-180, 213-214), i.e. columns wrap into neighbouring rows.
                    const unsigned char* f = ((q & 1) ? p.fb : p.fa) + dsc.pair * p.pair_stride;
                    unsigned char* tile = smem + S::TILE_OFF + q * T::BYTES;
                    const int last_y = p.H - 1, last_x = p.Wf - 1;
#pragma unroll 4
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
            }
            if (tx != 0) {
                if (lane == 0) {
                    fence_proxy_async();
                    mbar_arrive_expect_tx(bar, tx);
#pragma unroll 1
                    for (int q = 0; q < NW * 2; ++q) {
                        const TileDesc dsc = desc[q];
                        if (dsc.d >= 0)
                            tma_load_3d(smem_u32(smem + S::TILE_OFF + q * T::BYTES), (q & 1) ? &tmB : &tmA,
                                        bar, dsc.ox & ~15, dsc.oy, dsc.pair);
                    }
                }
                pending = true;
            }
            __syncwarp();
            // gathered tiles are stored from byte 0
            if (lane < NW * 2 && desc[lane].d < 0) desc[lane].d = 0;
            __syncwarp();
        }
    };

    // every warp of the CTA runs the same number of iterations (block barriers inside); a warp
    // without work re-does the last job and suppresses its output (n_total guard via `active`)
    int job = blockIdx.x * S::NWARPS + warp;
    int buf = 0;
    stage_tiles(min(job, njobs - 1), buf);

    for (int base = blockIdx.x * S::NWARPS; base < njobs; base += job_stride, job += job_stride, buf ^= 1) {
        const bool active = job < njobs;
        const int job_c = min(job, njobs - 1);
        if (p.sync_mask & 64) __syncthreads();
        if constexpr (T::kFrame) {
            if (pending) {
                mbar_wait(bar, parity);
                parity ^= 1u;
                pending = false;
            }
        }

        // ---- CWS: per-column tap descriptors (shared by all rows of a window) ------------
        bool anyflag = false;
        if constexpr (LOADER == LD_FRAME_CWS) {
            float2* xw = reinterpret_cast<float2*>(smem + S::XW_OFF);
            int* xf = reinterpret_cast<int*>(smem + S::XF_OFF);
            bool flag = false;
            const TileDesc* desc = reinterpret_cast<const TileDesc*>(smem + S::TD_OFF) + buf * (NW * 2);
#pragma unroll
            for (int e = lane; e < NW * 2 * W; e += 32) {
                const int j = e & (W - 1), q = e >> LOGW;      // q = wi*2 + frame
                const TileDesc dsc = desc[q];
                const AxisTap cx = cws_axis(dsc.c0 + j, dsc.vx);
                xw[e] = make_float2(cx.w1, cx.w0);
                xf[e] = (cx.exact ? 2 : 0) | ((cx.lo - (dsc.ox + j)) & 1);
                flag |= cx.exact;
                // rows: every row index appears as some j (square windows) -> same loop covers them
                const AxisTap cy = cws_axis(dsc.r0 + j, dsc.vy);
                flag |= cy.exact;
            }
            anyflag = __any_sync(FULL, flag);
            __syncwarp();
        }

        // ===================================================================================
        // Six FFT steps per lane around ONE shared transform body (the unrolled W-point FFT is the
        // bulk of the code; sharing it keeps every phase's working set inside the instruction cache):
        //   s = 0, 1  row FFTs of z = a + i b                  (phase R)  -> M
        //   s = 2, 3  forward FFTs of the column pair (k, W-k) (phase C)
        //   s = 4     inverse column FFT of the product        (phase C)  -> Q
        //   s = 5     inverse FFT of two packed Hermitian rows (phase I)  -> map rows in registers
        // ===================================================================================
        const int wi = lane / HALF;            // window of this lane in phases C, I, E
        const int kc = lane % HALF;            // column pair (kc, W-kc); kc == 0: columns 0 and W/2
        const int l = kc;                      // rows l and l + W/2 in phase I
        const int col1 = kc, col2 = kc ? W - kc : HALF;
        float2* const Mw = reinterpret_cast<float2*>(smem + S::EX_OFF + wi * G::MB);
        float sum_a = 0.f, sum_b = 0.f;        // pixel sums of the window (valid in lanes kc == 0)
        // bit w set: window w of the job is exactly constant in frame a (b).  The reference then gets
        // an exactly constant correlation map (all ties -> peak index 0, ratio 1); packing a + i b
        // into one transform would leak ~1e-7 of the other frame into it, so such windows are
        // detected here and their map is forced to zero.
        unsigned const_a = 0xffffffffu, const_b = 0xffffffffu;
        float lead_a = 0.f, lead_b = 0.f;
        float2 x[W];                           // the FFT operand
        float2 xs[(W <= 32) ? W : 1];          // W <= 32: spectrum of column k kept while column W-k is transformed
        constexpr int NSTEP = (SINK == SK_WIN) ? 2 : 6;
#pragma unroll 1
        for (int s = 0; s < NSTEP; ++s) {
            if ((p.sync_mask >> s) & 1) __syncthreads();        // lock step: shared instruction fetch
            // ---------------------------------------------------------------- load
            const int grow = lane + 32 * s;
            const int rwi = (grow >> LOGW) & (NW - 1), rt = grow & (W - 1);     // row mapping (s < 2)
            const int g = min(job_c * NW + rwi, n_total - 1);
            if (s < 2) {
                if constexpr (LOADER == LD_FRAME_INT) {
                    const unsigned char* ta = smem + S::TILE_OFF + (rwi * 2) * T::BYTES;
                    const unsigned char* tb = ta + T::BYTES;
                    const TileDesc* desc = reinterpret_cast<const TileDesc*>(smem + S::TD_OFF) + buf * (NW * 2);
                    uint32_t wa[W / 4], wb[W / 4];
                    load_row_words<W, LOADER, W / 4>(ta, rt, desc[rwi * 2].d, wa);
                    load_row_words<W, LOADER, W / 4>(tb, rt, desc[rwi * 2 + 1].d, wb);
                    static_for<0, W>([&](auto jc) {
                        constexpr int j = decltype(jc)::value;
                        x[j] = make_float2(u8f(wa[j >> 2], j & 3), u8f(wb[j >> 2], j & 3));
                    });
                } else if constexpr (LOADER == LD_FRAME_CWS) {
                    const TileDesc* desc = reinterpret_cast<const TileDesc*>(smem + S::TD_OFF) + buf * (NW * 2);
                    const float2* xw = reinterpret_cast<const float2*>(smem + S::XW_OFF);
                    const int* xf = reinterpret_cast<const int*>(smem + S::XF_OFF);
                    static_for<0, 2>([&](auto fc) {
                        constexpr int frame = decltype(fc)::value;
                        const TileDesc dsc = desc[rwi * 2 + frame];
                        const AxisTap cy = cws_axis(dsc.r0 + rt, dsc.vy);
                        const int jy = (cy.lo - (dsc.oy + rt)) & 1;
                        const unsigned char* tile = smem + S::TILE_OFF + (rwi * 2 + frame) * T::BYTES;
                        const int d = dsc.d;
                        uint32_t r0w[W / 4 + 1], r1w[W / 4 + 1];
                        load_row_words<W, LOADER, W / 4 + 1>(tile, rt, d, r0w);
                        load_row_words<W, LOADER, W / 4 + 1>(tile, rt + 1, d, r1w);
                        const float2* xwq = xw + (rwi * 2 + frame) * W;
                        const int* xfq = xf + (rwi * 2 + frame) * W;
                        float l0 = u8f(r0w[0], 0), l1 = u8f(r1w[0], 0);
                        if (!anyflag) {
                            static_for<0, W>([&](auto jc) {
                                constexpr int j = decltype(jc)::value;
                                const float h0 = u8f(r0w[(j + 1) >> 2], (j + 1) & 3);
                                const float h1 = u8f(r1w[(j + 1) >> 2], (j + 1) & 3);
                                const float2 wx = xwq[j];
                                // PB:187-192 evaluation order, no FMA contraction
                                float acc = __fmul_rn(__fmul_rn(l0, wx.x), cy.w1);
                                acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(h0, wx.y), cy.w1));
                                acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(l1, wx.x), cy.w0));
                                acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(h1, wx.y), cy.w0));
                                if constexpr (frame == 0) x[j].x = acc; else x[j].y = acc;
                                l0 = h0; l1 = h1;
                            });
                        } else {
                            static_for<0, W>([&](auto jc) {
                                constexpr int j = decltype(jc)::value;
                                const float h0 = u8f(r0w[(j + 1) >> 2], (j + 1) & 3);
                                const float h1 = u8f(r1w[(j + 1) >> 2], (j + 1) & 3);
                                const float2 wx = xwq[j];
                                const int fl = xfq[j];
                                float acc = __fmul_rn(__fmul_rn(l0, wx.x), cy.w1);
                                acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(h0, wx.y), cy.w1));
                                acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(l1, wx.x), cy.w0));
                                acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(h1, wx.y), cy.w0));
                                // exact-integer coordinate on either axis: tap (floor y, floor x) (PB:170, 193)
                                const float s0 = (fl & 1) ? h0 : l0, s1 = (fl & 1) ? h1 : l1;
                                const float q11 = jy ? s1 : s0;
                                acc = ((fl & 2) || cy.exact) ? q11 : acc;
                                if constexpr (frame == 0) x[j].x = acc; else x[j].y = acc;
                                l0 = h0; l1 = h1;
                            });
                        }
                    });
                } else if constexpr (LOADER == LD_EXPL_F32) {
                    const float4* ra = reinterpret_cast<const float4*>(
                        static_cast<const float*>(p.wa) + (static_cast<long long>(g) * W + rt) * W);
                    const float4* rb = reinterpret_cast<const float4*>(
                        static_cast<const float*>(p.wb) + (static_cast<long long>(g) * W + rt) * W);
                    static_for<0, W / 4>([&](auto cc) {
                        constexpr int c = decltype(cc)::value;
                        const float4 va = __ldg(ra + c), vb = __ldg(rb + c);
                        x[4 * c] = make_float2(va.x, vb.x);
                        x[4 * c + 1] = make_float2(va.y, vb.y);
                        x[4 * c + 2] = make_float2(va.z, vb.z);
                        x[4 * c + 3] = make_float2(va.w, vb.w);
                    });
                } else {
                    const uint4* ra = reinterpret_cast<const uint4*>(
                        static_cast<const unsigned char*>(p.wa) + (static_cast<long long>(g) * W + rt) * W);
                    const uint4* rb = reinterpret_cast<const uint4*>(
                        static_cast<const unsigned char*>(p.wb) + (static_cast<long long>(g) * W + rt) * W);
                    static_for<0, W / 16>([&](auto cc) {
                        constexpr int c = decltype(cc)::value;
                        const uint4 qa = __ldg(ra + c), qb = __ldg(rb + c);
                        const uint32_t wa[4] = {qa.x, qa.y, qa.z, qa.w};
                        const uint32_t wb[4] = {qb.x, qb.y, qb.z, qb.w};
                        static_for<0, 16>([&](auto bc) {
                            constexpr int b = decltype(bc)::value;
                            x[16 * c + b] = make_float2(u8f(wa[b >> 2], b & 3), u8f(wb[b >> 2], b & 3));
                        });
                    });
                }

            } else if (s == 2) {
                static_for<0, W>([&](auto tc) { constexpr int t = decltype(tc)::value; x[t] = Mw[t * PM + col1]; });
            } else if (s == 3) {
                static_for<0, W>([&](auto tc) { constexpr int t = decltype(tc)::value; x[t] = Mw[t * PM + col2]; });
            } else if (s == 5) {
                // two Hermitian rows l, l + W/2 of Q packed into one complex inverse FFT
                const float2* Qw = Mw;
                const float2 a0 = Qw[l * PQ], b0 = Qw[(l + HALF) * PQ];
                x[0] = make_float2(b0.x, a0.x);
                x[HALF] = make_float2(b0.y, a0.y);
                static_for<1, HALF>([&](auto kc_) {
                    constexpr int k = decltype(kc_)::value;
                    const float2 R1 = Qw[l * PQ + k], R2 = Qw[(l + HALF) * PQ + k];
                    x[k] = make_float2(R1.y + R2.x, R1.x - R2.y);
                    x[W - k] = make_float2(R2.x - R1.y, R1.x + R2.y);
                });
                __syncwarp();                   // Q fully read before the map overwrites it
            }

            if (s < 2 && SINK == SK_DISP) {
                uint32_t da = 0u, db = 0u;
                static_for<1, W>([&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    da |= __float_as_uint(x[j].x) ^ __float_as_uint(x[0].x);
                    db |= __float_as_uint(x[j].y) ^ __float_as_uint(x[0].y);
                });
                constexpr int SEG = (W < 32) ? W : 32;             // lanes holding rows of one window
                const int seg_lane0 = lane & ~(SEG - 1);
                // first pixel of the window: row 0 is read in step 0 for W = 64, in this step otherwise
                if (W < 64 || s == 0) {
                    lead_a = __shfl_sync(FULL, x[0].x, seg_lane0);
                    lead_b = __shfl_sync(FULL, x[0].y, seg_lane0);
                }
                const unsigned oka = __ballot_sync(FULL, da == 0u && x[0].x == lead_a);
                const unsigned okb = __ballot_sync(FULL, db == 0u && x[0].y == lead_b);
#pragma unroll
                for (int k = 0; k < 32 / SEG; ++k) {
                    const unsigned segmask = (SEG == 32) ? 0xffffffffu : (((1u << SEG) - 1u) << (k * SEG));
                    const int wdx = ((32 * s + k * SEG) >> LOGW) & (NW - 1);
                    if ((oka & segmask) != segmask) const_a &= ~(1u << wdx);
                    if ((okb & segmask) != segmask) const_b &= ~(1u << wdx);
                }
            }

            if constexpr (SINK == SK_WIN) {
                if (active && job_c * NW + rwi < n_total) {
                    float4* oa = reinterpret_cast<float4*>(p.win_a_out + (static_cast<long long>(g) * W + rt) * W);
                    float4* ob = reinterpret_cast<float4*>(p.win_b_out + (static_cast<long long>(g) * W + rt) * W);
                    static_for<0, W / 4>([&](auto cc) {
                        constexpr int c = decltype(cc)::value;
                        oa[c] = make_float4(x[4 * c].x, x[4 * c + 1].x, x[4 * c + 2].x, x[4 * c + 3].x);
                        ob[c] = make_float4(x[4 * c].y, x[4 * c + 1].y, x[4 * c + 2].y, x[4 * c + 3].y);
                    });
                }
                if (s == 1) {
                    __syncwarp();
                    if (base + job_stride < njobs) stage_tiles(min(job + job_stride, njobs - 1), buf ^ 1);
                }
                continue;
            }

            F::run(x);

            // ---------------------------------------------------------------- store
            if (s < 2) {
                float2* Mrow = reinterpret_cast<float2*>(smem + S::EX_OFF + rwi * G::MB) + rt * PM;
                static_for<0, W>([&](auto kc_) {
                    constexpr int k = decltype(kc_)::value;
                    Mrow[k] = x[F::pos(k)];
                });
                if (s == 1) {
                    __syncwarp();
                    // tiles are consumed: request the next job's while this one is transformed
                    if (base + job_stride < njobs) stage_tiles(min(job + job_stride, njobs - 1), buf ^ 1);
                }
            } else if (s == 2) {
                if constexpr (W <= 32) {
                    static_for<0, W>([&](auto ic) { constexpr int i = decltype(ic)::value; xs[i] = x[i]; });
                } else {
                    // W == 64: two 64-point spectra do not fit in registers; X is parked in its own
                    // (lane-private) column of M, natural order, and streamed back during the product
                    static_for<0, W>([&](auto rc) { constexpr int r = decltype(rc)::value; Mw[r * PM + col1] = x[F::pos(r)]; });
                }
            } else if (s == 3) {
                // x = spectrum Y of column W-k; X = spectrum of column k.  Product, (im, re) swapped
                // for the inverse transform, written back into x.
                auto Xn = [&](auto rc) -> float2 {
                    constexpr int r = decltype(rc)::value;
                    if constexpr (W <= 32) return xs[F::pos(r)];
                    else return Mw[r * PM + col1];
                };
                float2 pq[W];
                if (kc != 0) {
                    static_for<0, W>([&](auto rc) {
                        constexpr int r = decltype(rc)::value;
                        const float2 P = xcorr_bin(Xn(rc), x[F::pos((W - r) % W)]);
                        pq[r] = make_float2(P.y, P.x);
                    });
                } else {
                    const float2 dc = Xn(std::integral_constant<int, 0>{});
                    sum_a = dc.x;
                    sum_b = dc.y;
                    static_for<0, HALF + 1>([&](auto rc) {
                        constexpr int r = decltype(rc)::value;
                        constexpr int nr = (W - r) % W;
                        float2 P0 = xcorr_bin(Xn(rc), Xn(std::integral_constant<int, nr>{}));
                        const float2 Ph = xcorr_bin(x[F::pos(r)], x[F::pos(nr)]);
                        // drop the DC bin (mean product): a constant that `- amin` removes anyway
                        if constexpr (r == 0 && SINK == SK_DISP) P0 = make_float2(0.f, 0.f);
                        // out[r] = P0 + i Ph ; out[-r] = conj(P0) + i conj(Ph); stored (im, re)
                        pq[r] = make_float2(P0.y + Ph.x, P0.x - Ph.y);
                        if constexpr (nr != r) pq[nr] = make_float2(Ph.x - P0.y, P0.x + Ph.y);
                    });
                }
                static_for<0, W>([&](auto ic) { constexpr int i = decltype(ic)::value; x[i] = pq[i]; });
            } else if (s == 4) {
                __syncwarp();                  // M (and the parked X) fully read before Q overwrites it
                float2* Qw = Mw;
                static_for<0, W>([&](auto tc) {
                    constexpr int t = decltype(tc)::value;
                    const float2 o = x[F::pos(t)];
                    Qw[t * PQ + kc] = make_float2(o.y, o.x);
                });
                __syncwarp();
            }
        }
        if constexpr (SINK == SK_WIN) continue;

        // raw row l -> shifted row l + W/2 (values x[].y); raw row l + W/2 -> shifted row l (x[].x)
        float* mapw = reinterpret_cast<float*>(smem + S::EX_OFF + wi * G::MB);
        float mx_hi = -FLT_MAX, mx_lo = -FLT_MAX, mn = FLT_MAX;     // hi: shifted row l + HALF
        const bool degenerate = ((const_a | const_b) >> wi) & 1u;
        static_for<0, W>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            constexpr int sc = (j + HALF) % W;
            float2 o = x[F::pos(j)];
            if (SINK == SK_DISP && degenerate) o = make_float2(0.f, 0.f);
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
                const int g = job_c * NW + w2;
                if (!active || g >= n_total) break;
                const float* mw = reinterpret_cast<const float*>(smem + S::EX_OFF + w2 * G::MB);
                float* out = p.corr_out + static_cast<long long>(g) * W * W;
                for (int e = lane; e < W * W; e += 32)
                    out[e] = mw[(e >> LOGW) * PC + (e & (W - 1))] * (1.0f / G::K);
            }
            __syncwarp();
            continue;
        }

        // =============================== epilogue ==========================================
        if constexpr (SINK == SK_DISP) {
            float gmax = fmaxf(mx_hi, mx_lo), gmin = mn;
#pragma unroll
            for (int o = HALF / 2; o > 0; o >>= 1) {
                gmax = fmaxf(gmax, __shfl_xor_sync(FULL, gmax, o));
                gmin = fminf(gmin, __shfl_xor_sync(FULL, gmin, o));
            }
            // first maximum in flat (row-major) order of the shifted map (torch argmax, PB:383)
            constexpr int BIG = 1 << 20;
            int R = min(mx_lo == gmax ? l : BIG, mx_hi == gmax ? l + HALF : BIG);
#pragma unroll
            for (int o = HALF / 2; o > 0; o >>= 1) R = min(R, __shfl_xor_sync(FULL, R, o));
            R = min(R, W - 1);      // only reachable with NaN input
            int C = (mapw[R * PC + l] == gmax) ? l : ((mapw[R * PC + l + HALF] == gmax) ? l + HALF : BIG);
#pragma unroll
            for (int o = HALF / 2; o > 0; o >>= 1) C = min(C, __shfl_xor_sync(FULL, C, o));
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
            const double dmin = static_cast<double>(gmin);
            const double cm = (static_cast<double>(gmax) - dmin) + eps;
            const double cl = (static_cast<double>(at(il)) - dmin) + eps;
            const double cr = (static_cast<double>(at(ir)) - dmin) + eps;
            const double ct = (static_cast<double>(at(it_)) - dmin) + eps;
            const double cb = (static_cast<double>(at(ib)) - dmin) + eps;
            const double lm = log(cm), ll = log(cl), lr = log(cr), lt = log(ct), lb = log(cb);
            double du = static_cast<double>(C) + (lr - ll) / (2.0 * (ll + lr) - 4.0 * lm) - static_cast<double>(HALF);
            double dv = static_cast<double>(R) + (lb - lt) / (2.0 * (lb + lt) - 4.0 * lm) - static_cast<double>(HALF);
            // torch.nan_to_num (PB:418-419)
            du = isnan(du) ? 0.0 : (isinf(du) ? copysign(DBL_MAX, du) : du);
            dv = isnan(dv) ? 0.0 : (isinf(dv) ? copysign(DBL_MAX, dv) : dv);

            bool invalid = false;
            float ratio = 0.f;
            if (p.validate) {
                // second peak: maximum outside the 7x7 flat-index patch around m, each patch index
                // clamped to [0, N2-1] (PB:346-358).  Rows that cannot touch the patch reuse the
                // row maxima from registers; the <= 8 candidate rows are rescanned from smem.
                const int lo_f = m - 3 - 3 * W, hi_f = m + 3 + 3 * W;
                const int ra = max(lo_f, 0) >> LOGW, rb = min(hi_f, N2 - 1) >> LOGW;
                float s = -FLT_MAX;
                if (l < ra || l > rb) s = fmaxf(s, mx_lo);
                if (l + HALF < ra || l + HALF > rb) s = fmaxf(s, mx_hi);
                for (int rr = ra; rr <= rb; ++rr) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int cc = l + h * HALF;
                        const int f = rr * W + cc;
                        const int e = f - lo_f;                     // (i+3) + W (j+3)
                        bool in_patch = (e >= 0) && ((e & (W - 1)) <= 6) && ((e >> LOGW) <= 6);
                        in_patch |= (f == 0 && lo_f <= 0) || (f == N2 - 1 && hi_f >= N2 - 1);
                        if (!in_patch) s = fmaxf(s, mapw[rr * PC + cc]);
                    }
                }
#pragma unroll
                for (int o = HALF / 2; o > 0; o >>= 1) s = fmaxf(s, __shfl_xor_sync(FULL, s, o));
                const double c2 = (static_cast<double>(s) - dmin) + eps;
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
            const int g = job_c * NW + wi;
            if (active && l == 0 && g < n_total) {
                double uo = du + (p.base_u ? p.base_u[g] : 0.0);
                double vo = dv + (p.base_v ? p.base_v[g] : 0.0);
                if (p.pred_u) {
                    // PB:731-738: reject where the correction exceeds a positive predictor, or invalid
                    const double pu = p.pred_u[g], pv = p.pred_v[g];
                    if ((du > pu && rint(pu) > 0.0) || invalid) uo = pu;
                    if ((dv > pv && rint(pv) > 0.0) || invalid) vo = pv;
                }
                p.u[g] = uo;
                p.v[g] = vo;
                if (p.mask) p.mask[g] = invalid ? 1 : 0;
                if (p.ratio) p.ratio[g] = ratio;
            }
            __syncwarp();
        }
    }
}

}  // namespace pivb200
