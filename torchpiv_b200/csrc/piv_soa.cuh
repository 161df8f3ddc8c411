// Fused PIV pass kernel, pair-packed variant (64 / 32 / 16 px windows, displacement sink) for sm_100a: the
// kernel behind pivb200_pass_first / pivb200_pass_next (OfflinePIV, PIVPlan).
//
// Same job as piv_fused_kernel (window extraction by TMA -> CWS / DWS window shift -> 2-D cross-correlation
// -> fft-shifted peak search, sub-pixel fit, peak-ratio validation, predictor glue; PB:147-257, 346-422,
// 728-738 / 800-810) and the same decomposition -- persistent CTAs, a warp owns a job of 64 / W windows,
// H = W/2 lanes per window -- but every lane runs H-point transforms on PAIRS (fft_soa.cuh,
// piv_soa_math.cuh) instead of one W-point transform on (re, im)-packed registers:
//
//   R   lane l transforms window rows (2l, 2l+1) as the pair: real rows -> H-point complex FFT + split
//       step -> planes X [row pair l][column c] of real pairs and of imaginary pairs in shared memory (STS.64)
//   C   lane c reads column c: the pair is (even rows, odd rows); H-point FFT + radix-2 step across the
//       pair -> (Y[q], Y[q+H]).  Frame a's spectrum is PARKED in tensor memory (lane-private scratch:
//       tcgen05.st / ld) while frame b goes through R and C
//   P   conj(A^) B^ (packed); column 0 (= the two real columns 0 and W/2) is separated with the help of
//       the other lanes through a scratch area
//   C'  radix-2 step across the pair, inverse H-point FFT -> planes Q [row pair m][column c]
//   R'  lane l reads the half spectra of rows (2l, 2l+1), inverse real transform -> two rows of the map
//   E   epilogue: min / first-max / flat neighbours / second peak / FP32 ratio fit on the fft-shifted map
//
// Versus piv_fused_kernel: all FFT arithmetic is packed FADD2 / FMUL2 / FFMA2 with immediate twiddles (the
// radix-2 steps across a pair are the only scalar FP32 work), 13-22 % fewer executed instructions, 20 instead of 16
// resident warps at 32 px.  The map-dependent part of the epilogue runs per job; its scalar tail (sub-pixel fit,
// ratio test, predictor replacement, stores) is batched: soa_flush_tail handles 16 queued windows, one per lane.
// What bounds the kernel is latency per warp-instruction at 3-5 resident warps per scheduler (register file and
// shared memory are both full) on top of an FP32 pipe that the transform phases saturate with one or two warps.
// Instruction supply (per-job code 46-72 KB against 32 KB of L1.5) is the second-order term: the warps of one
// scheduler run in lock step from one meeting point per job (soa_launch.cuh) and share their fetches, worth 8 % at
// 64 px.  DESIGN.md section 3.1 has the measurements and the list of restructurings that were tried.
#pragma once
#include "piv_fused.cuh"
#include "piv_soa_math.cuh"

namespace pivb200 {

// byte b of `word` as a float without the quarter-rate XU pipe: PRMT (ALU) isolates the byte, I2FP.F32.S32 (a
// full-rate FMA-pipe conversion; inline PTX keeps ptxas from fusing the pair back into the XU's I2F.U8) converts it.
// PIVB200_XU_CVT=1 (compile time) restores I2F.U8 for A/B runs.
__device__ __forceinline__ float u8g(uint32_t word, int b) {
#ifdef PIVB200_XU_CVT
    return u8f(word, b);
#else
    const uint32_t x = __byte_perm(word, 0u, 0x4440u + static_cast<uint32_t>(b));
    float f;
    asm("cvt.rn.f32.s32 %0, %1;" : "=f"(f) : "r"(x));
    return f;
#endif
}

template <int W>
struct GeoS {
    static_assert(W == 16 || W == 32 || W == 64, "window size");
    static constexpr int NW = 64 / W;                 // windows per warp job
    static constexpr int H = W / 2;                   // lanes per window = transform length
    static constexpr int LOGW = (W == 64) ? 6 : ((W == 32) ? 5 : 4);
    static constexpr int PQ = H + 1;                  // pitch of X / Q rows in pairs (8 B); odd -> conflict free
    static constexpr int PLANE = H * PQ;              // pairs per plane: real parts first, imaginary parts behind them
    static constexpr int XB = 2 * PLANE * 8;
    static constexpr int PM = W + 1;                  // pitch of a map row PAIR in float2
    static constexpr int MAPB = H * PM * 8;
    static constexpr int SCRB = 6 * H * 8;            // column-0 scratch: (re, im) pair arrays of the a, b spectra and of V
    // the windows of a warp keep their data at staggered offsets so that scalar accesses of different windows
    // hit different banks; 128-bit accesses are issued per quarter warp = per window anyway
    static constexpr int STAGGER = 64 * (NW - 1);
    __host__ __device__ static constexpr int doff(int wi) { return wi * 64; }
    static constexpr int MAXB = XB > MAPB ? (XB > SCRB ? XB : SCRB) : (MAPB > SCRB ? MAPB : SCRB);
    static constexpr int REGION = ((MAXB + STAGGER + 255) / 256) * 256;
    static constexpr float K = 4.0f * W * W;          // map = K * sum_x a(x) b(x+s)
    static constexpr int TCOLS = 2 * W;               // TMEM columns one warp parks (H elements x 4 words)
};

// NROWS consecutive tile rows starting at `row0`, each realigned to start at byte `d` (0..15) of the staged row.
// The word part of d is the same for all lanes of a window, so a switch over statically indexed registers
// (uniform per half / quarter warp) replaces the two-level select network of load_row_words; the byte part is one
// funnel shift per word.
template <int W, int LOADER, int NROWS, int NOUT>
__device__ __forceinline__ void load_rows_realigned(const unsigned char* tile, int row0, int d, uint32_t (&out)[NROWS][NOUT]) {
    using T = Tile<W, LOADER>;
    constexpr int NL = T::BX / 4;
    static_assert(NOUT + 3 <= NL + 1, "one zero word pads the row");
    uint32_t L[NROWS][NL + 1];
    static_for<0, NROWS>([&](auto rc) {
        constexpr int r = decltype(rc)::value;
        static_for<0, T::BX / 16>([&](auto cc) {
            constexpr int c = decltype(cc)::value;
            const uint4 q = *reinterpret_cast<const uint4*>(tile + T::off(row0 + r, c));
            L[r][4 * c] = q.x; L[r][4 * c + 1] = q.y; L[r][4 * c + 2] = q.z; L[r][4 * c + 3] = q.w;
        });
        L[r][NL] = 0u;
    });
    const int sh = (d & 3) * 8;
    auto realign = [&](auto sc) {
        constexpr int s = decltype(sc)::value;
        static_for<0, NROWS>([&](auto rc) {
            constexpr int r = decltype(rc)::value;
            static_for<0, NOUT>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                out[r][k] = __funnelshift_r(L[r][k + s], L[r][k + s + 1], sh);
            });
        });
    };
    switch (d >> 2) {
        case 0: realign(std::integral_constant<int, 0>{}); break;
        case 1: realign(std::integral_constant<int, 1>{}); break;
        case 2: realign(std::integral_constant<int, 2>{}); break;
        default: realign(std::integral_constant<int, 3>{}); break;
    }
}

template <int W, int LOADER>
struct SmemS {
    using G = GeoS<W>;
    using T = Tile<W, LOADER>;
    static_assert(LOADER == LD_FRAME_INT || LOADER == LD_FRAME_CWS || LOADER == LD_FRAME_ALN, "frame loaders only");
    static_assert(T::TX <= G::REGION, "the tile is staged inside the window's buffer");
    static constexpr int REG_OFF = 0;
    static constexpr int XW_OFF = REG_OFF + G::NW * G::REGION;               // float4 [NW][W]  (CWS column taps, each weight twice)
    static constexpr int XW = (LOADER == LD_FRAME_CWS) ? G::NW * W * 16 : 0;
    static constexpr int XF_OFF = XW_OFF + XW;                               // int    [NW][W]
    static constexpr int XF = (LOADER == LD_FRAME_CWS) ? G::NW * W * 4 : 0;
    static constexpr int TD_OFF = ((XF_OFF + XF + 15) / 16) * 16;            // TileDesc [NW][2]
    static constexpr int TD = G::NW * 2 * 32;
    static constexpr int BAR_OFF = ((TD_OFF + TD + 7) / 8) * 8;
    static constexpr int TOTAL = BAR_OFF + 8;
    static constexpr int STRIDE = ((TOTAL + T::BASE_ALIGN - 1) / T::BASE_ALIGN) * T::BASE_ALIGN;
    static constexpr int SMEM_MAX = 232448 - 1024;
#ifndef PIVB200_S32_WARPS
#define PIVB200_S32_WARPS 20
#endif
#ifndef PIVB200_S16_WARPS
#define PIVB200_S16_WARPS 28
#endif
#ifndef PIVB200_S64_WARPS
#define PIVB200_S64_WARPS 12
#endif
    static constexpr int WARP_CAP = (W == 64) ? PIVB200_S64_WARPS : ((W == 32) ? PIVB200_S32_WARPS : PIVB200_S16_WARPS);
    static constexpr int TMEM_CAP = 4 * (512 / G::TCOLS);
    static constexpr int BY_SMEM = SMEM_MAX / STRIDE;
    static constexpr int NWARPS0 = BY_SMEM < WARP_CAP ? BY_SMEM : WARP_CAP;
    static constexpr int NWARPS = NWARPS0 < TMEM_CAP ? NWARPS0 : TMEM_CAP;
    // peak records waiting for the scalar tail of the epilogue (soa_flush_tail): QN records of 40 bytes per warp,
    // behind the warp slots
    static constexpr int QN = (W == 16) ? 8 : 16;
    static constexpr int QREC = 40;
    static constexpr int Q_OFF = NWARPS * STRIDE;
    static constexpr int CTA_BYTES = Q_OFF + NWARPS * QN * QREC;
    static_assert(QN % G::NW == 0, "whole jobs per flush");
    static_assert(CTA_BYTES <= SMEM_MAX, "shared memory");
};

// What the map-dependent part of the epilogue leaves per window; everything else (sub-pixel fit, ratio test,
// predictor replacement, stores: PB:394-422, 728-738) is scalar work that soa_flush_tail does for QN windows at once,
// one window per lane, instead of once per job with all lanes of a window computing the same numbers.
struct PeakRec {
    int g;              // global window index, -1 = padding
    int m;              // flat index of the peak in the fft-shifted map
    float cm0;          // peak - min
    float fl, fr, ft, fb;   // flat neighbours m+1, m-1, m+W, m-W (guarded, PB:385-392), each minus min
    float spd;          // second peak - min
    float sa, sb;       // pixel sums of the two windows
};
static_assert(sizeof(PeakRec) == 40, "PeakRec layout");

template <int W>
__device__ __noinline__ void soa_flush_tail(const PassParams& p, const PeakRec* q, int count, int lane) {
    using G = GeoS<W>;
    constexpr int H = G::H, LOGW = G::LOGW, N2 = W * W;
    if (lane < count) {
        const PeakRec r = q[lane];
        const int g = r.g;
        if (g >= 0) {
            // eps of PB:381 in the map's scale; pass 1 divides the frames by their means (PB:513-514), which scales eps
            float eps = G::K * 1e-7f;
            if (p.first_pass) eps *= (r.sa * (1.0f / N2)) * (r.sb * (1.0f / N2));
            // Three-point log-Gaussian fit (PB:394-407) in FP32 on RATIOS to the peak value:
            //   (ln c- - ln c+) / (2 ln c+ + 2 ln c- - 4 ln c0) = (B - A) / (2 (A + B)),  A = ln(c+ / c0), B = ln(c- / c0),
            // which needs no FP64: the differences of logarithms are formed directly instead of by cancellation.
            // Agreement with the FP64 evaluation of the reference: ~1e-7 px, below what the FP32 rounding of the map
            // itself contributes.  IEEE division: c / c == 1 exactly (featureless windows).
            const float cm = r.cm0 + eps;
            const float ll = logf((r.fl + eps) / cm), lr = logf((r.fr + eps) / cm);
            const float lt = logf((r.ft + eps) / cm), lb = logf((r.fb + eps) / cm);
            const float fu = (lr - ll) / (2.0f * (ll + lr)), fv = (lb - lt) / (2.0f * (lb + lt));
            const int R = r.m >> LOGW, C = r.m & (W - 1);
            // torch.nan_to_num (PB:418-419)
            double du = isnan(fu) ? 0.0 : (isinf(fu) ? copysign(DBL_MAX, static_cast<double>(fu))
                                                     : static_cast<double>(C - H) + static_cast<double>(fu));
            double dv = isnan(fv) ? 0.0 : (isinf(fv) ? copysign(DBL_MAX, static_cast<double>(fv))
                                                     : static_cast<double>(R - H) + static_cast<double>(fv));
            bool invalid = false;
            float ratio = 0.f;
            if (p.validate) {
                ratio = cm / (r.spd + eps);
                invalid = static_cast<double>(ratio) < p.val_ratio;
            }
            if (p.first_pass && (r.sa == 0.f || r.sb == 0.f)) {
                // black window: the reference divides by a zero mean (PB:513-514), every value is NaN,
                // nan_to_num gives 0 and the NaN ratio compares False (valid)
                du = dv = 0.0;
                invalid = false;
                ratio = 0.f;
            }
            double uo = du, vo = dv;
            if (p.base_u) { uo += p.base_u[g]; vo += p.base_v[g]; }
            if (p.pred_u) {
                // PB:731-738: reject where the correction exceeds a positive predictor, or invalid
                const double pred_u = p.pred_u[g], pred_v = p.pred_v[g];
                if ((du > pred_u && rint(pred_u) > 0.0) || invalid) uo = pred_u;
                if ((dv > pred_v && rint(pred_v) > 0.0) || invalid) vo = pred_v;
            }
            p.u[g] = uo;
            p.v[g] = vo;
            if (p.mask) p.mask[g] = invalid ? 1 : 0;
            if (p.ratio) p.ratio[g] = ratio;
        }
    }
    __syncwarp();
}


// Request the tiles of `frame` of the warp's current job (TileDesc table in shared memory): TMA for the windows that
// overlap the frame, byte gather for the rest.  The mbarrier is armed for EVERY (job, frame), also with nothing to
// wait for, so its phase parity at the matching wait is simply `frame`.  Out of line on purpose: three call sites,
// and the per-job code has to stay small (instruction cache).
template <int W, int LOADER>
__device__ __noinline__ void stage_tiles(const CUtensorMap* tmA, const CUtensorMap* tmB, const PassParams& p,
                                         unsigned char* smem, int frame, int lane) {
    using G = GeoS<W>;
    using T = Tile<W, LOADER>;
    using S = SmemS<W, LOADER>;
    constexpr int NW = G::NW;
    TileDesc* desc = reinterpret_cast<TileDesc*>(smem + S::TD_OFF);
    const uint32_t bar = smem_u32(smem + S::BAR_OFF);
    fence_proxy_async();            // generic accesses to the buffers are ordered before the TMA writes
    __syncwarp();
    // lane w2 < NW owns window w2 of the job
    const int w2 = lane & (NW - 1);
    const bool mine = lane < NW;
    const TileDesc dsc = desc[w2 * 2 + frame];
    const unsigned by_tma = __ballot_sync(0xffffffffu, mine && dsc.d >= 0);
    const unsigned by_gather = __ballot_sync(0xffffffffu, mine && dsc.d < 0);
    if (by_gather) {                // windows entirely outside the frame (rare): the whole warp gathers them
#pragma unroll 1
        for (int w3 = 0; w3 < NW; ++w3)
            if ((by_gather >> w3) & 1)
                gather_border_tile<W, LOADER>(p, desc[w3 * 2 + frame], smem + S::REG_OFF + w3 * G::REGION, frame, lane);
    }
    if (lane == 0) mbar_arrive_expect_tx(bar, static_cast<uint32_t>(__popc(by_tma)) * T::TX);
    __syncwarp();
    if (mine && dsc.d >= 0)
        tma_load_3d(smem_u32(smem + S::REG_OFF + w2 * G::REGION), frame ? tmB : tmA, bar, dsc.ox & ~15, dsc.oy, dsc.pair);
    if (mine && dsc.d < 0) desc[w2 * 2 + frame].d = 0;     // gathered tiles start at byte 0
    __syncwarp();
}

template <int W, int LOADER>
__global__ void __launch_bounds__(SmemS<W, LOADER>::NWARPS * 32, 1) piv_soa_kernel(const __grid_constant__ CUtensorMap tmA,
                                                        const __grid_constant__ CUtensorMap tmB,
                                                        const __grid_constant__ PassParams p) {
    using G = GeoS<W>;
    using T = Tile<W, LOADER>;
    using S = SmemS<W, LOADER>;
    using M = SoaMath<W>;
    constexpr int NW = G::NW, H = G::H, LOGW = G::LOGW, PQ = G::PQ, PM = G::PM;
    constexpr unsigned FULL = 0xffffffffu;

    extern __shared__ __align__(1024) unsigned char smem_cta[];
    __shared__ uint32_t tmem_base_sh;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    unsigned char* smem = smem_cta + warp * S::STRIDE;
    const int n_total = static_cast<int>(p.n_total);
    const int njobs = (n_total + NW - 1) / NW;
    const int job_stride = gridDim.x * nwarps;
    const uint32_t bar = smem_u32(smem + S::BAR_OFF);

    if (warp == 0) tmem_alloc_512(smem_u32(&tmem_base_sh));
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();
    // parking area of this warp: its 32 lanes, TCOLS columns per group of four warps
    const uint32_t tpark = tmem_base_sh + (static_cast<uint32_t>((warp & 3) * 32) << 16) +
                           static_cast<uint32_t>((warp >> 2) * G::TCOLS);
    if (lane == 0) mbar_init(bar, 1);
    __syncwarp();

    const int wi = lane / H;               // window of this lane inside the job
    const int l = lane % H;                // row pair l in phases R and R', column l in phases C and C'
    unsigned char* const region = smem + S::REG_OFF + wi * G::REGION;
    // X / Q of this lane's window: plane of real pairs [H][PQ], plane of imaginary pairs behind it.  (Two 64-bit
    // accesses per element instead of one 128-bit access: a quad would need the two pairs in four consecutive
    // registers, which costs four MOVs per element -- measured, profiles/r02a.)
    float2* const Xr = reinterpret_cast<float2*>(region + G::doff(wi));
    float2* const Xi = Xr + G::PLANE;

    auto make_desc = [&](int job) {
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
            const bool partial = p.use_tma && !interior && dsc.ox + T::USED > 0 && dsc.ox < p.Wf &&
                                 dsc.oy + T::BY > 0 && dsc.oy < p.H;
            dsc.d = interior ? (dsc.ox & 15) : (partial ? ((dsc.ox & 15) | kPatchFlag) : -1);
            desc[q] = dsc;
        }
        __syncwarp();
    };
    auto stage_issue = [&](int frame) { stage_tiles<W, LOADER>(&tmA, &tmB, p, smem, frame, lane); };

    // optional lock step (instruction-fetch sharing): bit i of p.sync_mask puts a named barrier over groups of
    // p.sync_group warps at phase boundary i (0 rows of frame a, 1 columns, 2 product, 3 inverse columns, 4 inverse rows,
    // 5 epilogue, 6 rows of frame b)
    auto lockstep = [&](int point) {
        if ((p.sync_mask >> point) & 1) {
            // sync_group == 0: the warps of one scheduler (warp & 3) form a group -- they share the scheduler's L0
            // instruction cache; otherwise groups of sync_group consecutive warps
            if (p.sync_group == 0) asm volatile("bar.sync %0, %1;" ::"r"(1 + (warp & 3)), "r"((nwarps >> 2) * 32) : "memory");
            else asm volatile("bar.sync %0, %1;" ::"r"(1 + warp / p.sync_group), "r"(p.sync_group * 32) : "memory");
        }
    };
    PeakRec* const queue = reinterpret_cast<PeakRec*>(smem_cta + S::Q_OFF + warp * (S::QN * S::QREC));
    int queued = 0;
    bool prefetched = false;
#pragma unroll 1
    for (int base = blockIdx.x * nwarps; base < njobs; base += job_stride) {
        const int job = min(base + warp, njobs - 1);
        const int g = min(job * NW + wi, n_total - 1);
        const bool g_valid = (base + warp < njobs) && (job * NW + wi < n_total);
        if (!prefetched) {
            make_desc(job);
            stage_issue(0);
        }

        float sum_a = 0.f, sum_b = 0.f;        // pixel sums of the window (valid in lanes l == 0)
        float2 x[W];
#pragma unroll 1
        for (int frame = 0; frame < 2; ++frame) {
            // ------------------------------------------------------------- tiles -> rows (2l, 2l+1) in x[j]
            lockstep(frame == 0 ? 0 : 6);
            mbar_wait(bar, static_cast<uint32_t>(frame));
            const TileDesc* desc = reinterpret_cast<const TileDesc*>(smem + S::TD_OFF);
            if constexpr (LOADER == LD_FRAME_INT || LOADER == LD_FRAME_CWS) {
#pragma unroll 1
                for (int w2 = 0; w2 < NW; ++w2) {
                    const TileDesc dsc = desc[w2 * 2 + frame];
                    if (dsc.d >= kPatchFlag)
                        patch_border_tile<W, LOADER>(p, dsc, smem + S::REG_OFF + w2 * G::REGION, frame, lane);
                }
                __syncwarp();
            }
            if constexpr (LOADER == LD_FRAME_INT || LOADER == LD_FRAME_ALN) {
                const int d = desc[wi * 2 + frame].d & 15;
                uint32_t w[2][W / 4];
                if constexpr (LOADER == LD_FRAME_ALN) {
                    load_row_words<W, LOADER, W / 4>(region, 2 * l, d, w[0]);
                    load_row_words<W, LOADER, W / 4>(region, 2 * l + 1, d, w[1]);
                } else {
                    load_rows_realigned<W, LOADER, 2, W / 4>(region, 2 * l, d, w);
                }
                static_for<0, W>([&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    x[j] = make_float2(u8g(w[0][j >> 2], j & 3), u8g(w[1][j >> 2], j & 3));
                });
            } else {
                float4* xw = reinterpret_cast<float4*>(smem + S::XW_OFF);
                int* xf = reinterpret_cast<int*>(smem + S::XF_OFF);
                // per-column tap descriptors of this frame (shared by all rows of a window).  (A variant that keeps the two
                // weights in registers when all columns of the window lie in one float binade -- then they are identical --
                // was measured: no gain, and the second copy of the tap loop costs instruction-cache space.)
                bool flag = false;
#pragma unroll
                for (int e = lane; e < NW * W; e += 32) {
                    const int j = e & (W - 1), w2 = e >> LOGW;
                    const TileDesc d2 = desc[w2 * 2 + frame];
                    const AxisTap cx = cws_axis(d2.c0 + j, d2.vx);
                    xw[e] = make_float4(cx.w1, cx.w1, cx.w0, cx.w0);      // pairs: operands of the packed taps
                    xf[e] = (cx.exact ? 2 : 0) | ((cx.lo - (d2.ox + j)) & 1);
                    flag |= cx.exact;
                }
                __syncwarp();
                const TileDesc dsc = desc[wi * 2 + frame];
                const float4* xwq = xw + wi * W;
                const int* xfq = xf + wi * W;
                const int ra = 2 * l;
                const AxisTap cyA = cws_axis(dsc.r0 + ra, dsc.vy), cyB = cws_axis(dsc.r0 + ra + 1, dsc.vy);
                // general = some coordinate of the job is an exact integer (scalar tap loop below)
                const bool general = __any_sync(FULL, flag || cyA.exact || cyB.exact);
                uint32_t wABC[3][W / 4 + 1];
                load_rows_realigned<W, LOADER, 3, W / 4 + 1>(region, ra, dsc.d & 15, wABC);
                const uint32_t (&wA)[W / 4 + 1] = wABC[0], (&wB)[W / 4 + 1] = wABC[1], (&wC)[W / 4 + 1] = wABC[2];
                float cA = u8g(wA[0], 0), cB = u8g(wB[0], 0), cC = u8g(wC[0], 0);
                if (!general) {
                    // per-column weights from the table, packed taps
                    const float2 wy1 = make_float2(cyA.w1, cyB.w1), wy0 = make_float2(cyA.w0, cyB.w0);
                    float2 vc = pfma(make_float2(cA, cB), wy1, pmul(make_float2(cB, cC), wy0));
                    static_for<0, W>([&](auto jc) {
                        constexpr int j = decltype(jc)::value;
                        const float nA = u8g(wA[(j + 1) >> 2], (j + 1) & 3);
                        const float nB = u8g(wB[(j + 1) >> 2], (j + 1) & 3);
                        const float nC = u8g(wC[(j + 1) >> 2], (j + 1) & 3);
                        const float2 vn = pfma(make_float2(nA, nB), wy1, pmul(make_float2(nB, nC), wy0));
                        const float4 wx = xwq[j];
                        x[j] = pfma(vc, make_float2(wx.x, wx.y), pmul(vn, make_float2(wx.z, wx.w)));
                        vc = vn;
                    });
                } else {
                    // some coordinate of the job is an exact integer: there the reference's weights all vanish and
                    // the value is patched to the tap at (floor y, floor x) (PB:170, 193)
                    const bool jyA = (cyA.lo - (dsc.oy + ra)) & 1, jyB = (cyB.lo - (dsc.oy + ra + 1)) & 1;
                    static_for<0, W>([&](auto jc) {
                        constexpr int j = decltype(jc)::value;
                        const float nA = u8g(wA[(j + 1) >> 2], (j + 1) & 3);
                        const float nB = u8g(wB[(j + 1) >> 2], (j + 1) & 3);
                        const float nC = u8g(wC[(j + 1) >> 2], (j + 1) & 3);
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
            }
            __syncwarp();                       // tiles fully read before X overwrites the buffer

            // ------------------------------------------------------------- R: rows -> X
            M::row_forward(x);
            static_for<0, H>([&](auto cc) {
                constexpr int c = decltype(cc)::value;
                constexpr int e = M::pos(c);
                Xr[l * PQ + c] = x[2 * e];
                Xi[l * PQ + c] = x[2 * e + 1];
            });
            __syncwarp();
            // ------------------------------------------------------------- C: column l
            lockstep(1);
            static_for<0, H>([&](auto tc) {
                constexpr int t = decltype(tc)::value;
                x[2 * t] = Xr[t * PQ + l];
                x[2 * t + 1] = Xi[t * PQ + l];
            });
            __syncwarp();                       // X fully read: the buffer is dead
            if (frame == 0) stage_issue(1);     // frame b's tiles arrive while column a is transformed
            M::col_forward(x);
            if (frame == 0) {
                // one pair per store: a wider store wants its registers consecutive (MOVs)
                static_for<0, W>([&](auto ic) { constexpr int i = decltype(ic)::value; tmem_st_pair(tpark + 2 * i, x[i]); });
                tmem_wait_st();
            }
        }

        // ------------------------------------------------------------- P: conj(A^) B^
        lockstep(2);
        {
            // Column 0 is the packed pair of real columns (bins 0 and W/2): C = c0^ + i cH^.  Its lane drops both
            // spectra into a scratch area AS PAIRS (element q = bins (q, q+H), real pairs and imaginary pairs in
            // separate arrays: no register shuffling), the H lanes of the window do the H + 1 small separation
            // jobs in parallel, and the lane reads the packed product back.
            float2* const sBr = Xr;                 // [H] each
            float2* const sBi = Xr + H;
            float2* const sAr = Xr + 2 * H;
            float2* const sAi = Xr + 3 * H;
            float2* const sVr = Xr + 4 * H;
            float2* const sVi = Xr + 5 * H;
            if (l == 0) {
                static_for<0, H>([&](auto qc) {
                    constexpr int q = decltype(qc)::value;
                    constexpr int e = M::pos(q);
                    sBr[q] = x[2 * e];
                    sBi[q] = x[2 * e + 1];
                });
            }
            static_for<0, W / 8>([&](auto cc) {
                constexpr int c = decltype(cc)::value;             // 4 elements = 8 pairs = 16 columns
                float2 A[8];
                tmem_ld8(tpark + 16 * c, A);
                tmem_wait_ld();
                if (l == 0) {
                    static_for<0, 4>([&](auto ic) {
                        constexpr int i = decltype(ic)::value;
                        constexpr int q = M::posinv(4 * c + i);             // element 4c + i holds bins (q, q+H)
                        sAr[q] = A[2 * i];
                        sAi[q] = A[2 * i + 1];
                    });
                }
                static_for<0, 4>([&](auto ic) {
                    constexpr int i = decltype(ic)::value;
                    constexpr int e = 4 * c + i;
                    const float2 ar = A[2 * i], ai = A[2 * i + 1], br = x[2 * e], bi = x[2 * e + 1];
                    x[2 * e] = pfma(ar, br, pmul(ai, bi));
                    x[2 * e + 1] = pfma(ar, bi, pneg(pmul(ai, br)));
                });
            });
            __syncwarp();
            {
                // bin q of a spectrum: component q & 1... no: element q % H, half q / H
                const float* fAr = reinterpret_cast<const float*>(sAr), *fAi = reinterpret_cast<const float*>(sAi);
                const float* fBr = reinterpret_cast<const float*>(sBr), *fBi = reinterpret_cast<const float*>(sBi);
                float* fVr = reinterpret_cast<float*>(sVr), *fVi = reinterpret_cast<float*>(sVi);
                // lane l > 0: bins q = l (element l, half 0) and W - l (element H - l, half 1); lane 0: bins 0 and H
                const int iq = (l == 0) ? 0 : 2 * l, in = (l == 0) ? 1 : 2 * (H - l) + 1;
                const float2 Aq = make_float2(fAr[iq], fAi[iq]), An = make_float2(fAr[in], fAi[in]);
                const float2 Bq = make_float2(fBr[iq], fBi[iq]), Bn = make_float2(fBr[in], fBi[in]);
                float2 Vq, Vn;
                if (l == 0) {
                    // bins 0 and W/2 are their own partners: everything is real
                    sum_a = 0.5f * Aq.x;
                    sum_b = 0.5f * Bq.x;
                    // the DC bin (mean product) is dropped: a constant that `- amin` removes anyway
                    Vq = make_float2(0.f, Aq.y * Bq.y);
                    Vn = make_float2(An.x * Bn.x, An.y * Bn.y);
                } else {
                    const float2 a0 = make_float2(Aq.x + An.x, Aq.y - An.y);      // 2 c0^[q]
                    const float2 ah = make_float2(Aq.y + An.y, An.x - Aq.x);      // 2 cH^[q]
                    const float2 b0 = make_float2(0.25f * (Bq.x + Bn.x), 0.25f * (Bq.y - Bn.y));
                    const float2 bh = make_float2(0.25f * (Bq.y + Bn.y), 0.25f * (Bn.x - Bq.x));
                    const float2 P0 = make_float2(fmaf(a0.x, b0.x, a0.y * b0.y), fmaf(a0.x, b0.y, -a0.y * b0.x));
                    const float2 Ph = make_float2(fmaf(ah.x, bh.x, ah.y * bh.y), fmaf(ah.x, bh.y, -ah.y * bh.x));
                    // V[q] = P0 + i Ph, V[W-q] = conj(P0) + i conj(Ph)
                    Vq = make_float2(P0.x - Ph.y, P0.y + Ph.x);
                    Vn = make_float2(P0.x + Ph.y, Ph.x - P0.y);
                }
                fVr[iq] = Vq.x; fVi[iq] = Vq.y;
                fVr[in] = Vn.x; fVi[in] = Vn.y;
            }
            __syncwarp();
            if (l == 0) {
                static_for<0, H>([&](auto qc) {
                    constexpr int q = decltype(qc)::value;
                    constexpr int e = M::pos(q);
                    x[2 * e] = sVr[q];
                    x[2 * e + 1] = sVi[q];
                });
            }
            __syncwarp();                       // scratch fully read before Q overwrites the buffer
        }

        // ------------------------------------------------------------- C': inverse column -> Q
        lockstep(3);
        M::col_inverse(x);                      // result in swapped slots: real pair in the odd one
        static_for<0, H>([&](auto mc) {
            constexpr int m = decltype(mc)::value;
            constexpr int e = M::pos(m);
            Xr[m * PQ + l] = x[2 * e + 1];
            Xi[m * PQ + l] = x[2 * e];
        });
        __syncwarp();
        // ------------------------------------------------------------- R': rows (2l, 2l+1) of the map
        lockstep(4);
        static_for<0, H>([&](auto cc) {
            constexpr int c = decltype(cc)::value;
            x[2 * c + 1] = Xr[l * PQ + c];        // swapped slots: the shared forward body then runs the inverse transform
            x[2 * c] = Xi[l * PQ + c];
        });
        __syncwarp();                           // Q fully read before the map overwrites it
        M::row_inverse(x);

        // raw rows (2l, 2l+1) -> rows (sr0, sr0 + 1) of the fft-shifted map, a row PAIR per float2:
        // element (R, C) of the shifted map lives at float index ((R >> 1) * PM + C) * 2 + (R & 1)
        float* const mapw = reinterpret_cast<float*>(Xr);
        float2* const mapw2 = Xr;
        const int sr0 = (2 * l + H) & (W - 1);
        float mx0 = -FLT_MAX, mx1 = -FLT_MAX, mn = FLT_MAX;
        static_for<0, W>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            constexpr int sc = (j + H) % W;
            const float2 o = x[2 * M::pos(j >> 1) + ((j & 1) ? 0 : 1)];      // swapped slots
            mapw2[(sr0 >> 1) * PM + sc] = o;
            mx0 = fmaxf(mx0, o.x);
            mx1 = fmaxf(mx1, o.y);
            mn = fminf(mn, fminf(o.x, o.y));
        });
        __syncwarp();

        // =============================== epilogue (PB:360-422) ===================================
        // Here: everything that needs the map.  The scalar rest happens in soa_flush_tail, QN windows at a time.
        lockstep(5);
        {
            auto at2 = [&](int R, int C) { return mapw[((R >> 1) * PM + C) * 2 + (R & 1)]; };
            const float gmax = group_max<H>(fmaxf(mx0, mx1)), gmin = group_min<H>(mn);
            // first maximum in flat (row-major) order of the shifted map (torch argmax, PB:383)
            constexpr int BIG = 1 << 20;
            int R = group_min_int<H>(min(mx0 == gmax ? sr0 : BIG, mx1 == gmax ? sr0 + 1 : BIG));
            R = min(R, W - 1);      // only reachable with NaN input
            int C = group_min_int<H>((at2(R, l) == gmax) ? l : ((at2(R, l + H) == gmax) ? l + H : BIG));
            C = min(C, W - 1);
            constexpr int N2 = W * W;
            const int m = R * W + C;
            auto at = [&](int f) { return at2(f >> LOGW, f & (W - 1)); };
            // flat neighbours, guarded only at the array ends (PB:385-392); lanes 1..4 of the window fetch one each
            const int il = (m + 1 >= N2 - 1) ? m : m + 1;
            const int ir = (m - 1 <= 0) ? m : m - 1;
            const int it_ = (m + W >= N2 - 1) ? m : m + W;
            const int ib = (m - W <= 0) ? m : m - W;
            const float fsel = at((l == 1) ? il : (l == 2) ? ir : (l == 3) ? it_ : ib) - gmin;
            // second peak: maximum outside the 7x7 flat-index patch around m, each patch index clamped to
            // [0, N2-1] (PB:346-358).  Rows that cannot touch the patch reuse the row maxima from registers; the up to
            // eight rows that can are read back, two columns per lane.  Element (rr, cc) has the patch coordinate
            // e = rr W + cc - lo_f: its column part (e mod W) depends on cc only, its row part is rr + ((cc - lo_f) >> LOGW).
            float sp = -FLT_MAX;
            if (p.validate) {
                const int lo_f = m - 3 - 3 * W, hi_f = m + 3 + 3 * W;
                const int ra = max(lo_f, 0) >> LOGW, rb = min(hi_f, N2 - 1) >> LOGW;
                if (sr0 < ra || sr0 > rb) sp = fmaxf(sp, mx0);
                if (sr0 + 1 < ra || sr0 + 1 > rb) sp = fmaxf(sp, mx1);
                const int d0 = l - lo_f, d1 = l + H - lo_f;
                const bool col0 = (d0 & (W - 1)) <= 6, col1 = (d1 & (W - 1)) <= 6;
                const int ro0 = d0 >> LOGW, ro1 = d1 >> LOGW;
                // the clamped ends: index 0 belongs to the patch when the patch reaches below 0, N2-1 when beyond the end
                const bool z0 = (lo_f <= 0) && (l == 0), zN = (hi_f >= N2 - 1) && (l == H - 1);
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int rr = ra + t;
                    if (rr <= rb) {
                        const float* row = mapw + (rr >> 1) * (2 * PM) + (rr & 1);
                        const bool in0 = (col0 && static_cast<unsigned>(rr + ro0) <= 6u) || (t == 0 && z0);
                        const bool in1 = (col1 && static_cast<unsigned>(rr + ro1) <= 6u) || (rr == rb && zN);
                        const float v0 = row[2 * l], v1 = row[2 * (l + H)];
                        sp = fmaxf(sp, in0 ? -FLT_MAX : v0);
                        sp = fmaxf(sp, in1 ? -FLT_MAX : v1);
                    }
                }
                sp = group_max<H>(sp);
            }
            const int l0 = wi * H;
            const float f_l = __shfl_sync(FULL, fsel, l0 + 1), f_r = __shfl_sync(FULL, fsel, l0 + 2),
                        f_t = __shfl_sync(FULL, fsel, l0 + 3), f_b = __shfl_sync(FULL, fsel, l0 + 4);
            if (l == 0) {
                PeakRec rec;
                rec.g = g_valid ? g : -1;
                rec.m = m;
                rec.cm0 = gmax - gmin;
                rec.fl = f_l; rec.fr = f_r; rec.ft = f_t; rec.fb = f_b;
                rec.spd = sp - gmin;
                rec.sa = sum_a; rec.sb = sum_b;
                queue[queued + wi] = rec;
            }
            queued += NW;
            // The map has been read for the last time: the next job's descriptors and frame-a tiles are requested now
            __syncwarp();
            prefetched = false;
            if (base + job_stride < njobs) {
                make_desc(min(base + job_stride + warp, njobs - 1));
                stage_issue(0);
                prefetched = true;
            }
            if (queued == S::QN) {
                soa_flush_tail<W>(p, queue, queued, lane);
                queued = 0;
            }
        }
    }
    if (queued) soa_flush_tail<W>(p, queue, queued, lane);

    tmem_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc_512(tmem_base_sh);
}

}  // namespace pivb200
