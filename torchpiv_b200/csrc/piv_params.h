// Internal (host <-> device) parameter block of the fused PIV pass kernels.
#pragma once
#include <cstdint>

namespace pivb200 {

enum Loader : int {
    LD_FRAME_INT = 0,   // windows cut from the frame with an integer shift (pass 1, DWS)
    LD_FRAME_CWS = 1,   // windows cut with a per-window float32 shift + 2x2 bilinear taps (CWS)
    LD_EXPL_F32 = 2,    // windows already materialised as [N, w, w] float32
    LD_EXPL_U8 = 3,     // ... or uint8
    LD_FRAME_ALN = 4    // LD_FRAME_INT without shifts on a grid whose step is a multiple of 16 px (the usual
                        // first pass): every tile row starts 16-byte aligned, so the TMA box is exactly the
                        // window and the loader needs no realignment network
    ,
    LD_FRAME_TC = 5     // LD_FRAME_ALN for 64 px windows with the ROW transform on the tensor cores: the uint8 rows
                        // become fp16 operands (exact) of tcgen05.mma against a hi + lo split DFT matrix, the
                        // spectra come back from tensor memory -- no row FFT and no u8 -> f32 conversion on the
                        // FP32 pipe
};

enum Sink : int {
    SK_DISP = 0,        // displacement + validation mask (the product path)
    SK_CORR = 1,        // fft-shifted correlation maps [N, w, w] float32 (correalte_fft API)
    SK_WIN = 2          // the (shifted) windows themselves [N, w, w] float32 (parity of the loader)
};

// Unsigned division by a run-time invariant (Granlund-Montgomery): q = (t + ((n - t) >> s1)) >> s2 with
// t = umulhi(M, n); exact for every 32-bit n.  Replaces the ~40-instruction emulated integer division
// in the per-job window geometry.
struct FastDiv {
    uint32_t M, s1, s2;
};
inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    uint32_t l = 0;
    while ((1ull << l) < d) ++l;                       // ceil(log2 d)
    f.M = static_cast<uint32_t>(((1ull << 32) * ((1ull << l) - d)) / d + 1);
    f.s1 = l < 1 ? l : 1;
    f.s2 = l > 1 ? l - 1 : 0;
    return f;
}

struct PassParams {
    // frames: pair p of frame a starts at fa + p * pair_stride (bytes); rows are `pitch` bytes apart
    const unsigned char* fa;
    const unsigned char* fb;
    long long pair_stride;
    int H, Wf, pitch;
    int n_rows, n_cols, step;        // window grid of one pair, window origin = (r*step, c*step)
    FastDiv div_n, div_c;            // division by n_rows * n_cols and by n_cols
    long long n_total;               // n_pairs * n_rows * n_cols (or number of explicit windows)
    int first_pass;                  // 1: eps scaled by mean(a)*mean(b) and black-window rule (PB:513-514)
    const float* sxf;                // CWS shift per window (+ for frame b, - for frame a)
    const float* syf;
    const int* sxi;                  // integer shift per window (DWS); null = 0 (pass 1)
    const int* syi;
    const double* base_u;            // u = base_u + du (null: 0)
    const double* base_v;
    const double* pred_u;            // zeroed predictor u0 for the PB:731-738 replacement (null: none)
    const double* pred_v;
    int validate;
    double val_ratio;
    double* u;
    double* v;
    unsigned char* mask;             // 1 = invalid (peak ratio test), may be null when !validate
    float* ratio;                    // optional peak / second-peak ratio
    int use_tma;
    int sync_mask;                   // bit s: barrier at phase boundary s (piv_soa.cuh: 0..6)
    int sync_group;                  // 1: barriers span groups of four warps instead of the CTA
    int skew_ns;                     // initial delay of group g: g * skew_ns
    // explicit loaders / debug sinks
    const void* wa;
    const void* wb;
    float* corr_out;
    float* win_a_out;
    float* win_b_out;
};

}  // namespace pivb200
