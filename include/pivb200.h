/* pivb200 -- C ABI of the B200-native (sm_100a) PIV cross-correlation hot path.
 *
 * Drop-in boundary for the hot path of NikNazarov/TorchPIV (reference file
 * src/torchPIV/PIVbackend.py, "PB" below).  The reference has no FFI of its own: its hot
 * path is eager PyTorch.  These entry points are what a binding of that path binds (the
 * ctypes stub a maintainer would add is shown in INTEGRATION.md; this repo's own host side,
 * torchpiv_b200/, calls exactly these symbols and nothing else).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless it says "host";
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work is
 *     enqueued asynchronously on it, nothing synchronises;
 *   - return value: 0 = ok, negative = PIVB200_E_* argument/geometry error (nothing was
 *     launched), positive = cudaError_t of a failed runtime call;
 *   - masks: 1 = INVALID vector (the reference's `validation_mask` True);
 *   - window index n = row * n_cols + col (row-major), pair index outermost: g = pair * n + n;
 *   - interrogation windows of 16, 32 or 64 px run the fused in-register FFT kernels; any other
 *     size from 4 to 256 px runs a general kernel (in-place mixed-radix FFT, radices 2-16,
 *     direct sums for sizes with a prime factor above 13; the window lives in shared memory up to
 *     160 px and in an L2-resident scratch slab above; same semantics; scratch memory comes from
 *     cudaMallocAsync on `stream`).  ODD sizes reproduce the reference's behaviour: torch.fft.irfft2
 *     without an explicit size returns a [w, w-1] correlation map for them (PB:255), and so do
 *     pivb200_correlate and the passes internally.  Sizes below 4 or above 256 px return
 *     PIVB200_E_WINDOW.  There is no CPU fallback anywhere in this library.
 */
#ifndef PIVB200_H_
#define PIVB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PIVB200_OK 0
#define PIVB200_E_WINDOW (-1)   /* window size < 4 or > 256                              */
#define PIVB200_E_OVERLAP (-2)  /* overlap >= window (PB:503-504 raises ValueError)      */
#define PIVB200_E_FRAME (-3)    /* window larger than the frame (PB:506-507), bad pitch  */
#define PIVB200_E_ARG (-4)      /* null / misaligned pointer, non-positive count         */
#define PIVB200_E_DRIVER (-5)   /* cuTensorMapEncodeTiled unavailable or failed          */
#define PIVB200_E_SIZE (-6)     /* problem too large for 32-bit window indexing          */

#define PIVB200_MODE_DWS 0
#define PIVB200_MODE_CWS 1

/* Library / build identification. */
int pivb200_version(void);
const char* pivb200_error_string(int code);

/* Field shape of one pass: (size - w) / (w - ovl) + 1 per axis.  Replaces get_field_shape, PB:425-456. */
int pivb200_field_shape(int H, int W, int wind, int overlap, int* n_rows, int* n_cols);

/* ------------------------------------------------------------------------------------------
 * First pass, fully fused.  Replaces extended_search_area_piv (PB:459-520) =
 * moving_window_array (PB:220-247) + mean normalisation (PB:513-514) + correalte_fft
 * (PB:249-257) + `corr - amin` (PB:518) + correlation_to_displacement (PB:360-422) +
 * peak2peak_secondpeak (PB:346-358).
 *   frames_a/b : uint8, n_pairs frames of H rows, `pitch` bytes per row, consecutive pairs
 *                `pair_stride` bytes apart (ignored when n_pairs == 1)
 *   u, v       : float64 [n_pairs * n_rows * n_cols] displacement in px (x = columns, y = rows)
 *   mask       : uint8, same length, 1 = peak ratio < val_ratio; may be NULL iff !validate
 *   ratio      : optional float32 peak-to-second-peak ratio (NULL to skip; additive output)
 */
int pivb200_pass_first(const uint8_t* frames_a, const uint8_t* frames_b, int n_pairs,
                       long long pair_stride, int H, int W, int pitch, int wind, int overlap,
                       int validate, double val_ratio, double* u, double* v, uint8_t* mask,
                       float* ratio, void* stream);

/* ------------------------------------------------------------------------------------------
 * Later passes, fully fused.  Replaces the device part of piv_iteration_CWS.__call__
 * (PB:714-738) / piv_iteration_DWS.__call__ (PB:782-810): window shift by -/+ the predictor
 * half (biliniar_interpolation_CWS PB:147-194 / interpolation_DWS PB:197-216), correlation,
 * peak fit, validation and the predictor replacement logic.
 *   mode      : PIVB200_MODE_CWS -> shift_x/y are float32 [N] (= float32(u0/2), PB:714-717)
 *               PIVB200_MODE_DWS -> shift_x/y are int32   [N] (= rint(u0/2),     PB:784-790)
 *   base_u/v  : float64 [N], added to the measured correction (2*u2, PB:728-729 / 800-801); NULL = 0
 *   pred_u/v  : float64 [N], the (mask-zeroed) predictor used for the replacement rule
 *               (PB:731-738); NULL = no replacement (raw correction is returned)
 */
int pivb200_pass_next(const uint8_t* frames_a, const uint8_t* frames_b, int n_pairs,
                      long long pair_stride, int H, int W, int pitch, int wind, int overlap,
                      int mode, const void* shift_x, const void* shift_y, const double* base_u,
                      const double* base_v, const double* pred_u, const double* pred_v,
                      int validate, double val_ratio, double* u, double* v, uint8_t* mask,
                      float* ratio, void* stream);

/* ------------------------------------------------------------------------------------------
 * Predictor resampling between passes, on the device.  Replaces the three host-side
 * scipy RectBivariateSpline evaluations + mask thresholding + shift preparation of
 * PB:700-717 (CWS) / PB:769-790 (DWS).  The bicubic interpolating spline is a fixed linear
 * operator per geometry: S = Ay * U * Ax^T (Ay: [n1, n0], Ax: [m1, m0], float64, row-major,
 * built once on the host -- torchpiv_b200/geometry.py).
 *   u_prev, v_prev : float64 [n_pairs, n0, m0]; mask_prev: uint8 (NULL = no validation mask)
 *   tmp            : float64 scratch, 3 * n_pairs * n0 * m1 elements
 *   outputs (all [n_pairs, n1, m1]): shift_x/y (float32 for CWS, int32 for DWS), base_u/v,
 *   pred_u/v as consumed by pivb200_pass_next.
 */
int pivb200_predictor(const double* u_prev, const double* v_prev, const uint8_t* mask_prev,
                      int n_pairs, int n0, int m0, int n1, int m1, const double* Ay,
                      const double* Ax, int mode, double* tmp, void* shift_x, void* shift_y,
                      double* base_u, double* base_v, double* pred_u, double* pred_v,
                      void* stream);

/* ------------------------------------------------------------------------------------------
 * Function-level entry points (same maths as the fused passes, exposed so that every
 * backend function of the reference has a counterpart).
 */

/* correalte_fft (PB:249-257) on materialised windows [n, wind, wind]; dtype 0 = float32,
 * 1 = uint8 (promoted like torch.fft does).  corr: float32 [n, wind, wind], fft-shifted. */
int pivb200_correlate(const void* windows_a, const void* windows_b, int dtype, long long n,
                      int wind, float* corr, void* stream);

/* correlation_to_displacement (PB:360-422) on arbitrary maps [n, d, k] (any d, k >= 2);
 * dtype 0 = float32, 1 = float64.  Like the reference it MODIFIES corr (adds 1e-7, zeroes the
 * (2*val_window+1)^2 flat-index patch around the peak when validate != 0). */
int pivb200_corr_to_disp(void* corr, int dtype, long long n, int d, int k, int validate,
                         double val_ratio, int val_window, double* u, double* v, uint8_t* mask,
                         void* stream);

/* The shifted interrogation windows as the fused pass sees them (TMA tile + in-tile taps),
 * float32 [N, wind, wind] each.  mode/shift as in pivb200_pass_next; shift_x = NULL with
 * MODE_DWS gives the unshifted windows = moving_window_array (PB:220-247).
 * Unshifted and MODE_DWS windows are BIT-EXACT with the reference.  MODE_CWS windows are NOT: the fused
 * loader evaluates the reference's four-term bilinear sum (PB:187-192) in separable form (vertical tap,
 * then horizontal tap), which differs from it by FP32 rounding -- at most 2^-14 grey levels (observed
 * <= 4e-5 of 0..255), far inside the 1e-3 px tolerance of the displacements.  pivb200_bilinear_cws below
 * keeps the reference's exact evaluation order and IS bit-exact. */
int pivb200_windows(const uint8_t* frames_a, const uint8_t* frames_b, int n_pairs,
                    long long pair_stride, int H, int W, int pitch, int wind, int overlap,
                    int mode, const void* shift_x, const void* shift_y, float* win_a,
                    float* win_b, void* stream);

/* biliniar_interpolation_CWS (PB:147-194) with the reference's own argument layout: frame uint8
 * [H, W] (dense), grid int64 [n_elem] of flat pixel indices, vel_x/vel_y float32 per WINDOW
 * (elem_per_window = w*w).  out float32 [n_elem].  Bit-exact with the reference. */
int pivb200_bilinear_cws(const uint8_t* frame, int H, int W, const int64_t* grid, long long n_elem,
                         int elem_per_window, const float* vel_x, const float* vel_y, float* out,
                         void* stream);

/* interpolation_DWS (PB:197-216): out uint8 [n_elem] = frame.flat[clamp(grid + vy*W + vx)]. */
int pivb200_shift_dws(const uint8_t* frame, int H, int W, const int64_t* grid, long long n_elem,
                      int elem_per_window, const int64_t* vel_x, const int64_t* vel_y,
                      uint8_t* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * On-device field post-processing (additive; the reference has no device counterpart -- it fills
 * holes on the host with SciPy Delaunay interpolation, PB:266-344 + 884-892, and keeps its running
 * statistics in workers.py:79-119).  Fields are float64 [n_pairs][n_rows][n_cols].
 *
 * pivb200_nmt: normalised median test on the 3x3 neighbourhood (neighbours flagged in `mask` are
 *   ignored; `mask` may be NULL).  outlier[g] = mask[g] | (|u - med(u_nb)| / (med|u_nb - med| + eps)
 *   > threshold, same for v); vectors with fewer than two usable neighbours are not tested.
 *   `outlier` must not alias `mask`.
 * pivb200_replace: vectors flagged in `invalid` become the median of their usable 3x3 neighbours;
 *   Jacobi sweeps (at most max_sweeps, rounded up to even) let holes fill from the rim inwards;
 *   `invalid` is cleared where a value was produced; vectors never reached become 0 and stay
 *   flagged.  In place; `workspace` is pivb200_replace_workspace_bytes(...) bytes, 8-byte aligned.
 * pivb200_stats_accumulate: running moments[5][n_rows][n_cols] of the fields seen so far -- mean u,
 *   mean v, then the sums of (u-mean_u)^2, (v-mean_v)^2, (u-mean_u)(v-mean_v) -- are merged with this
 *   batch (pairwise update, free of the cancellation of raw power sums).  `n_before` = number of fields
 *   already merged (0 with a zeroed buffer).  The streaming form of workers.py:79-95's stacked arrays. */
int pivb200_nmt(const double* u, const double* v, const uint8_t* mask, int n_pairs, int n_rows,
                int n_cols, double threshold, double eps, uint8_t* outlier, void* stream);
long long pivb200_replace_workspace_bytes(int n_pairs, int n_rows, int n_cols);
int pivb200_replace(double* u, double* v, uint8_t* invalid, int n_pairs, int n_rows, int n_cols,
                    int max_sweeps, void* workspace, void* stream);
int pivb200_stats_accumulate(const double* u, const double* v, int n_pairs, int n_rows, int n_cols,
                             long long n_before, double* moments, void* stream);

/* ------------------------------------------------------------------------------------------
 * Measurement helpers (bench.py): sustained FP32 FFMA throughput of the current device in
 * TFLOP/s (2 flops per FFMA), timed with CUDA events over `iters` launches.  HOST pointer. */
int pivb200_measure_fp32_peak(int iters, double* tflops_host, void* stream);

/* Number of kernel launches issued by this library in the calling process (all entry points). */
long long pivb200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* PIVB200_H_ */
