"""CPU oracle for the TorchPIV cross-correlation hot path  --  TEST INFRASTRUCTURE ONLY.

This module is a NumPy/SciPy restatement of the algorithm that the reference
(NikNazarov/TorchPIV, ``src/torchPIV/PIVbackend.py``, abbreviated ``PB`` below)
runs for ``OfflinePIV``.  It is the *checker* for the CUDA path and the "port"
CPU baseline of ``bench.py``; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product package (``torchpiv_b200``) never imports anything from ``oracle/``.

Parity status: PINNED.  Every function here is compared against the unmodified
reference (imported from ``/root/reference`` by ``tests/golden/make_golden.py``)
and the resulting input/output vectors are committed under ``tests/golden/``;
``tests/test_oracle_golden.py`` replays them without the reference being present.

Third-party arithmetic the reference delegates to (and the oracle therefore also
delegates to, same libraries, versions as installed in this image):
  * FFT: ``torch.fft.rfft2 / irfft2`` (PB:255-256).  The oracle uses
    ``scipy.fft`` (pocketfft, scipy 1.18.1) -- same transform, same dtype rules
    (float32 -> complex64, float64 -> complex128); results agree to rounding.
  * predictor resampling: ``scipy.interpolate.RectBivariateSpline`` (FITPACK,
    PB:700-713, 769-780).
  * hole filling: ``scipy.interpolate.LinearNDInterpolator`` (Qhull) and a 3x3
    cross dilation (``cv2.dilate`` in the reference, PB:266-308).

All index conventions are the reference's: flat index ``m = r*k + c`` on the
fft-shifted map, neighbours / 7x7 exclusion patch addressed in *flat* order
(wrapping across row ends), ``True`` in a mask means INVALID.
"""
from __future__ import annotations

import numpy as np
from scipy import fft as _sfft
from scipy import interpolate as _interp

EPS = 1e-7  # PB:380


# --------------------------------------------------------------------------------------
# geometry (PB:425-456, 522-597)
# --------------------------------------------------------------------------------------
def get_field_shape(image_size, search_area_size, overlap):
    """PB:453-456: ``(size - w) // (w - ovl) + 1`` per axis."""
    size = np.asarray(image_size)
    return (size - search_area_size) // (search_area_size - overlap) + 1


def get_coordinates(image_size, search_area_size, overlap):
    """PB:562-597: window-centre coordinates, centred by an integer offset."""
    n_rows, n_cols = get_field_shape(image_size, search_area_size, overlap)[-2:]
    step = search_area_size - overlap
    x = np.arange(n_cols, dtype=np.int32) * step + search_area_size / 2.0
    y = np.arange(n_rows, dtype=np.int32) * step + search_area_size / 2.0
    x += (image_size[-1] - 1 - ((n_cols - 1) * step + (search_area_size - 1))) // 2
    y += (image_size[-2] - 1 - ((n_rows - 1) * step + (search_area_size - 1))) // 2
    return np.meshgrid(x, y)


def moving_window_array(array: np.ndarray, window_size: int, overlap: int) -> np.ndarray:
    """PB:232-247: [n_rows*n_cols, w, w] windows, row-major, origin at pixel 0."""
    h, w = array.shape[-2:]
    step = window_size - overlap
    n_r = int((h - window_size) / step) + 1
    n_c = int((w - window_size) / step) + 1
    s0, s1 = array.strides[-2:]
    view = np.lib.stride_tricks.as_strided(
        array, shape=(n_r, n_c, window_size, window_size),
        strides=(s0 * step, s1 * step, s0, s1), writeable=False)
    return view.reshape(-1, window_size, window_size)


def window_index_grid(frame_shape, window_size, overlap) -> np.ndarray:
    """PB:684-687: flat pixel indices of every window, int64 [N, w, w]."""
    h, w = frame_shape[-2:]
    flat = np.arange(h * w, dtype=np.int64).reshape(h, w)
    return moving_window_array(flat, window_size, overlap)


# --------------------------------------------------------------------------------------
# window shifting (PB:147-216)
# --------------------------------------------------------------------------------------
def bilinear_interpolation_cws(array: np.ndarray, grid: np.ndarray,
                               vel_x: np.ndarray, vel_y: np.ndarray) -> np.ndarray:
    """PB:147-194.  ``array`` uint8 [H,W]; ``grid`` int64 [N,w,w]; ``vel_*`` float32 [N,1,1].

    torch promotes ``int64 + float32`` to float32, so the shifted coordinate is rounded
    at the magnitude of the absolute pixel coordinate; NumPy would promote to float64,
    hence the explicit casts.  Taps are addressed by *flat* index clamped to
    ``[0, H*W-1]``; where either coordinate is an exact integer the value is the tap at
    ``(floor y, floor x)`` (PB:170, 193).
    """
    f32 = np.float32
    wf = array.shape[-1]
    n = array.size
    gy = (grid // wf).astype(f32)
    gx = (grid % wf).astype(f32)
    new_y = gy + vel_y.astype(f32)
    new_x = gx + vel_x.astype(f32)
    up_x = np.ceil(new_x).astype(np.int64)
    up_y = np.ceil(new_y).astype(np.int64)
    dn_x = np.floor(new_x).astype(np.int64)
    dn_y = np.floor(new_y).astype(np.int64)
    exact = (up_x - dn_x) * (up_y - dn_y) == 0
    flat = array.reshape(-1)

    def tap(yy, xx):
        return flat[np.clip(yy * wf + xx, 0, n - 1)].astype(f32)

    q11 = tap(dn_y, dn_x)
    q12 = tap(up_y, dn_x)
    q21 = tap(dn_y, up_x)
    q22 = tap(up_y, up_x)
    wx1 = up_x.astype(f32) - new_x
    wx0 = new_x - dn_x.astype(f32)
    wy1 = up_y.astype(f32) - new_y
    wy0 = new_y - dn_y.astype(f32)
    out = q11 * wx1 * wy1 + q21 * wx0 * wy1 + q12 * wx1 * wy0 + q22 * wx0 * wy0
    out[exact] = q11[exact]
    return out.astype(f32, copy=False)


def interpolation_dws(array: np.ndarray, grid: np.ndarray,
                      vel_x: np.ndarray, vel_y: np.ndarray) -> np.ndarray:
    """PB:197-216: integer window shift by flat-index arithmetic, clamped, dtype kept."""
    wf = array.shape[-1]
    idx = grid + vel_y.astype(np.int64) * wf + vel_x.astype(np.int64)
    np.clip(idx, 0, array.size - 1, out=idx)
    return array.reshape(-1)[idx]


# --------------------------------------------------------------------------------------
# correlation (PB:249-257)
# --------------------------------------------------------------------------------------
def correlate_fft(images_a: np.ndarray, images_b: np.ndarray, workers: int = -1) -> np.ndarray:
    """PB:255-256: fftshift(irfft2(conj(rfft2(a)) * rfft2(b))).  uint8 is promoted to
    float32 (what torch.fft does), float32 stays float32, float64 stays float64."""
    if images_a.dtype == np.uint8:
        images_a = images_a.astype(np.float32)
    if images_b.dtype == np.uint8:
        images_b = images_b.astype(np.float32)
    fa = _sfft.rfft2(images_a, workers=workers)
    fb = _sfft.rfft2(images_b, workers=workers)
    np.conjugate(fa, out=fa)
    fa *= fb
    # no `s=`: like torch.fft.irfft2 in the reference, the last axis comes back with 2 (m - 1) samples -- w for an
    # even window, w - 1 for an odd one (the reference's map for odd windows is [w, w - 1]; PB:255)
    corr = _sfft.irfft2(fa, workers=workers)
    return _sfft.fftshift(corr, axes=(-2, -1))


# --------------------------------------------------------------------------------------
# peak finding, sub-pixel fit, validation (PB:346-422)
# --------------------------------------------------------------------------------------
def peak2peak_secondpeak(flat: np.ndarray, imax: np.ndarray, k: int, wind: int = 2) -> np.ndarray:
    """PB:346-358: zero the (2*wind+1)^2 flat-index patch around ``imax`` IN PLACE
    (each index clamped to [0, d*k-1]), return argmax of what is left."""
    c, n2 = flat.shape
    rows = np.arange(c)
    for i in range(-wind, wind + 1):
        for j in range(-wind, wind + 1):
            ids = np.clip(imax + i + k * j, 0, n2 - 1)
            flat[rows, ids] = 0.0
    return flat.argmax(axis=1)


def correlation_to_displacement(corr: np.ndarray, n_rows: int, n_cols: int,
                                validate: bool = True, val_ratio: float = 1.2,
                                validation_window: int = 3):
    """PB:360-422.  ``corr`` [c, d, k] is modified in place (``+= eps``; patch zeroing).
    Returns ``u, v`` float64 [n_rows, n_cols] and the boolean INVALID mask (or None)."""
    c, d, k = corr.shape
    n2 = d * k
    corr += corr.dtype.type(EPS)                      # PB:381, in corr's dtype
    flat = corr.reshape(c, n2)
    # PB:382: .type(float64) copies float32 data but aliases float64 data
    cor = flat if flat.dtype == np.float64 else flat.astype(np.float64)
    m = flat.argmax(axis=1)                           # first maximum in flat order
    left, right, top, bot = m + 1, m - 1, m + k, m - k
    left = np.where(left >= n2 - 1, m, left)          # PB:389-392: only the array ends are guarded
    right = np.where(right <= 0, m, right)
    top = np.where(top >= n2 - 1, m, top)
    bot = np.where(bot <= 0, m, bot)
    rows = np.arange(c)
    cm, cl, cr = cor[rows, m], cor[rows, left], cor[rows, right]
    ct, cb = cor[rows, top], cor[rows, bot]
    with np.errstate(all="ignore"):
        lm, ll, lr, lt, lb = np.log(cm), np.log(cl), np.log(cr), np.log(ct), np.log(cb)
        nom1 = lr - ll
        den1 = 2 * (ll + lr) - 4 * lm
        nom2 = lb - lt
        den2 = 2 * (lb + lt) - 4 * lm
        v = (m // d) + nom2 / den2
        u = (m % k) + nom1 / den1
    mask = None
    if validate:
        m2 = peak2peak_secondpeak(flat, m, k, validation_window)
        with np.errstate(all="ignore"):
            mask = (cm / cor[rows, m2]) < val_ratio
        mask = mask.reshape(n_rows, n_cols)
    v = v - int(d / 2)
    u = u - int(k / 2)
    u = np.nan_to_num(u).reshape(n_rows, n_cols)
    v = np.nan_to_num(v).reshape(n_rows, n_cols)
    return u, v, mask


# --------------------------------------------------------------------------------------
# passes (PB:459-520, 677-812)
# --------------------------------------------------------------------------------------
def extended_search_area_piv(frame_a: np.ndarray, frame_b: np.ndarray, window_size: int = 32,
                             overlap: int = 0, validate: bool = False,
                             validation_ratio: float = 1.2, workers: int = -1,
                             compute_dtype=np.float64, stash: dict = None):
    """PB:459-520 (first pass, float64 after the mean normalisation).  ``stash`` (test hook, not
    reference behaviour) receives a copy of the min-subtracted correlation maps under "corr".  ``compute_dtype=float32``
    is NOT the reference: it re-evaluates the same maths in single precision so that tests can
    tell ill-conditioned vectors (those that move when the arithmetic changes) from real errors."""
    if overlap >= window_size:
        raise ValueError("Overlap has to be smaller than the window_size")
    if window_size > frame_a.shape[-2] or window_size > frame_a.shape[-1]:
        raise ValueError("window size cannot be larger than the image")
    n_rows, n_cols = get_field_shape(frame_a.shape, window_size, overlap)
    x, y = get_coordinates(frame_a.shape, window_size, overlap)
    aa = moving_window_array(frame_a, window_size, overlap)
    bb = moving_window_array(frame_b, window_size, overlap)
    with np.errstate(all="ignore"):
        aa = aa / aa.mean(axis=(-2, -1), dtype=np.float64, keepdims=True)
        bb = bb / bb.mean(axis=(-2, -1), dtype=np.float64, keepdims=True)
    if compute_dtype != np.float64:
        aa, bb = aa.astype(compute_dtype), bb.astype(compute_dtype)
    corr = correlate_fft(aa, bb, workers=workers)
    corr = corr - corr.min(axis=(-2, -1), keepdims=True)
    if stash is not None:
        stash["corr"] = corr.copy()
    u, v, mask = correlation_to_displacement(corr, n_rows, n_cols, validate, validation_ratio)
    return u, v, x, y, mask


def resample_predictor(x0, y0, field, slice_y, slice_x):
    """PB:700-704: bicubic FITPACK spline through the old grid, evaluated on the new one."""
    spl = _interp.RectBivariateSpline(y0[:, 0], x0[0, :], field)
    return spl(slice_y, slice_x)


class PivIteration:
    """Common skeleton of PB:677-740 (CWS) and PB:744-812 (DWS)."""

    mode = "CWS"

    def __init__(self, frame_shape, wind_size, overlap, workers: int = -1, compute_dtype=None):
        # compute_dtype=float64 is NOT the reference (which correlates in float32): conditioning probe
        self.compute_dtype = compute_dtype
        self.frame_shape = tuple(frame_shape)
        self.wind_size, self.overlap = wind_size, overlap
        self.n_rows, self.n_cols = get_field_shape(frame_shape, wind_size, overlap)
        self.x, self.y = get_coordinates(frame_shape, wind_size, overlap)
        self.slice_x, self.slice_y = self.x[0, :], self.y[:, 0]
        self.idx = window_index_grid(frame_shape, wind_size, overlap)
        self.workers = workers

    def shifted_windows(self, frame_a, frame_b, u0, v0):
        raise NotImplementedError

    def __call__(self, frame_a, frame_b, x0, y0, u0, v0, validation_mask):
        u0 = resample_predictor(x0, y0, u0, self.slice_y, self.slice_x)
        v0 = resample_predictor(x0, y0, v0, self.slice_y, self.slice_x)
        val_pred = None
        if validation_mask is not None:
            val_pred = resample_predictor(x0, y0, validation_mask, self.slice_y, self.slice_x) >= .5
        aa, bb, u_base, v_base, u0, v0 = self.shifted_windows(frame_a, frame_b, u0, v0, val_pred)
        if self.compute_dtype is not None:
            aa, bb = aa.astype(self.compute_dtype), bb.astype(self.compute_dtype)
        corr = correlate_fft(aa, bb, workers=self.workers)
        corr = corr - corr.min(axis=(-2, -1), keepdims=True)
        self.last_corr = corr.copy()            # test hook (conditioning analysis), not reference behaviour
        du, dv, val = correlation_to_displacement(corr, self.n_rows, self.n_cols,
                                                  validation_mask is not None)
        u = u_base + du
        v = v_base + dv
        mask_u = (du > u0) * (np.rint(u0) > 0)      # PB:731-732 / 803-804
        mask_v = (dv > v0) * (np.rint(v0) > 0)
        if val is not None:
            mask_u[val] = True
            mask_v[val] = True
        v[mask_v] = v0[mask_v]
        u[mask_u] = u0[mask_u]
        return u, v, self.x, self.y, val


class PivIterationCWS(PivIteration):
    mode = "CWS"

    def shifted_windows(self, frame_a, frame_b, u0, v0, val_pred):
        u2, v2 = u0 / 2, v0 / 2                      # PB:705-706: BEFORE the invalid zeroing
        if val_pred is not None:
            u0[val_pred] = 0.0
            v0[val_pred] = 0.0
        u2t = u2.astype(np.float32).reshape(-1, 1, 1)
        v2t = v2.astype(np.float32).reshape(-1, 1, 1)
        aa = bilinear_interpolation_cws(frame_a, self.idx, -u2t, -v2t)
        bb = bilinear_interpolation_cws(frame_b, self.idx, u2t, v2t)
        return aa, bb, 2 * u2, 2 * v2, u0, v0


class PivIterationDWS(PivIteration):
    mode = "DWS"

    def shifted_windows(self, frame_a, frame_b, u0, v0, val_pred):
        if val_pred is not None:                     # PB:775-780: zeroing FIRST
            u0[val_pred] = 0.0
            v0[val_pred] = 0.0
        u2, v2 = np.rint(u0 / 2), np.rint(v0 / 2)    # round-half-even
        u2t = u2.astype(np.int64).reshape(-1, 1, 1)
        v2t = v2.astype(np.int64).reshape(-1, 1, 1)
        aa = interpolation_dws(frame_a, self.idx, -u2t, -v2t)
        bb = interpolation_dws(frame_b, self.idx, u2t, v2t)
        return aa, bb, 2 * np.rint(u2), 2 * np.rint(v2), u0, v0


ITER_MODES = {"CWS": PivIterationCWS, "DWS": PivIterationDWS}


# --------------------------------------------------------------------------------------
# post-processing (PB:266-344, 884-900)
# --------------------------------------------------------------------------------------
def _dilate_cross3(mask: np.ndarray) -> np.ndarray:
    """cv2.dilate with the 3x3 MORPH_ELLIPSE element (= a plus sign), zero border (PB:275-279)."""
    out = mask.copy()
    out[1:, :] |= mask[:-1, :]
    out[:-1, :] |= mask[1:, :]
    out[:, 1:] |= mask[:, :-1]
    out[:, :-1] |= mask[:, 1:]
    return out


def fill_missing_values(target: np.ndarray):
    """PB:284-308.  Returns None when the reference would (empty ring, Qhull failure,
    too many invalid vectors)."""
    invalid = np.isnan(target)
    ring = _dilate_cross3(invalid) & ~invalid
    points = np.argwhere(ring)
    values = target[ring]
    if points.size < ring.size / 2:
        try:
            interp = _interp.LinearNDInterpolator(points, values)
            target[invalid] = interp(np.argwhere(invalid))
        except Exception:
            return None
    else:
        return None
    return target


def interpolate_borders(vec: np.ndarray) -> np.ndarray:
    """PB:328-344: 1-D linear fill of NaNs along the four edges (edge values extended)."""
    if not np.isnan(vec).any():
        return vec
    for sl in ((0, slice(None)), (-1, slice(None)), (slice(None), 0), (slice(None), -1)):
        line = vec[sl]
        nans = np.isnan(line)
        if not nans.all():
            pos = np.arange(line.size)
            line[nans] = np.interp(pos[nans], pos[~nans], line[~nans])
    return vec


def piv_passes(frame_a: np.ndarray, frame_b: np.ndarray, wind_size: int, overlap: int,
               multipass: int = 1, multipass_mode: str = "CWS", multipass_scale: float = 2.0,
               workers: int = -1, iter_objs=None):
    """PB:874-882: the per-pair pass loop (no post-processing).  Returns u, v, x, y, val
    of the last pass, plus the list of per-pass (u, v, val) for function-boundary parity."""
    u, v, x, y, val = extended_search_area_piv(frame_a, frame_b, wind_size, overlap,
                                                validate=True, workers=workers)
    history = [(u.copy(), v.copy(), None if val is None else val.copy())]
    w, o = wind_size, overlap
    for it in range(multipass - 1):
        w = int(w // multipass_scale)
        o = int(o // multipass_scale)
        obj = iter_objs[it] if iter_objs is not None else ITER_MODES[multipass_mode](
            frame_a.shape, w, o, workers=workers)
        u, v, x, y, val = obj(frame_a, frame_b, x, y, u, v, val)
        history.append((u.copy(), v.copy(), None if val is None else val.copy()))
    return u, v, x, y, val, history


def postprocess(u, v, x, y, val, scale: float = 1.0, dt: float = 1.0):
    """PB:884-900.  Returns (x, y, u, v) or None when the reference skips the pair."""
    if val is not None:
        u[val] = np.nan
        v[val] = np.nan
        u = interpolate_borders(u)
        v = interpolate_borders(v)
        u = fill_missing_values(u)
        v = fill_missing_values(v)
        if u is None or v is None:
            return None
    u = np.flip(u, axis=0)
    v = -np.flip(v, axis=0)
    u = u * scale / dt * 1000
    v = v * scale / dt * 1000
    return x * scale, y * scale, u, v


def offline_piv_pair(frame_a, frame_b, wind_size, overlap, multipass=1, multipass_mode="CWS",
                     dt=1, scale=1.0, multipass_scale=2.0, workers: int = -1, iter_objs=None):
    """One iteration of OfflinePIV.__call__ (PB:868-901) on already decoded frames."""
    u, v, x, y, val, _ = piv_passes(frame_a, frame_b, wind_size, overlap, multipass,
                                    multipass_mode, multipass_scale, workers, iter_objs)
    return postprocess(u, v, x, y, val, scale, dt)
