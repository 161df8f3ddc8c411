"""Host-side geometry of a PIV pass: field shape, window-centre coordinates and the predictor
resampling operator.  Pure NumPy; tiny; computed once per geometry.

Mirrors ``get_field_shape`` / ``get_coordinates`` of the reference (PIVbackend.py:425-456,
522-597) and replaces its per-pair ``scipy.interpolate.RectBivariateSpline`` calls
(PIVbackend.py:700-713, 769-780) by a precomputed linear operator: the interpolating bicubic
spline FITPACK builds for ``s=0`` is the tensor product of two 1-D interpolating cubic splines
with the not-a-knot knot vector, so ``spline(U)(y_new, x_new) == Ay @ U @ Ax.T`` where ``Ay``,
``Ax`` depend only on the old and new grid coordinates.  Evaluation points outside the old
grid are clamped to its bounding box, as FITPACK's ``bispev`` does.
"""
from __future__ import annotations

import numpy as np

__all__ = ["get_field_shape", "get_coordinates", "spline_operator"]

FUSED_WINDOWS = (16, 32, 64)          # fused in-register FFT kernels
MAX_WINDOW = 256                      # general-size path (mixed-radix FFT in shared memory): any size up to this
SUPPORTED_WINDOWS = FUSED_WINDOWS     # kept for callers that ask which sizes take the fast path


def window_supported(wind: int) -> bool:
    """Sizes the library can process: 4..256 px (16/32/64 px take the fused kernels; odd sizes give the
    reference's [w, w-1] correlation maps)."""
    return int(wind) == wind and 4 <= wind <= MAX_WINDOW


def get_field_shape(image_size, search_area_size, overlap):
    """Number of interrogation windows per axis: ``(size - w) // (w - ovl) + 1``."""
    size = np.asarray(image_size)
    return (size - search_area_size) // (search_area_size - overlap) + 1


def get_coordinates(image_size, search_area_size, overlap):
    """Window-centre coordinates ``x, y`` (float64 ``[n_rows, n_cols]``).  Windows themselves
    start at pixel 0; only the reported coordinates are centred, by a whole number of pixels."""
    n_rows, n_cols = (int(v) for v in get_field_shape(image_size, search_area_size, overlap)[-2:])
    step = search_area_size - overlap
    span_x = (n_cols - 1) * step + (search_area_size - 1)
    span_y = (n_rows - 1) * step + (search_area_size - 1)
    x = np.arange(n_cols, dtype=np.float64) * step + search_area_size / 2.0
    y = np.arange(n_rows, dtype=np.float64) * step + search_area_size / 2.0
    x += (int(image_size[-1]) - 1 - span_x) // 2
    y += (int(image_size[-2]) - 1 - span_y) // 2
    return np.meshgrid(x, y)


def _bspline_basis(t: np.ndarray, k: int, x: np.ndarray, n: int) -> np.ndarray:
    """Dense matrix B[i, j] = B_j(x_i) of the n B-splines of degree k on knot vector t
    (Cox - de Boor recursion on the k+1 splines that are non-zero at each point)."""
    x = np.asarray(x, dtype=np.float64)
    # interval index l with t[l] <= x < t[l+1], the right end belongs to the last interval
    l = np.searchsorted(t, x, side="right") - 1
    l = np.clip(l, k, n - 1)
    vals = np.zeros((x.size, k + 1))
    vals[:, 0] = 1.0
    for d in range(1, k + 1):
        saved = np.zeros(x.size)
        for r in range(d):
            tr = t[l + r + 1]
            tl = t[l + r + 1 - d]
            denom = tr - tl
            term = np.where(denom != 0, vals[:, r] / np.where(denom != 0, denom, 1.0), 0.0)
            vals[:, r] = saved + (tr - x) * term
            saved = (x - tl) * term
        vals[:, d] = saved
    out = np.zeros((x.size, n))
    rows = np.arange(x.size)
    for r in range(k + 1):
        out[rows, l - k + r] = vals[:, r]
    return out


def spline_operator(old: np.ndarray, new: np.ndarray) -> np.ndarray:
    """Matrix ``A [len(new), len(old)]`` with ``A @ f == s(clip(new))`` where ``s`` is the cubic
    interpolating (not-a-knot) spline through ``(old, f)`` -- one axis of FITPACK's ``s=0``
    ``RectBivariateSpline``.  ``old`` must be strictly increasing with at least 4 points."""
    old = np.asarray(old, dtype=np.float64)
    new = np.asarray(new, dtype=np.float64)
    n, k = old.size, 3
    if n < k + 1:
        raise ValueError("the predictor field needs at least 4 points per axis for the bicubic spline")
    if np.any(np.diff(old) <= 0):
        raise ValueError("old grid coordinates must be strictly increasing")
    t = np.concatenate([[old[0]] * (k + 1), old[2:n - 2], [old[-1]] * (k + 1)])
    colloc = _bspline_basis(t, k, old, n)
    evalm = _bspline_basis(t, k, np.clip(new, old[0], old[-1]), n)
    # A = evalm @ inv(colloc); solve the transposed system for accuracy
    return np.linalg.solve(colloc.T, evalm.T).T.copy()
