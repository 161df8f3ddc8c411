"""Per-pair host post-processing of the final vector field (NumPy / SciPy, float64).

Same behaviour as the tail of the reference's ``OfflinePIV.__call__`` (PIVbackend.py:884-900):
invalid vectors -> NaN, linear fill along the four borders (``interpolate_boarders``,
PIVbackend.py:328-344), Delaunay-linear fill of the remaining holes from the ring of valid
neighbours (``fillMissingValues`` / ``getPixelsForInterp``, PIVbackend.py:266-308), vertical
flip, sign of v, physical units.  The field is a few thousand vectors; this is not on the
GPU hot path."""
from __future__ import annotations

import numpy as np
from scipy.interpolate import LinearNDInterpolator

__all__ = ["fill_borders", "fill_holes", "finalize_field", "HoleFillPool"]


def fill_borders(field: np.ndarray) -> np.ndarray:
    """In-place 1-D linear interpolation of NaNs on the first/last row and column.  An edge
    that is entirely NaN is left alone; NaNs beyond the outermost valid sample take its value."""
    if not np.isnan(field).any():
        return field
    edges = (field[0, :], field[-1, :], field[:, 0], field[:, -1])
    for edge in edges:                      # views: assignments write through to `field`
        bad = np.isnan(edge)
        if bad.all() or not bad.any():
            continue
        where = np.flatnonzero
        edge[bad] = np.interp(where(bad), where(~bad), edge[~bad])
    return field


def _neighbour_ring(invalid: np.ndarray) -> np.ndarray:
    """Valid cells that touch an invalid one through an edge (3x3 plus-shaped dilation of the
    invalid set with a zero border, minus the set itself)."""
    grown = invalid.copy()
    grown[1:] |= invalid[:-1]
    grown[:-1] |= invalid[1:]
    grown[:, 1:] |= invalid[:, :-1]
    grown[:, :-1] |= invalid[:, 1:]
    return grown & ~invalid


def _delaunay_fill(support: np.ndarray, values: np.ndarray, targets: np.ndarray):
    """Piecewise-linear (Qhull Delaunay) interpolation of `values` [n, k] given at `support` [n, 2] onto `targets`
    [m, 2]; None when the triangulation fails.  Module-level so that a worker process can run it."""
    try:
        return LinearNDInterpolator(support, values)(targets)
    except Exception:
        return None


class HoleFillPool:
    """Worker PROCESSES for the Delaunay hole filling of the reference-exact mode.

    Qhull (SciPy) holds the GIL for the 2-6 ms a 4 MP field takes, so threads cannot spread it over the host
    cores; processes can.  Only the ring points, their values and the hole coordinates travel (tens of KB per
    pair).  The workers are forked once, when the pool is built (before the decode threads start), and only ever run
    NumPy / SciPy code -- the same arrangement as torch's DataLoader workers next to a CUDA parent; "spawn" would
    re-import the user's main script in every worker."""

    def __init__(self, workers: int):
        import multiprocessing as mp
        from concurrent.futures import ProcessPoolExecutor
        self.workers = int(workers)
        self._pool = ProcessPoolExecutor(max_workers=self.workers, mp_context=mp.get_context("fork"))
        # start every worker now (the first submit would otherwise pay the interpreter start-up one by one)
        list(self._pool.map(_warm, range(self.workers)))

    def submit(self, support, values, targets):
        return self._pool.submit(_delaunay_fill, support, values, targets)

    def shutdown(self):
        self._pool.shutdown(wait=False, cancel_futures=True)


def _warm(i):
    return i


def fill_holes(field: np.ndarray, *more: np.ndarray, pool: "HoleFillPool | None" = None):
    """Fill NaNs by piecewise-linear (Delaunay) interpolation over the ring of valid neighbours.
    Returns None -- the caller then skips the pair, as the reference does -- when the ring is
    empty (no invalid vector at all), when the triangulation fails, or when the ring covers a
    quarter of the field or more ("too many false vectors").

    ``more``: further fields with the SAME NaN pattern (u and v of one pair).  They share one
    triangulation (the reference triangulates the identical point set twice, PB:885-890; Qhull is
    deterministic, so the values are the same and the dominant cost is halved); the return value
    is then a tuple of the filled fields."""
    fields = (field,) + more
    invalid = np.isnan(field)
    for other in more:
        if not np.array_equal(np.isnan(other), invalid):      # different holes: fall back to one by one
            out = tuple(fill_holes(f) for f in fields)          # (rare; synchronous)
            return None if any(o is None for o in out) else out
    ring = _neighbour_ring(invalid)
    support = np.argwhere(ring)
    if not support.size < ring.size / 2:
        print("Warning! to many false vectors")
        return None
    values = np.stack([f[ring] for f in fields], axis=1)
    targets = np.argwhere(invalid)
    if pool is not None:
        # asynchronous: the caller resolves the future with finish_fill()
        return _PendingFill(fields, invalid, pool.submit(support, values, targets), bool(more))
    filled = _delaunay_fill(support, values, targets)
    if filled is None:
        return None
    for k, f in enumerate(fields):
        f[invalid] = filled[:, k]
    return fields if more else field


class _PendingFill:
    """A hole filling running in a worker process."""

    def __init__(self, fields, invalid, future, many):
        self.fields, self.invalid, self.future, self.many = fields, invalid, future, many

    def result(self):
        filled = self.future.result()
        if filled is None:
            return None
        for k, f in enumerate(self.fields):
            f[self.invalid] = filled[:, k]
        return self.fields if self.many else self.fields[0]


def finalize_field(u, v, x, y, invalid, scale: float = 1.0, dt: float = 1.0, pool=None):
    """u, v in px (float64, modified in place) -> (x, y, u, v) in mm and m/s, or None to skip.
    With ``pool`` (HoleFillPool) the Delaunay fill runs in a worker process and a zero-argument callable that
    completes the field is returned instead."""
    if invalid is not None:
        u[invalid] = np.nan
        v[invalid] = np.nan
        filled = fill_holes(fill_borders(u), fill_borders(v), pool=pool)
        if filled is None:
            return None
        if isinstance(filled, _PendingFill):
            def complete(pending=filled):
                done = pending.result()
                return None if done is None else _to_units(done[0], done[1], x, y, scale, dt)
            return complete
        u, v = filled
    return _to_units(u, v, x, y, scale, dt)


def _to_units(u, v, x, y, scale, dt):
    u = np.flip(u, axis=0) * scale / dt * 1000
    v = -np.flip(v, axis=0) * scale / dt * 1000
    return x * scale, y * scale, u, v
