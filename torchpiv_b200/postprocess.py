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
    [m, 2]; None when the triangulation fails."""
    try:
        return LinearNDInterpolator(support, values)(targets)
    except Exception:
        return None


# Shared slots between the parent and its forked workers: {"U": [slots, B, n, m] f64, "V": ..., "M": uint8, "OK": [slots, B]}
_SLOTS = None


def _finalize_slot(slot, count, scale, dt):
    """Worker-process task: finalize_uv for the `count` pairs of a shared slot, IN PLACE (the finished u, v replace the
    raw ones; OK[slot, i] = 0 marks a pair the reference would skip)."""
    U, V, M, OK = _SLOTS["U"], _SLOTS["V"], _SLOTS["M"], _SLOTS["OK"]
    for i in range(count):
        out = finalize_uv(U[slot, i], V[slot, i], M[slot, i].view(np.bool_), scale, dt)
        if out is None:
            OK[slot, i] = 0
        else:
            OK[slot, i] = 1
            U[slot, i] = out[0]
            V[slot, i] = out[1]
    return count


def _warm(i):
    return i


class HoleFillPool:
    """Worker PROCESSES for the host post-processing of the reference-exact mode (one task per batch of pairs).

    Qhull (SciPy) holds the GIL for the 2-6 ms a 4 MP field takes, and the NumPy glue around it is GIL-bound too,
    so threads cannot spread it over the host cores; processes can.  The fields travel through anonymous shared
    memory mapped BEFORE the workers are forked (a batch is 8.5 MB; pickling it through the executor's pipes costs
    more than the triangulations), only slot numbers go through the task queue.  The workers are forked once, when
    the pool is built (before the decode threads start), and only ever run NumPy / SciPy code -- the same
    arrangement as torch's DataLoader workers next to a CUDA parent; "spawn" would re-import the user's main script
    in every worker."""

    def __init__(self, workers: int, batch_pairs: int, n_rows: int, n_cols: int):
        import mmap
        import multiprocessing as mp
        from concurrent.futures import ProcessPoolExecutor
        global _SLOTS
        self.workers = int(workers)
        self.n_slots = 2 * self.workers + 2
        B, n, m = int(batch_pairs), int(n_rows), int(n_cols)
        f_bytes, m_bytes = self.n_slots * B * n * m * 8, self.n_slots * B * n * m
        self._maps = [mmap.mmap(-1, f_bytes), mmap.mmap(-1, f_bytes), mmap.mmap(-1, m_bytes), mmap.mmap(-1, self.n_slots * B)]
        shape = (self.n_slots, B, n, m)
        self._views = {"U": np.frombuffer(self._maps[0], np.float64).reshape(shape),
                       "V": np.frombuffer(self._maps[1], np.float64).reshape(shape),
                       "M": np.frombuffer(self._maps[2], np.uint8).reshape(shape),
                       "OK": np.frombuffer(self._maps[3], np.uint8).reshape(self.n_slots, B)}
        _SLOTS = self._views                      # inherited by the forked workers
        self._free = list(range(self.n_slots))
        self._pool = ProcessPoolExecutor(max_workers=self.workers, mp_context=mp.get_context("fork"))
        # start every worker now (the first submit would otherwise pay the fork one by one)
        list(self._pool.map(_warm, range(self.workers)))

    @property
    def capacity(self) -> int:
        """Batches that can be in flight."""
        return self.n_slots

    def submit_batch(self, u, v, invalid, scale, dt):
        """u, v [k, n, m] float64, invalid [k, n, m] bool -> a handle for :meth:`collect`."""
        slot = self._free.pop()
        k = u.shape[0]
        self._views["U"][slot, :k] = u
        self._views["V"][slot, :k] = v
        self._views["M"][slot, :k] = invalid
        return self._pool.submit(_finalize_slot, slot, k, scale, dt), slot, k

    def collect(self, handle):
        """Wait for a batch; returns a list of (u, v) copies or None per pair and frees the slot."""
        fut, slot, k = handle
        fut.result()
        U, V, OK = self._views["U"], self._views["V"], self._views["OK"]
        out = [(U[slot, i].copy(), V[slot, i].copy()) if OK[slot, i] else None for i in range(k)]
        self._free.append(slot)
        return out

    def shutdown(self):
        self._pool.shutdown(wait=False, cancel_futures=True)


def fill_holes(field: np.ndarray, *more: np.ndarray):
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
    filled = _delaunay_fill(support, values, targets)
    if filled is None:
        return None
    for k, f in enumerate(fields):
        f[invalid] = filled[:, k]
    return fields if more else field


def finalize_uv(u, v, invalid, scale: float = 1.0, dt: float = 1.0):
    """u, v in px (float64, modified in place) -> (u, v) in m/s, rows flipped, v negated; None to skip the pair."""
    if invalid is not None:
        u[invalid] = np.nan
        v[invalid] = np.nan
        filled = fill_holes(fill_borders(u), fill_borders(v))
        if filled is None:
            return None
        u, v = filled
    return np.flip(u, axis=0) * scale / dt * 1000, -np.flip(v, axis=0) * scale / dt * 1000


def finalize_field(u, v, x, y, invalid, scale: float = 1.0, dt: float = 1.0):
    """u, v in px (float64, modified in place) -> (x, y, u, v) in mm and m/s, or None to skip."""
    out = finalize_uv(u, v, invalid, scale, dt)
    return None if out is None else (x * scale, y * scale, out[0], out[1])
