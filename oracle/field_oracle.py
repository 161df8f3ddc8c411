"""CPU oracle for what follows the correlation passes  --  TEST INFRASTRUCTURE ONLY.

Only ``tests/`` may import this module (see the header of ``oracle/piv_oracle.py``).

Parity status per function:

* ``statistics_table``  PINNED.  Restates the statistics block of the reference's worker
  (``src/torchPIV/workers.py:79-119``: stacked mean, Reynolds stresses, ``np.gradient``,
  vorticity / strain).  ``tests/golden/make_golden.py::gen_statistics`` executes those very
  source lines of the unmodified reference (read from ``/root/reference`` at generation time,
  nothing copied) and commits inputs' SHA + outputs as ``tests/golden/statistics.npz``.
* ``normalized_median_test`` / ``stencil_replace``  PARITY UNPINNED.  The reference has neither a
  normalised median test nor a stencil replacement ("N* vs reference" items
  2-3); these restate the published algorithm (Westerweel & Scarano, Exp. Fluids 39 (2005)
  1096-1100: residual ``|u - median(nb)| / (median|nb - median(nb)| + eps)`` over the 3x3
  neighbourhood, threshold 2, eps 0.1 px) and a Jacobi-sweep median fill, in plain Python loops.
"""
from __future__ import annotations

import numpy as np

TABLE_KEYS = ("x[mm]", "y[mm]", "Vx[m/s]", "Vy[m/s]", "(vx-Vx)(vy-Vy)[m^2/s^2]", "(vx-Vx)^2[m^2/s^2]",
              "(vy-Vy)^2[m^2/s^2]", "dVx/dx[1/s]", "dVx/dy[1/s]", "dVy/dx[1/s]", "dVy/dy[1/s]", "W[1/s]",
              "S[1/s]")


def statistics_table(x, y, u_list, v_list) -> dict:
    """workers.py:79-119.  ``u_list`` / ``v_list``: the fields the generator yielded (m/s), x / y in mm."""
    u_inst = np.stack([np.asarray(u, dtype=np.float64) for u in u_list], axis=0)
    v_inst = np.stack([np.asarray(v, dtype=np.float64) for v in v_list], axis=0)
    avg_u = np.mean(u_inst, axis=0, dtype=np.float64)                      # workers.py:83-84
    avg_v = np.mean(v_inst, axis=0, dtype=np.float64)
    uu = np.mean((u_inst - avg_u) ** 2, axis=0, dtype=np.float64)          # workers.py:86-90
    vv = np.mean((v_inst - avg_v) ** 2, axis=0, dtype=np.float64)
    uv = np.mean((u_inst - avg_u) * (v_inst - avg_v), axis=0, dtype=np.float64)
    mid_i, mid_j = x.shape[-2] // 2, x.shape[-1] // 2                      # workers.py:95-99
    dx = (x[mid_i, mid_j + 1] - x[mid_i, mid_j]) / 1000
    dy = (y[mid_i + 1, mid_j] - y[mid_i, mid_j]) / 1000
    # NB the reference passes (dx, dy) as the spacings of axis 0 and axis 1 and names the axis-0
    # derivative "d/dy": kept as is
    dUy, dUx = np.gradient(avg_u, dx, dy, edge_order=2)
    dVy, dVx = np.gradient(avg_v, dx, dy, edge_order=2)
    values = (x, y, avg_u, avg_v, uv, uu, vv, dUx, dUy, dVx, dVy, dVx - dUy, dVx + dUy)
    return dict(zip(TABLE_KEYS, values))


def _ring(field, bad, r, c):
    n_rows, n_cols = field.shape
    out = []
    for dr in (-1, 0, 1):
        for dc in (-1, 0, 1):
            rr, cc = r + dr, c + dc
            if (dr == 0 and dc == 0) or rr < 0 or rr >= n_rows or cc < 0 or cc >= n_cols:
                continue
            if bad is not None and bad[rr, cc]:
                continue
            out.append(field[rr, cc])
    return np.array(out, dtype=np.float64)


def normalized_median_test(u, v, mask=None, threshold: float = 2.0, eps: float = 0.1) -> np.ndarray:
    """One field ``[n_rows, n_cols]``.  Flagged neighbours are ignored; flagged vectors stay flagged;
    vectors with fewer than two usable neighbours are not tested."""
    n_rows, n_cols = u.shape
    out = np.zeros((n_rows, n_cols), dtype=bool)
    for r in range(n_rows):
        for c in range(n_cols):
            if mask is not None and mask[r, c]:
                out[r, c] = True
                continue
            a, b = _ring(u, mask, r, c), _ring(v, mask, r, c)
            if a.size < 2:
                continue
            mu, mv = np.median(a), np.median(b)
            ru, rv = np.median(np.abs(a - mu)), np.median(np.abs(b - mv))
            out[r, c] = (abs(u[r, c] - mu) / (ru + eps) > threshold) or (abs(v[r, c] - mv) / (rv + eps) > threshold)
    return out


def stencil_replace(u, v, invalid, max_sweeps: int):
    """Jacobi sweeps: a flagged vector with >= 1 usable 3x3 neighbour becomes their median (values of
    the previous sweep) and is unflagged.  The sweep count is rounded up to even; vectors never
    reached become 0 and stay flagged.  Returns new ``(u, v, invalid)``."""
    u, v, bad = u.astype(np.float64).copy(), v.astype(np.float64).copy(), invalid.astype(bool).copy()
    for _ in range(max_sweeps + (max_sweeps & 1)):
        nu, nv, nb = u.copy(), v.copy(), bad.copy()
        for r, c in np.argwhere(bad):
            a, b = _ring(u, bad, r, c), _ring(v, bad, r, c)
            if a.size:
                nu[r, c], nv[r, c], nb[r, c] = np.median(a), np.median(b), False
        u, v, bad = nu, nv, nb
    u[bad] = 0.0
    v[bad] = 0.0
    return u, v, bad
