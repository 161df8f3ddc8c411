"""On-device post-processing of batches of vector fields.

Thin host wrappers over the C ABI entry points ``pivb200_nmt`` / ``pivb200_replace`` /
``pivb200_stats_accumulate`` (``include/pivb200.h``, kernels in ``csrc/field_ops.cu``):

* :func:`normalized_median_test`, :func:`replace_invalid` and :class:`StencilPost` -- the
  "stencil" alternative to the reference's host hole filling (``interpolate_boarders`` +
  ``fillMissingValues``, PIVbackend.py:266-344, 884-892).  A documented deviation: the reference
  has no counterpart, so values of replaced vectors differ from its Delaunay interpolation, and a
  pair is never skipped.  ``OfflinePIV(..., replace="stencil")`` runs it on the compute stream
  between the last pass and the D2H copy.
* :class:`FieldStatistics` -- streaming form of the statistics block of the reference's worker
  (workers.py:79-119): five running moments per vector stay on the device, the mean / Reynolds-stress /
  gradient table is formed once at the end.  The reference stacks every field of a run in host
  memory.

Fields are float64 CUDA tensors ``[B, n_rows, n_cols]``; masks are uint8 (1 = invalid).  There is no
CPU fallback: CPU tensors raise."""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _lib

__all__ = ["normalized_median_test", "replace_invalid", "StencilPost", "FieldStatistics", "merge_states", "finalize_field_stencil",
           "TABLE_KEYS"]

TABLE_KEYS = ("x[mm]", "y[mm]", "Vx[m/s]", "Vy[m/s]", "(vx-Vx)(vy-Vy)[m^2/s^2]", "(vx-Vx)^2[m^2/s^2]",
              "(vy-Vy)^2[m^2/s^2]", "dVx/dx[1/s]", "dVx/dy[1/s]", "dVy/dx[1/s]", "dVy/dy[1/s]", "W[1/s]",
              "S[1/s]")


def _stream(device, stream):
    return torch.cuda.current_stream(device).cuda_stream if stream is None else stream


def _check_fields(u: torch.Tensor, v: torch.Tensor):
    for t in (u, v):
        if not t.is_cuda:
            raise RuntimeError("torchpiv_b200 post-processing runs on CUDA tensors only (no CPU fallback)")
        if t.dtype != torch.float64 or t.dim() != 3 or not t.is_contiguous():
            raise TypeError("fields must be contiguous float64 CUDA tensors [B, n_rows, n_cols]")
    if u.shape != v.shape or u.device != v.device:
        raise ValueError("u and v must have the same shape and device")
    return tuple(int(s) for s in u.shape)


def _check_mask(mask: torch.Tensor, like: torch.Tensor):
    if mask.dtype != torch.uint8 or mask.shape != like.shape or mask.device != like.device or not mask.is_contiguous():
        raise TypeError("masks must be contiguous uint8 tensors shaped like the fields, on the same device")


def normalized_median_test(u: torch.Tensor, v: torch.Tensor, mask: Optional[torch.Tensor] = None,
                           threshold: float = 2.0, eps: float = 0.1, out: Optional[torch.Tensor] = None,
                           stream=None) -> torch.Tensor:
    """Westerweel-Scarano normalised median test on the 3x3 neighbourhood.  Returns a uint8 tensor:
    ``mask | outlier``.  Neighbours flagged in ``mask`` are ignored."""
    B, n_rows, n_cols = _check_fields(u, v)
    if mask is not None:
        _check_mask(mask, u)
    if out is None:
        out = torch.empty(u.shape, dtype=torch.uint8, device=u.device)
    else:
        _check_mask(out, u)
    with torch.cuda.device(u.device):      # the C ABI launches on the CURRENT device
        _lib.check(_lib.lib().pivb200_nmt(u.data_ptr(), v.data_ptr(), mask.data_ptr() if mask is not None else None,
                                          B, n_rows, n_cols, float(threshold), float(eps), out.data_ptr(),
                                          _stream(u.device, stream)))
    return out


def replace_workspace(u: torch.Tensor) -> torch.Tensor:
    B, n_rows, n_cols = (int(s) for s in u.shape)
    nbytes = int(_lib.lib().pivb200_replace_workspace_bytes(B, n_rows, n_cols))
    return torch.empty((nbytes + 7) // 8, dtype=torch.float64, device=u.device)


def replace_invalid(u: torch.Tensor, v: torch.Tensor, invalid: torch.Tensor, max_sweeps: Optional[int] = None,
                    workspace: Optional[torch.Tensor] = None, stream=None) -> None:
    """In place: flagged vectors become the median of their usable 3x3 neighbours (Jacobi sweeps, holes
    fill from the rim inwards; default sweep count = the longer field side, enough for any hole).
    ``invalid`` is cleared where a value was produced."""
    B, n_rows, n_cols = _check_fields(u, v)
    _check_mask(invalid, u)
    if max_sweeps is None:
        max_sweeps = max(n_rows, n_cols)
    if workspace is None:
        workspace = replace_workspace(u)
    with torch.cuda.device(u.device):      # the C ABI launches on the CURRENT device
        _lib.check(_lib.lib().pivb200_replace(u.data_ptr(), v.data_ptr(), invalid.data_ptr(), B, n_rows, n_cols,
                                              int(max_sweeps), workspace.data_ptr(), _stream(u.device, stream)))


class StencilPost:
    """``post(u, v, mask, stream)`` hook for :class:`engine.FramePipeline`: optional normalised median
    test, then stencil replacement, in place on the plan's result tensors."""

    def __init__(self, nmt: bool = False, threshold: float = 2.0, eps: float = 0.1, max_sweeps: int = 8):
        self.nmt, self.threshold, self.eps, self.max_sweeps = nmt, threshold, eps, int(max_sweeps)
        self._ws = None
        self._flags = None

    def __call__(self, u: torch.Tensor, v: torch.Tensor, mask: torch.Tensor, stream=None) -> None:
        _check_fields(u, v)          # in place: views of the plan's workspace must be contiguous
        _check_mask(mask, u)
        if self._flags is None or self._flags.shape != u.shape or self._flags.device != u.device:
            self._ws = replace_workspace(u)
            self._flags = torch.empty(u.shape, dtype=torch.uint8, device=u.device)
        flags = mask
        if self.nmt:
            flags = normalized_median_test(u, v, mask, self.threshold, self.eps, out=self._flags, stream=stream)
        replace_invalid(u, v, flags, self.max_sweeps, self._ws, stream=stream)
        if flags is not mask:
            mask.copy_(flags)        # caller's current stream == `stream` (engine.FramePipeline)


def finalize_field_stencil(u, v, x, y, invalid, scale: float = 1.0, dt: float = 1.0):
    """Host tail of ``OfflinePIV`` in stencil mode: the holes were already filled on the device, only
    the flip / sign / unit conversion of PIVbackend.py:894-900 is left."""
    u = np.flip(u, axis=0) * scale / dt * 1000
    v = -np.flip(v, axis=0) * scale / dt * 1000
    return x * scale, y * scale, u, v


def merge_states(states):
    """Combine ``(count, moments)`` states of several accumulators (one per GPU / shard) into one, with
    the same pairwise update the kernel uses.  Pure NumPy: runs on rank 0 after the gather."""
    total, acc = 0, None
    for count, mom in states:
        if count == 0:
            continue
        mom = np.array(mom, dtype=np.float64)
        if acc is None:
            total, acc = count, mom
            continue
        n = total + count
        w = total * count / n
        du, dv = mom[0] - acc[0], mom[1] - acc[1]
        acc[0] += du * count / n
        acc[1] += dv * count / n
        acc[2] += mom[2] + du * du * w
        acc[3] += mom[3] + dv * dv * w
        acc[4] += mom[4] + du * dv * w
        total = n
    if acc is None:
        raise ValueError("no field was accumulated")
    return total, acc


class FieldStatistics:
    """Running mean / Reynolds stresses of a sequence of vector fields, accumulated on the device.

    ``add(u, v)`` takes the fields as they leave the passes: pixel units, un-flipped, float64
    ``[B, n_rows, n_cols]`` CUDA tensors (NumPy arrays are uploaded).  ``table(x, y, scale, dt)``
    applies the generator's flip / sign / unit conversion (linear, so it commutes with the sums) and
    returns the reference's statistics table (workers.py:100-119; same keys, same order)."""

    def __init__(self, n_rows: int, n_cols: int, device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("torchpiv_b200 runs on CUDA devices only (no CPU fallback)")
        self.n_rows, self.n_cols = int(n_rows), int(n_cols)
        # mean u, mean v, then centred second-moment sums uu, vv, uv (merged batch by batch on the device)
        self.moments_dev = torch.zeros((5, self.n_rows, self.n_cols), dtype=torch.float64, device=self.device)
        # the zero fill runs on the current stream, add() may run on another (non-blocking) one: order them
        self._zeroed = torch.cuda.Event()
        self._zeroed.record(torch.cuda.current_stream(self.device))
        self.count = 0

    def add(self, u, v, stream=None) -> None:
        if isinstance(u, np.ndarray):
            u = torch.from_numpy(np.ascontiguousarray(u, dtype=np.float64)).to(self.device)
            v = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).to(self.device)
        if u.dim() == 2:
            u, v = u[None], v[None]
        u, v = u.contiguous(), v.contiguous()
        B, n_rows, n_cols = _check_fields(u, v)
        if (n_rows, n_cols) != (self.n_rows, self.n_cols):
            raise ValueError("field shape does not match the accumulator")
        if self._zeroed is not None:
            # first use: `stream` is a raw handle (or None = current stream) -- a one-off host wait is the simplest order
            self._zeroed.synchronize()
            self._zeroed = None
        with torch.cuda.device(self.device):      # the C ABI launches on the CURRENT device
            _lib.check(_lib.lib().pivb200_stats_accumulate(u.data_ptr(), v.data_ptr(), B, n_rows, n_cols, self.count,
                                                           self.moments_dev.data_ptr(), _stream(self.device, stream)))
        self.count += B

    def state(self):
        """``(count, moments)`` with the moments as a NumPy array ``[5, n_rows, n_cols]`` -- what a rank
        sends to rank 0 (``sharding.gather_results``) to be combined by :func:`merge_states`."""
        return self.count, self.moments_dev.cpu().numpy()

    def moments(self, state=None):
        """(mean_u, mean_v, uu, vv, uv) in pixel units, un-flipped, as NumPy arrays."""
        count, s = self.state() if state is None else state
        if count == 0:
            raise ValueError("no field was accumulated")
        return s[0], s[1], s[2] / count, s[3] / count, s[4] / count

    def table(self, x: np.ndarray, y: np.ndarray, scale: float = 1.0, dt: float = 1.0, state=None) -> dict:
        """``x``, ``y``: window centres in px (``get_coordinates``).  ``state``: a merged state of several
        accumulators (multi-GPU runs) instead of this object's own."""
        mu, mv, uu, vv, uv = self.moments(state)
        k = scale / dt * 1000
        avg_u = np.flip(mu, axis=0) * k
        avg_v = -np.flip(mv, axis=0) * k
        uu, vv, uv = np.flip(uu, axis=0) * k * k, np.flip(vv, axis=0) * k * k, -np.flip(uv, axis=0) * k * k
        x, y = x * scale, y * scale
        mid_i, mid_j = x.shape[-2] // 2, x.shape[-1] // 2
        dx = (x[mid_i, mid_j + 1] - x[mid_i, mid_j]) / 1000
        dy = (y[mid_i + 1, mid_j] - y[mid_i, mid_j]) / 1000
        # the reference hands (dx, dy) to np.gradient as the spacings of axis 0 and axis 1 (workers.py:98-99)
        dUy, dUx = np.gradient(avg_u, dx, dy, edge_order=2)
        dVy, dVx = np.gradient(avg_v, dx, dy, edge_order=2)
        values = (x, y, avg_u, avg_v, uv, uu, vv, dUx, dUy, dVx, dVy, dVx - dUy, dVx + dUy)
        return dict(zip(TABLE_KEYS, values))
