"""Device-resident multipass PIV plan: the host side of the fused sm_100a path.

``PIVPlan`` owns everything one (frame shape, window, overlap, passes, mode) configuration needs
on one GPU -- per-pass geometry, the predictor spline operators (uploaded once), and the
workspaces -- and runs a BATCH of image pairs through

    pivb200_pass_first -> [pivb200_predictor -> pivb200_pass_next] * (multipass - 1)

entirely on the device: 1 + 3 * (multipass - 1) kernel launches per batch, no host round trip
between passes (the reference does 3 D2H syncs and 3 host spline evaluations per pass,
PIVbackend.py:413-421, 700-713).  torch is used for device memory and streams only."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import torch

from . import _lib
from .geometry import MAX_WINDOW, get_coordinates, get_field_shape, spline_operator, window_supported

__all__ = ["PassGeometry", "PIVPlan", "pass_schedule", "HostPipeline", "FramePipeline"]


@dataclass
class PassGeometry:
    wind: int
    overlap: int
    n_rows: int
    n_cols: int
    x: np.ndarray = field(repr=False)
    y: np.ndarray = field(repr=False)

    @property
    def n(self) -> int:
        return self.n_rows * self.n_cols


def pass_schedule(wind_size: int, overlap: int, multipass: int, multipass_scale: float):
    """Window / overlap of every pass: ``int(w // scale)`` per extra pass (PIVbackend.py:855-857)."""
    sched = [(int(wind_size), int(overlap))]
    w, o = wind_size, overlap
    for _ in range(int(multipass) - 1):
        w, o = int(w // multipass_scale), int(o // multipass_scale)
        sched.append((w, o))
    return sched


def _geometry(frame_shape, wind, overlap) -> PassGeometry:
    if overlap >= wind:
        raise ValueError("Overlap has to be smaller than the window_size")
    if wind > frame_shape[-2] or wind > frame_shape[-1]:
        raise ValueError("window size cannot be larger than the image")
    if not window_supported(wind):
        raise ValueError(f"interrogation window must be 4..{MAX_WINDOW} px (got {wind}): 16/32/64 px run the fused "
                         "in-register FFT kernels, other sizes the general mixed-radix kernel; there is no "
                         "CPU fallback")
    n_rows, n_cols = (int(v) for v in get_field_shape(frame_shape, wind, overlap)[-2:])
    x, y = get_coordinates(frame_shape, wind, overlap)
    return PassGeometry(wind, overlap, n_rows, n_cols, x, y)


class PIVPlan:
    """Multipass PIV of batches of ``[B, H, W]`` uint8 frame pairs on one CUDA device."""

    def __init__(self, frame_shape, wind_size: int, overlap: int, multipass: int = 1,
                 multipass_mode: str = "CWS", multipass_scale: float = 2.0,
                 device="cuda", val_ratio: float = 1.2, want_ratio: bool = False):
        self.lib = _lib.lib()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("torchpiv_b200 runs on CUDA devices only (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.mode_name = multipass_mode
        self.mode = _lib.MODES[multipass_mode]          # KeyError for unknown modes, like IterModMap
        self.H, self.W = int(frame_shape[-2]), int(frame_shape[-1])
        self.val_ratio = float(val_ratio)
        self.want_ratio = want_ratio
        self.passes: List[PassGeometry] = [
            _geometry((self.H, self.W), w, o)
            for (w, o) in pass_schedule(wind_size, overlap, multipass, multipass_scale)]
        # predictor operators of pass k (old grid = pass k-1, new grid = pass k)
        self._Ay: List[Optional[torch.Tensor]] = [None]
        self._Ax: List[Optional[torch.Tensor]] = [None]
        for prev, cur in zip(self.passes[:-1], self.passes[1:]):
            ay = spline_operator(prev.y[:, 0], cur.y[:, 0])
            ax = spline_operator(prev.x[0, :], cur.x[0, :])
            self._Ay.append(torch.from_numpy(ay).to(self.device))
            self._Ax.append(torch.from_numpy(ax).to(self.device))
        self._ws_pairs = 0
        self._ws = None

    # ------------------------------------------------------------------ workspaces
    def _workspace(self, n_pairs: int):
        if self._ws is not None and self._ws_pairs >= n_pairs:
            return self._ws
        dev, f64 = self.device, torch.float64
        ws = []
        for k, g in enumerate(self.passes):
            w = {
                "u": torch.empty((n_pairs, g.n_rows, g.n_cols), dtype=f64, device=dev),
                "v": torch.empty((n_pairs, g.n_rows, g.n_cols), dtype=f64, device=dev),
                "mask": torch.empty((n_pairs, g.n_rows, g.n_cols), dtype=torch.uint8, device=dev),
            }
            if self.want_ratio:
                w["ratio"] = torch.empty((n_pairs, g.n_rows, g.n_cols), dtype=torch.float32, device=dev)
            if k > 0:
                prev = self.passes[k - 1]
                sdt = torch.float32 if self.mode == _lib.MODE_CWS else torch.int32
                w["tmp"] = torch.empty(3 * n_pairs * prev.n_rows * g.n_cols, dtype=f64, device=dev)
                for name in ("sx", "sy"):
                    w[name] = torch.empty((n_pairs, g.n), dtype=sdt, device=dev)
                for name in ("base_u", "base_v", "pred_u", "pred_v"):
                    w[name] = torch.empty((n_pairs, g.n), dtype=f64, device=dev)
            ws.append(w)
        self._ws, self._ws_pairs = ws, n_pairs
        return ws

    # ------------------------------------------------------------------ execution
    def run(self, frames_a: torch.Tensor, frames_b: torch.Tensor, stream=None, validate: bool = True):
        """frames: uint8 CUDA tensors ``[B, H, W]`` (or ``[H, W]``), last dim contiguous.
        Returns per-pass results of the LAST pass as device tensors ``(u, v, mask)`` shaped
        ``[B, n_rows, n_cols]`` (float64, float64, uint8; views into the plan's workspace --
        they are overwritten by the next ``run``)."""
        if frames_a.dim() == 2:
            frames_a, frames_b = frames_a[None], frames_b[None]
        self._check_frames(frames_a, frames_b)
        B = frames_a.shape[0]
        # the C ABI launches on the CURRENT device: make it the plan's for the duration of the call
        with torch.cuda.device(self.device):
            return self._run(frames_a, frames_b, B, stream, validate)

    def _run(self, frames_a, frames_b, B, stream, validate):
        ws = self._workspace(B)
        L = self.lib
        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        pair_stride, pitch = frames_a.stride(0), frames_a.stride(1)
        ptr = lambda t: t.data_ptr() if t is not None else None   # noqa: E731
        g0, w0 = self.passes[0], ws[0]
        _lib.check(L.pivb200_pass_first(
            ptr(frames_a), ptr(frames_b), B, pair_stride, self.H, self.W, pitch, g0.wind, g0.overlap,
            1 if validate else 0, self.val_ratio, ptr(w0["u"]), ptr(w0["v"]), ptr(w0["mask"]),
            ptr(w0.get("ratio")), stream))
        for k in range(1, len(self.passes)):
            prev, cur, wp, wk = self.passes[k - 1], self.passes[k], ws[k - 1], ws[k]
            _lib.check(L.pivb200_predictor(
                ptr(wp["u"]), ptr(wp["v"]), ptr(wp["mask"]) if validate else None, B,
                prev.n_rows, prev.n_cols, cur.n_rows, cur.n_cols, ptr(self._Ay[k]), ptr(self._Ax[k]),
                self.mode, ptr(wk["tmp"]), ptr(wk["sx"]), ptr(wk["sy"]), ptr(wk["base_u"]),
                ptr(wk["base_v"]), ptr(wk["pred_u"]), ptr(wk["pred_v"]), stream))
            _lib.check(L.pivb200_pass_next(
                ptr(frames_a), ptr(frames_b), B, pair_stride, self.H, self.W, pitch, cur.wind,
                cur.overlap, self.mode, ptr(wk["sx"]), ptr(wk["sy"]), ptr(wk["base_u"]),
                ptr(wk["base_v"]), ptr(wk["pred_u"]), ptr(wk["pred_v"]), 1 if validate else 0,
                self.val_ratio, ptr(wk["u"]), ptr(wk["v"]), ptr(wk["mask"]), ptr(wk.get("ratio")),
                stream))
        last = ws[-1]
        return last["u"][:B], last["v"][:B], last["mask"][:B]

    def pass_results(self, k: int, n_pairs: int):
        """(u, v, mask) device tensors of pass ``k`` from the most recent ``run``."""
        w = self._ws[k]
        return w["u"][:n_pairs], w["v"][:n_pairs], w["mask"][:n_pairs]

    @property
    def launches_per_batch(self):
        """Kernel launches per ``run`` when every pass takes the fused kernels (16/32/64 px windows): one
        per pass plus two predictor kernels between passes.  None when a pass takes the general-size path,
        whose launch count depends on the batch size (chunked scratch buffer)."""
        from .geometry import FUSED_WINDOWS
        if any(g.wind not in FUSED_WINDOWS for g in self.passes):
            return None
        return 1 + 3 * (len(self.passes) - 1)

    @property
    def out_geometry(self) -> PassGeometry:
        return self.passes[-1]

    def _check_frames(self, a: torch.Tensor, b: torch.Tensor) -> None:
        for t in (a, b):
            if t.dtype != torch.uint8 or not t.is_cuda or t.dim() != 3:
                raise TypeError("frames must be uint8 CUDA tensors of shape [B, H, W]")
            if t.device != self.device:
                raise ValueError(f"frames live on {t.device}, the plan on {self.device}")
            if tuple(t.shape[-2:]) != (self.H, self.W):
                raise ValueError(f"frame shape {tuple(t.shape[-2:])} != plan shape {(self.H, self.W)}")
            if t.stride(2) != 1:
                raise ValueError("frame rows must be contiguous")
        if a.shape != b.shape or a.stride() != b.stride():
            raise ValueError("frame_a and frame_b batches must have identical shape and strides")


class HostPipeline:
    """End-to-end batches from HOST memory: pinned staging -> H2D -> PIVPlan -> D2H of the result.

    Two slots are cycled so that the copy-in of batch i+1 (copy stream) overlaps the kernels of
    batch i (compute stream); every batch still pays its own H2D and D2H.  This is the path
    ``OfflinePIV`` and ``bench.py``'s ``e2e`` number go through."""

    def __init__(self, plan: PIVPlan, max_pairs: int):
        self.plan = plan
        dev = plan.device
        g = plan.out_geometry
        self.max_pairs = int(max_pairs)
        self.compute = torch.cuda.Stream(dev)
        self.copy = torch.cuda.Stream(dev)
        self.slots = []
        for _ in range(2):
            self.slots.append({
                "a": torch.empty((max_pairs, plan.H, plan.W), dtype=torch.uint8, device=dev),
                "b": torch.empty((max_pairs, plan.H, plan.W), dtype=torch.uint8, device=dev),
                "u": torch.empty((max_pairs, g.n_rows, g.n_cols), dtype=torch.float64).pin_memory(),
                "v": torch.empty((max_pairs, g.n_rows, g.n_cols), dtype=torch.float64).pin_memory(),
                "m": torch.empty((max_pairs, g.n_rows, g.n_cols), dtype=torch.uint8).pin_memory(),
                "in_ready": torch.cuda.Event(), "done": torch.cuda.Event(), "free": torch.cuda.Event(),
                "n": 0,
            })
        self._next = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def submit(self, host_a: torch.Tensor, host_b: torch.Tensor) -> int:
        """Enqueue one batch (pinned uint8 ``[B, H, W]`` host tensors).  Returns the slot id to
        pass to :meth:`result`.  At most two batches may be in flight."""
        B = host_a.shape[0]
        if B > self.max_pairs:
            raise ValueError("batch larger than the pipeline was built for")
        sid = self._next
        self._next ^= 1
        s = self.slots[sid]
        with torch.cuda.stream(self.copy):
            self.copy.wait_event(s["free"])            # previous user of this slot has been read back
            s["a"][:B].copy_(host_a, non_blocking=True)
            s["b"][:B].copy_(host_b, non_blocking=True)
            s["in_ready"].record(self.copy)
        with torch.cuda.stream(self.compute):
            self.compute.wait_event(s["in_ready"])
            u, v, m = self.plan.run(s["a"][:B], s["b"][:B], stream=self.compute.cuda_stream)
            s["u"][:B].copy_(u, non_blocking=True)
            s["v"][:B].copy_(v, non_blocking=True)
            s["m"][:B].copy_(m, non_blocking=True)
            s["done"].record(self.compute)
            s["free"].record(self.compute)
        s["n"] = B
        self.h2d_bytes += 2 * host_a[:B].numel()
        self.d2h_bytes += B * (s["u"][0].numel() * 16 + s["m"][0].numel())
        return sid

    def result(self, sid: int):
        """Block until batch ``sid`` is back on the host; returns NumPy views ``(u, v, invalid)``
        (valid until the slot is reused two submits later)."""
        s = self.slots[sid]
        s["done"].synchronize()
        B = s["n"]
        return s["u"][:B].numpy(), s["v"][:B].numpy(), s["m"][:B].numpy().astype(bool)


class FramePipeline:
    """Batches of pairs from pinned host FRAME stacks (the ``OfflinePIV`` data path).

    A slot holds up to ``2 * batch_pairs`` frames on the host (pinned; image decoders write straight
    into it) and on the device.  ``submit`` uploads only the frames a batch really has -- ``K + 1``
    for a chained (sequential-folder) batch, whose two frame stacks are the overlapping device views
    ``[0:K]`` and ``[1:K+1]``, else ``2 K`` -- on the copy stream, runs the plan on the compute stream
    and brings ``u, v, mask`` back; two slots alternate so that decoding / upload of batch i+1
    overlaps the kernels of batch i."""

    def __init__(self, plan: PIVPlan, batch_pairs: int, post=None, stats=None):
        self.plan = plan
        # post(u, v, mask, stream): in-place device post-processing between the last pass and the D2H
        # copy (postprocess_device.StencilPost); stats: postprocess_device.FieldStatistics fed after it
        self.post = post
        self.stats = stats
        dev = plan.device
        g = plan.out_geometry
        K = self.batch_pairs = int(batch_pairs)
        self.compute = torch.cuda.Stream(dev)
        self.copy = torch.cuda.Stream(dev)
        self.slots = []
        for _ in range(2):
            self.slots.append({
                "host": torch.empty((2 * K, plan.H, plan.W), dtype=torch.uint8).pin_memory(),
                "dev": torch.empty((2 * K, plan.H, plan.W), dtype=torch.uint8, device=dev),
                "u": torch.empty((K, g.n_rows, g.n_cols), dtype=torch.float64).pin_memory(),
                "v": torch.empty((K, g.n_rows, g.n_cols), dtype=torch.float64).pin_memory(),
                "m": torch.empty((K, g.n_rows, g.n_cols), dtype=torch.uint8).pin_memory(),
                "uploaded": torch.cuda.Event(), "done": torch.cuda.Event(), "n": 0,
            })
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def host_frames(self, sid: int) -> np.ndarray:
        """Writable ``[2K, H, W]`` uint8 view of slot ``sid``'s pinned staging memory.  Blocks until
        the previous upload out of this slot has finished."""
        s = self.slots[sid]
        s["uploaded"].synchronize()
        return s["host"].numpy()

    def submit(self, sid: int, n_pairs: int, chained: bool, keep=None) -> None:
        """``keep``: optional per-pair flags; pairs flagged False (unreadable frames) are left out of
        the statistics."""
        K = int(n_pairs)
        if not 0 < K <= self.batch_pairs:
            raise ValueError("batch larger than the pipeline was built for")
        s = self.slots[sid]
        n_frames = K + 1 if chained else 2 * K
        with torch.cuda.stream(self.copy):
            self.copy.wait_event(s["done"])        # the kernels of the slot's previous batch have read it
            s["dev"][:n_frames].copy_(s["host"][:n_frames], non_blocking=True)
            s["uploaded"].record(self.copy)
        fa = s["dev"][:K]
        fb = s["dev"][1:K + 1] if chained else s["dev"][K:2 * K]
        with torch.cuda.stream(self.compute):
            self.compute.wait_event(s["uploaded"])
            u, v, m = self.plan.run(fa, fb, stream=self.compute.cuda_stream)
            if self.post is not None:
                self.post(u, v, m, stream=self.compute.cuda_stream)
            if self.stats is not None:
                if keep is None or all(keep):
                    self.stats.add(u, v, stream=self.compute.cuda_stream)
                elif any(keep):
                    sel = torch.tensor([i for i, k in enumerate(keep) if k], device=u.device)
                    self.stats.add(u[sel], v[sel], stream=self.compute.cuda_stream)
            s["u"][:K].copy_(u, non_blocking=True)
            s["v"][:K].copy_(v, non_blocking=True)
            s["m"][:K].copy_(m, non_blocking=True)
            s["done"].record(self.compute)
        s["n"] = K
        self.h2d_bytes += n_frames * self.plan.H * self.plan.W
        self.d2h_bytes += K * (s["u"][0].numel() * 16 + s["m"][0].numel())

    def result(self, sid: int):
        """``(u, v, invalid)`` NumPy arrays of the slot's batch (views, valid until the slot is
        submitted again)."""
        s = self.slots[sid]
        s["done"].synchronize()
        K = s["n"]
        return s["u"][:K].numpy(), s["v"][:K].numpy(), s["m"][:K].numpy().astype(bool)
