"""Reference-facing API of the B200-native PIV hot path.

Every public name, argument and return convention mirrors ``src/torchPIV/PIVbackend.py`` of
NikNazarov/TorchPIV ("PB"), so scripts written against the reference run unchanged:

    from torchpiv_b200 import OfflinePIV
    for x, y, u, v in OfflinePIV(folder, device, "bmp", wind_size=64, overlap=32,
                                 multipass=2, multipass_mode="CWS")():
        ...

All numerical work goes through the C ABI of ``libpivb200.so`` (hand-written sm_100a kernels).
There is no CPU path: ``device="cpu"`` raises, and a missing library raises on first use.
Differences from the reference, all documented in DESIGN.md: interrogation windows are 4..256 px
(16/32/64 px take the fused kernels, other sizes -- odd ones included -- a general mixed-radix kernel); the
first pass is evaluated in FP32 (the reference uses FP64) -- results agree within 1e-3 px; exact
ties between correlation values may resolve differently.
"""
from __future__ import annotations

from typing import Generator, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .dataset import (PIVDataset, ToTensor, natural_keys, plan_batches, read_gray,  # noqa: F401
                      read_gray_into, shard_range)
from .engine import PIVPlan, pass_schedule
from .geometry import get_coordinates, get_field_shape, spline_operator  # noqa: F401
from .postprocess import finalize_field

__all__ = [
    "DeviceMap", "IterModMap", "OfflinePIV", "extended_search_area_piv", "piv_iteration_CWS",
    "piv_iteration_DWS", "correalte_fft", "correlation_to_displacement", "moving_window_array",
    "get_field_shape", "get_coordinates", "biliniar_interpolation_CWS", "interpolation_DWS",
    "PIVDataset", "ToTensor", "natural_keys",
]


# --------------------------------------------------------------------------------------------
# device registry (PB:13-18)
# --------------------------------------------------------------------------------------------
class _Devices(dict):
    """name -> torch.device.  Like the reference, every CUDA device is registered under
    ``torch.cuda.get_device_name(i)`` (identical GPUs collide; the last index wins); in addition
    ``"cuda"`` and ``"cuda:N"`` address a specific GPU.  ``"cpu"`` is listed so that the lookup
    itself behaves like the reference's, but using it raises: there is no CPU fallback."""

    def __init__(self):
        super().__init__()
        self._filled = False

    def _fill(self):
        if self._filled:
            return
        self._filled = True
        for i in range(torch.cuda.device_count()):
            self.setdefault(f"cuda:{i}", torch.device("cuda", i))
        if torch.cuda.device_count():
            self.setdefault("cuda", torch.device("cuda", 0))
        for i in range(torch.cuda.device_count()):
            dict.__setitem__(self, torch.cuda.get_device_name(i), torch.device("cuda", i))
        self.setdefault("cpu", torch.device("cpu"))

    def __getitem__(self, key):
        self._fill()
        if isinstance(key, torch.device):
            return key
        return super().__getitem__(key)

    def keys(self):
        self._fill()
        return super().keys()

    def __contains__(self, key):
        self._fill()
        return super().__contains__(key)


class DeviceMap:
    devicies = _Devices()


def _cuda_device(device) -> torch.device:
    dev = DeviceMap.devicies[device] if not isinstance(device, torch.device) else device
    if dev.type != "cuda":
        raise RuntimeError("torchpiv_b200 has no CPU path: pick a CUDA device (e.g. 'cuda:0' or "
                           "torch.cuda.get_device_name(0))")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def _stream(dev: torch.device):
    return torch.cuda.current_stream(dev).cuda_stream


def _frame_on(frame, dev: torch.device) -> torch.Tensor:
    if isinstance(frame, np.ndarray):
        frame = torch.from_numpy(np.ascontiguousarray(frame))
    if frame.dtype != torch.uint8:
        raise TypeError("frames must be uint8 grey-level images (the reference decodes 8-bit "
                        "grayscale, PB:136-137)")
    return frame.to(dev, non_blocking=True).contiguous()


# --------------------------------------------------------------------------------------------
# function-level API
# --------------------------------------------------------------------------------------------
def moving_window_array(array: torch.Tensor, window_size, overlap) -> torch.Tensor:
    """``[n_rows*n_cols, w, w]`` stack of interrogation windows (PB:220-247).  Pure memory
    plumbing (strided view + copy); the fused passes never materialise this tensor -- they read
    windows with TMA tile loads straight from the frame."""
    h, w = array.shape[-2:]
    step = window_size - overlap
    n_r = int((h - window_size) / step) + 1
    n_c = int((w - window_size) / step) + 1
    s0, s1 = array.stride()[-2:]
    view = torch.as_strided(array, (n_r, n_c, window_size, window_size), (s0 * step, s1 * step, s0, s1),
                            array.storage_offset())
    return view.reshape(-1, window_size, window_size)


def correalte_fft(images_a: torch.Tensor, images_b: torch.Tensor) -> torch.Tensor:
    """fft-shifted circular cross-correlation of two window stacks ``[c, w, w]`` (PB:249-257),
    ``corr[s] = sum_x a[x] b[x+s]`` with zero lag at ``(w/2, w/2)``; for odd ``w`` the map is ``[c, w, w-1]`` like
    the reference's (``irfft2`` without an explicit size).  uint8 and float32 inputs
    give float32 like the reference; float64 inputs are evaluated in FP32 and returned as
    float64."""
    if images_a.shape != images_b.shape or images_a.dim() != 3 or images_a.shape[-1] != images_a.shape[-2]:
        raise ValueError("expected two [c, w, w] stacks of square windows")
    dev = _cuda_device(images_a.device)
    out_dtype = torch.float64 if images_a.dtype == torch.float64 else torch.float32
    if images_a.dtype == torch.uint8 and images_b.dtype == torch.uint8:
        code, a, b = 1, images_a.contiguous(), images_b.contiguous()
    else:
        code, a, b = 0, images_a.float().contiguous(), images_b.float().contiguous()
    c, w, _ = a.shape
    # odd windows: [c, w, w - 1], what torch.fft.irfft2 returns in the reference (PB:255)
    corr = torch.empty((c, w, w - (w & 1)), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):      # the C ABI launches on the CURRENT device
        _lib.check(_lib.lib().pivb200_correlate(a.data_ptr(), b.data_ptr(), code, c, w, corr.data_ptr(),
                                                _stream(dev)))
    return corr.to(out_dtype)


def correlation_to_displacement(corr: torch.Tensor, n_rows, n_cols, validate: bool = True,
                                val_ratio=1.2, validation_window=3
                                ) -> Tuple[np.ndarray, np.ndarray, Optional[np.ndarray]]:
    """Peak location, 3-point log-Gaussian sub-pixel fit and peak-ratio validation of correlation
    maps ``[c, d, k]`` (PB:360-422).  Like the reference it modifies ``corr`` in place (adds
    1e-7, zeroes the patch around the peak when validating) and returns float64 NumPy arrays
    ``u, v [n_rows, n_cols]`` and the boolean INVALID mask (None without validation)."""
    if corr.dtype not in (torch.float32, torch.float64) or not corr.is_contiguous():
        raise TypeError("corr must be a contiguous float32/float64 CUDA tensor")
    dev = _cuda_device(corr.device)
    c, d, k = corr.shape
    u = torch.empty(c, dtype=torch.float64, device=dev)
    v = torch.empty(c, dtype=torch.float64, device=dev)
    m = torch.empty(c, dtype=torch.uint8, device=dev) if validate else None
    with torch.cuda.device(dev):      # the C ABI launches on the CURRENT device
        _lib.check(_lib.lib().pivb200_corr_to_disp(
            corr.data_ptr(), 0 if corr.dtype == torch.float32 else 1, c, d, k, 1 if validate else 0,
            float(val_ratio), int(validation_window), u.data_ptr(), v.data_ptr(),
            m.data_ptr() if validate else None, _stream(dev)))
    mask = m.cpu().numpy().astype(bool).reshape(n_rows, n_cols) if validate else None
    return u.cpu().numpy().reshape(n_rows, n_cols), v.cpu().numpy().reshape(n_rows, n_cols), mask


def biliniar_interpolation_CWS(array: torch.Tensor, grid: torch.Tensor, vel_x: torch.Tensor,
                               vel_y: torch.Tensor) -> torch.Tensor:
    """Per-window bilinear translation with the reference's argument layout (PB:147-194):
    ``grid`` int64 ``[c, w, h]`` flat pixel indices, ``vel_*`` float32 ``[c, 1, 1]``.  Bit-exact."""
    dev = _cuda_device(array.device)
    frame = array.contiguous()
    if frame.dtype != torch.uint8:
        raise TypeError("array must be a uint8 frame")
    grid = grid.to(dev, torch.int64).contiguous()
    vx = vel_x.to(dev, torch.float32).reshape(-1).contiguous()
    vy = vel_y.to(dev, torch.float32).reshape(-1).contiguous()
    out = torch.empty(grid.shape, dtype=torch.float32, device=dev)
    epw = grid[0].numel()
    with torch.cuda.device(dev):      # the C ABI launches on the CURRENT device
        _lib.check(_lib.lib().pivb200_bilinear_cws(frame.data_ptr(), frame.shape[-2], frame.shape[-1],
                                                   grid.data_ptr(), grid.numel(), epw, vx.data_ptr(),
                                                   vy.data_ptr(), out.data_ptr(), _stream(dev)))
    return out


def interpolation_DWS(array: torch.Tensor, grid: torch.Tensor, vel_x: torch.Tensor,
                      vel_y: torch.Tensor) -> torch.Tensor:
    """Per-window integer translation by flat-index arithmetic (PB:197-216); uint8 in, uint8 out."""
    dev = _cuda_device(array.device)
    frame = array.contiguous()
    if frame.dtype != torch.uint8:
        raise TypeError("array must be a uint8 frame")
    grid = grid.to(dev, torch.int64).contiguous()
    vx = vel_x.to(dev, torch.int64).reshape(-1).contiguous()
    vy = vel_y.to(dev, torch.int64).reshape(-1).contiguous()
    out = torch.empty(grid.shape, dtype=torch.uint8, device=dev)
    epw = grid[0].numel()
    with torch.cuda.device(dev):      # the C ABI launches on the CURRENT device
        _lib.check(_lib.lib().pivb200_shift_dws(frame.data_ptr(), frame.shape[-2], frame.shape[-1],
                                                grid.data_ptr(), grid.numel(), epw, vx.data_ptr(),
                                                vy.data_ptr(), out.data_ptr(), _stream(dev)))
    return out


def extended_search_area_piv(frame_a, frame_b, window_size=32, overlap=0, validate: bool = False,
                             validation_ratio: float = 1.2) -> Tuple[np.ndarray, ...]:
    """First (zero-order) PIV pass, one fused kernel launch (PB:459-520).  ``frame_a/b`` are
    uint8 ``[H, W]`` CUDA tensors.  Returns ``u, v, x, y, validation_mask`` as NumPy arrays
    (float64; mask bool, True = invalid, None when ``validate`` is False)."""
    if overlap >= window_size:
        raise ValueError("Overlap has to be smaller than the window_size")
    if (window_size > frame_a.shape[-2]) or (window_size > frame_a.shape[-1]):
        raise ValueError("window size cannot be larger than the image")
    dev = _cuda_device(frame_a.device)
    fa, fb = _frame_on(frame_a, dev), _frame_on(frame_b, dev)
    h, w = fa.shape[-2:]
    n_rows, n_cols = (int(v) for v in get_field_shape((h, w), window_size, overlap))
    x, y = get_coordinates((h, w), window_size, overlap)
    u = torch.empty((n_rows, n_cols), dtype=torch.float64, device=dev)
    v = torch.empty_like(u)
    m = torch.empty((n_rows, n_cols), dtype=torch.uint8, device=dev) if validate else None
    with torch.cuda.device(dev):      # the C ABI launches on the CURRENT device
        _lib.check(_lib.lib().pivb200_pass_first(
            fa.data_ptr(), fb.data_ptr(), 1, 0, h, w, fa.stride(0), int(window_size), int(overlap),
            1 if validate else 0, float(validation_ratio), u.data_ptr(), v.data_ptr(),
            m.data_ptr() if validate else None, None, _stream(dev)))
    mask = m.cpu().numpy().astype(bool) if validate else None
    return u.cpu().numpy(), v.cpu().numpy(), x, y, mask


class _PivIteration:
    """Later pass (PB:677-812): resample the previous field onto this pass's grid (bicubic
    spline, as a precomputed operator applied on the device), shift the windows by -/+ half the
    predictor, correlate, validate, and merge with the predictor -- two small kernels plus one
    fused kernel.  Same constructor / call signature as the reference classes."""

    MODE = "CWS"

    def __init__(self, frame_shape, wind_size, overlap, device) -> None:
        self.device = _cuda_device(device)
        self.frame_shape = (int(frame_shape[-2]), int(frame_shape[-1]))
        self.wind_size, self.overlap = int(wind_size), int(overlap)
        self.n_rows, self.n_cols = (int(v) for v in get_field_shape(self.frame_shape, wind_size, overlap))
        self.x, self.y = get_coordinates(self.frame_shape, wind_size, overlap)
        self.slice_x, self.slice_y = self.x[0, :], self.y[:, 0]
        self._ops = {}

    def _operators(self, x0: np.ndarray, y0: np.ndarray):
        key = (y0[:, 0].tobytes(), x0[0, :].tobytes())
        if key not in self._ops:
            ay = spline_operator(y0[:, 0], self.slice_y)
            ax = spline_operator(x0[0, :], self.slice_x)
            self._ops[key] = (torch.from_numpy(ay).to(self.device), torch.from_numpy(ax).to(self.device))
        return self._ops[key]

    def __call__(self, frame_a, frame_b, x0: np.ndarray, y0: np.ndarray, u0: np.ndarray,
                 v0: np.ndarray, validation_mask: Optional[np.ndarray]) -> Tuple[np.ndarray, ...]:
        dev, L = self.device, _lib.lib()
        fa, fb = _frame_on(frame_a, dev), _frame_on(frame_b, dev)
        ay, ax = self._operators(np.asarray(x0), np.asarray(y0))
        n0, m0 = u0.shape
        n1, m1 = self.n_rows, self.n_cols
        f64 = torch.float64
        up = torch.from_numpy(np.ascontiguousarray(u0, dtype=np.float64)).to(dev)
        vp = torch.from_numpy(np.ascontiguousarray(v0, dtype=np.float64)).to(dev)
        validate = validation_mask is not None
        mp = (torch.from_numpy(np.ascontiguousarray(validation_mask).astype(np.uint8)).to(dev)
              if validate else None)
        mode = _lib.MODES[self.MODE]
        sdt = torch.float32 if mode == _lib.MODE_CWS else torch.int32
        n = n1 * m1
        tmp = torch.empty(3 * n0 * m1, dtype=f64, device=dev)
        sx, sy = torch.empty(n, dtype=sdt, device=dev), torch.empty(n, dtype=sdt, device=dev)
        base_u, base_v, pred_u, pred_v, u, v = (torch.empty(n, dtype=f64, device=dev) for _ in range(6))
        m = torch.empty(n, dtype=torch.uint8, device=dev) if validate else None
        st = _stream(dev)
        with torch.cuda.device(dev):      # the C ABI launches on the CURRENT device
            _lib.check(L.pivb200_predictor(up.data_ptr(), vp.data_ptr(), mp.data_ptr() if validate else None,
                                           1, n0, m0, n1, m1, ay.data_ptr(), ax.data_ptr(), mode,
                                           tmp.data_ptr(), sx.data_ptr(), sy.data_ptr(), base_u.data_ptr(),
                                           base_v.data_ptr(), pred_u.data_ptr(), pred_v.data_ptr(), st))
        h, w = self.frame_shape
        with torch.cuda.device(dev):      # the C ABI launches on the CURRENT device
            _lib.check(L.pivb200_pass_next(fa.data_ptr(), fb.data_ptr(), 1, 0, h, w, fa.stride(0),
                                           self.wind_size, self.overlap, mode, sx.data_ptr(), sy.data_ptr(),
                                           base_u.data_ptr(), base_v.data_ptr(), pred_u.data_ptr(),
                                           pred_v.data_ptr(), 1 if validate else 0, 1.2, u.data_ptr(),
                                           v.data_ptr(), m.data_ptr() if validate else None, None, st))
        val = m.cpu().numpy().astype(bool).reshape(n1, m1) if validate else None
        return (u.cpu().numpy().reshape(n1, m1), v.cpu().numpy().reshape(n1, m1), self.x, self.y, val)


class piv_iteration_CWS(_PivIteration):   # noqa: N801  (reference spelling)
    MODE = "CWS"


class piv_iteration_DWS(_PivIteration):   # noqa: N801
    MODE = "DWS"


class IterModMap:
    functions = {"DWS": piv_iteration_DWS, "CWS": piv_iteration_CWS}


# --------------------------------------------------------------------------------------------
# OfflinePIV (PB:824-903)
# --------------------------------------------------------------------------------------------
class OfflinePIV:
    """Generator-style multipass PIV over a folder of images; same constructor as the reference.

    ``piv = OfflinePIV(...); for x, y, u, v in piv(): ...`` yields, per processed pair, four
    float64 ``[n_rows, n_cols]`` arrays of the LAST pass grid: coordinates in mm
    (``px * scale``) and velocities in m/s (``px * scale / dt * 1000``, dt in microseconds),
    rows flipped and v negated exactly like the reference.  Pairs are skipped when an image
    cannot be read or when the hole filling declines (see postprocess.fill_holes).

    Keyword-only extensions (the reference processes one pair at a time on one device):

    ``batch_pairs``     pairs per kernel launch / H2D copy (default 8)
    ``decode_threads``  image-decoding threads that fill the pinned staging memory (default 4)
    ``shard``           ``(rank, world)``: this object processes only its contiguous block of the pair
                        list (one process per GPU, no collective; see dataset.shard_range).  ``len()``
                        is then the block's length and ``pair_indices`` the global pair numbers.
    ``replace``         ``"reference"`` (default): host hole filling exactly like the reference
                        (SciPy Delaunay); ``"stencil"``: on-device 3x3 replacement, see
                        postprocess_device.py (a documented deviation; never skips a pair);
                        ``"stencil+nmt"``: a normalised median test first widens the invalid set.
    ``fill_workers``    (reference mode) worker PROCESSES for the Delaunay hole filling; 0 (default) fills in the
                        decode threads (Qhull holds the GIL: ~450 pairs/s at 4 MP), N > 0 spreads it over N cores
    After every ``yield`` the attribute ``last_pair_index`` holds the global index of the pair the field belongs
    to (pairs may be skipped, so counting the yields is not enough).

    ``statistics``      (stencil modes only) accumulate the running sums of the reference worker's
                        statistics (workers.py:79-119) on the device; ``statistics_table()`` returns
                        the table after the run."""

    def __init__(self, folder: str, device: str, file_fmt: str, wind_size: int, overlap: int,
                 multipass: int = 1, multipass_mode: str = "CWS", dt: int = 1, scale: float = 1.,
                 multipass_scale: float = 2., folder_mode: str = "pairs", *, batch_pairs: int = 8,
                 decode_threads: int = 4, shard: Optional[Tuple[int, int]] = None,
                 replace: str = "reference", statistics: bool = False, fill_workers: int = 0) -> None:
        self._wind_size = wind_size
        self._overlap = overlap
        self._dt = dt
        self._iter = multipass
        self._iter_scale = multipass_scale
        self._scale = scale
        self._device = DeviceMap.devicies[device]            # KeyError for unknown names
        self._dataset = PIVDataset(folder, file_fmt, folder_mode, transform=None)
        self._iter_function = IterModMap.functions[multipass_mode]   # KeyError for unknown modes
        self._mode = multipass_mode
        if replace not in ("reference", "stencil", "stencil+nmt"):
            raise KeyError(replace)
        if statistics and replace == "reference":
            raise ValueError("device statistics need the fields complete on the device: use replace='stencil'")
        self._replace = replace
        self._want_stats = bool(statistics)
        self.statistics = None
        self._batch_pairs = max(1, int(batch_pairs))
        self._decode_threads = max(1, int(decode_threads))
        self._fill_workers = max(0, int(fill_workers))
        self._fill_pool = None
        self.last_pair_index = None         # global index (into the folder's pair list) of the pair yielded last
        rank, world = shard if shard is not None else (0, 1)
        self.pair_indices = shard_range(len(self._dataset), rank, world)
        self._plan = None
        self._pipe = None
        if not len(self):
            return
        self._device = _cuda_device(self._device)
        frame_a, _ = self._dataset[self.pair_indices.start]
        if frame_a is not None:
            self._plan = self._make_plan(frame_a.shape)

    def _make_plan(self, shape) -> PIVPlan:
        return PIVPlan(shape, self._wind_size, self._overlap, self._iter, self._mode,
                       self._iter_scale, device=self._device)

    def __len__(self) -> int:
        return len(self.pair_indices)

    # -- decoding ------------------------------------------------------------------------------
    def _decode_into(self, path: str, dst: np.ndarray) -> bool:
        """Decode one frame straight into pinned staging memory; False = unreadable / wrong shape."""
        if self._plan is None:
            return False
        return read_gray_into(path, dst)

    def _ensure_plan(self, batches) -> bool:
        """The constructor could not read the first frame: find the first readable one."""
        if self._plan is not None:
            return True
        for batch in batches:
            for path in batch.files:
                img = read_gray(path)
                if img is not None:
                    self._plan = self._make_plan(img.shape)
                    return True
        return False

    def __call__(self) -> Generator:
        """Yield ``(x, y, u, v)`` per processed pair, in pair order.

        Three stages overlap: ``decode_threads`` workers decode the frames of batch i+1 into pinned
        memory (OpenCV releases the GIL) while the GPU runs batch i (one H2D of the unique frames, the
        fused passes, one D2H of u, v, mask) and the caller's thread post-processes batch i-1."""
        from concurrent.futures import ThreadPoolExecutor
        from .engine import FramePipeline
        batches = plan_batches(self._dataset.img_pairs, self._batch_pairs, self.pair_indices)
        if not batches or not self._ensure_plan(batches):
            return
        if self._pipe is None or self._pipe.plan is not self._plan:
            post = None
            if self._replace != "reference":
                from .postprocess_device import FieldStatistics, StencilPost
                g = self._plan.out_geometry
                # every sweep is one kernel launch per batch and fills one more ring of a hole: 24 sweeps close
                # holes up to 48 vectors across (more than a third of a 4 MP field at 16 px spacing -- the
                # reference gives up on such pairs, PB:305-307); vectors left over are zeroed and stay flagged
                post = StencilPost(nmt=self._replace.endswith("nmt"), max_sweeps=min(max(g.n_rows, g.n_cols), 24))
                if self._want_stats:
                    self.statistics = FieldStatistics(g.n_rows, g.n_cols, self._plan.device)
            self._pipe = FramePipeline(self._plan, self._batch_pairs, post=post, stats=self.statistics)
        pipe = self._pipe
        geo = self._plan.out_geometry
        if self._fill_workers and self._replace == "reference" and self._fill_pool is None:
            from .postprocess import HoleFillPool
            self._fill_pool = HoleFillPool(self._fill_workers, self._batch_pairs, geo.n_rows, geo.n_cols)

        with ThreadPoolExecutor(max_workers=self._decode_threads) as pool:
            def stage(n):
                frames = pipe.host_frames(n & 1)
                return [pool.submit(self._decode_into, path, frames[j])
                        for j, path in enumerate(batches[n].files)]

            def finish(n, ok):
                """Results of batch n -> finished fields of its readable pairs, in order.  The per-pair host
                post-processing (SciPy Delaunay fill in the reference mode) runs on the worker pool too."""
                u, v, bad = pipe.result(n & 1)
                batch = batches[n]
                jobs = [(batch.first_pair + i, pool.submit(self._finalize, u[i].copy(), v[i].copy(), geo, bad[i]))
                        for i in range(len(batch)) if ok[batch.index_a[i]] and ok[batch.index_b[i]]]
                for pair_index, job in jobs:
                    out = job.result()
                    if out is not None:
                        # pairs can be skipped (unreadable frame, hole filling declined): tell which one this is
                        self.last_pair_index = pair_index
                        yield out

            if self._replace != "reference":
                # stencil modes: the holes were filled on the device; what is left (flip, sign, units: PB:894-900) is
                # done for the whole batch at once -- per-pair NumPy calls in the (GIL-sharing) decode threads were
                # the bottleneck of the from-files path
                def finish(n, ok):          # noqa: F811 - replaces the in-thread version above
                    u, v, bad = pipe.result(n & 1)
                    batch = batches[n]
                    # same elementwise operation order as the reference (PB:896-897)
                    U = np.flip(u[:len(batch)], axis=1) * self._scale / self._dt * 1000
                    V = -np.flip(v[:len(batch)], axis=1) * self._scale / self._dt * 1000
                    for i in range(len(batch)):
                        if ok[batch.index_a[i]] and ok[batch.index_b[i]]:
                            self.last_pair_index = batch.first_pair + i
                            yield geo.x * self._scale, geo.y * self._scale, U[i], V[i]

            if self._fill_pool is not None:
                # reference mode with worker processes: a batch's host post-processing is ONE task of a worker
                # process; up to 2 x workers batches are in flight, results are yielded in pair order
                from collections import deque
                waiting = deque()

                def finish(n, ok):          # noqa: F811 - replaces the in-thread version above
                    u, v, bad = pipe.result(n & 1)
                    batch = batches[n]
                    sel = [i for i in range(len(batch)) if ok[batch.index_a[i]] and ok[batch.index_b[i]]]
                    if sel:
                        waiting.append(([batch.first_pair + i for i in sel],
                                        self._fill_pool.submit_batch(u[sel], v[sel], bad[sel], self._scale, self._dt)))
                    yield from drain(self._fill_pool.capacity - 1)

                def drain(limit):
                    while len(waiting) > limit:
                        ids, handle = waiting.popleft()
                        for pair_index, out in zip(ids, self._fill_pool.collect(handle)):
                            if out is not None:
                                self.last_pair_index = pair_index
                                yield geo.x * self._scale, geo.y * self._scale, out[0], out[1]
            else:
                def drain(limit):
                    return iter(())

            decoding = stage(0)
            previous = None                       # (batch number, ok flags) in flight on the GPU
            for n, batch in enumerate(batches):
                ok = [f.result() for f in decoding]
                keep = [ok[batch.index_a[i]] and ok[batch.index_b[i]] for i in range(len(batch))]
                pipe.submit(n & 1, len(batch), batch.chained, keep=keep)
                if n + 1 < len(batches):
                    # slot (n + 1) & 1 was used by batch n - 1, whose upload has long finished
                    decoding = stage(n + 1)
                if previous is not None:
                    yield from finish(*previous)
                previous = (n, ok)
            yield from finish(*previous)
            yield from drain(0)

    def statistics_table(self) -> dict:
        """The reference worker's statistics table (workers.py:100-119) of the pairs processed so far."""
        if self.statistics is None:
            raise ValueError("construct OfflinePIV with replace='stencil', statistics=True")
        g = self._plan.out_geometry
        return self.statistics.table(g.x, g.y, self._scale, self._dt)

    def _finalize(self, u, v, geo, invalid):
        if self._replace != "reference":
            from .postprocess_device import finalize_field_stencil
            return finalize_field_stencil(u, v, geo.x, geo.y, invalid, self._scale, self._dt)
        return finalize_field(u, v, geo.x, geo.y, invalid, self._scale, self._dt)

    def close(self) -> None:
        """Stop the hole-filling worker processes (``fill_workers`` > 0)."""
        if self._fill_pool is not None:
            self._fill_pool.shutdown()
            self._fill_pool = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
