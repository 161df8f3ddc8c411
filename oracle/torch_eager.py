"""The oracle's 2-pass chain restated with stock eager PyTorch ops  --  BENCH / TEST INFRASTRUCTURE ONLY.

Purpose: BASELINE.json's north_star sets the bar "at least the reference's torch-CUDA path on the same
B200".  The reference itself cannot travel to the GPU box (its sources must not be copied and
``/root/reference`` does not exist there), so ``bench.py`` times THIS restatement with ``device="cuda"``:
the same algorithm as ``oracle/piv_oracle.py`` expressed with the library operations an eager-PyTorch
implementation consists of (``as_strided`` window stacks, ``torch.fft.rfft2/irfft2``, elementwise index
arithmetic + gathers for the window shift, ``argmax``, one small scatter per patch element, three
device-to-host copies per pass, SciPy splines on the host between passes).  It says what stock
aten / cuFFT kernels achieve on this GPU for this workload; it is not the product and the product never
imports it.

Parity status: PINNED on the CPU -- ``tests/test_oracle_golden.py::test_torch_eager_*`` compares it (with
``device="cpu"``) against the golden vectors of the unmodified reference.
"""
from __future__ import annotations

import numpy as np
import torch

from . import piv_oracle as O


def windows(frame: torch.Tensor, w: int, o: int) -> torch.Tensor:
    h, wf = frame.shape
    step = w - o
    n_r, n_c = (h - w) // step + 1, (wf - w) // step + 1
    s0, s1 = frame.stride()
    return torch.as_strided(frame, (n_r, n_c, w, w), (s0 * step, s1 * step, s0, s1)).reshape(-1, w, w)


def correlate(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    spec = torch.conj(torch.fft.rfft2(a)) * torch.fft.rfft2(b)
    return torch.fft.fftshift(torch.fft.irfft2(spec, s=a.shape[-2:]), dim=(-2, -1))


def _second_peak(flat: torch.Tensor, m: torch.Tensor, k: int, wind: int) -> torch.Tensor:
    n2 = flat.shape[1]
    zero = torch.zeros((flat.shape[0], 1), dtype=flat.dtype, device=flat.device)
    for i in range(-wind, wind + 1):
        for j in range(-wind, wind + 1):
            flat.scatter_(1, (m + i + k * j).clamp_(0, n2 - 1), zero)
    return flat.argmax(dim=1, keepdim=True)


def corr_to_disp(corr: torch.Tensor, n_rows: int, n_cols: int, validate: bool = True, val_ratio: float = 1.2,
                 wind: int = 3):
    c, d, k = corr.shape
    n2 = d * k
    corr += O.EPS
    flat = corr.view(c, n2)
    cor = flat.to(torch.float64)
    m = flat.argmax(dim=1, keepdim=True)
    left, right, top, bot = m + 1, m - 1, m + k, m - k
    left = torch.where(left >= n2 - 1, m, left)
    right = torch.where(right <= 0, m, right)
    top = torch.where(top >= n2 - 1, m, top)
    bot = torch.where(bot <= 0, m, bot)
    lm, ll, lr = (torch.log(torch.gather(cor, 1, i)) for i in (m, left, right))
    lt, lb = (torch.log(torch.gather(cor, 1, i)) for i in (top, bot))
    v = (m // d) + (lb - lt) / (2 * (lb + lt) - 4 * lm) - int(d / 2)
    u = (m % k) + (lr - ll) / (2 * (ll + lr) - 4 * lm) - int(k / 2)
    mask = None
    if validate:
        cm = torch.gather(cor, 1, m)
        m2 = _second_peak(flat, m, k, wind)
        mask = ((cm / torch.gather(cor, 1, m2)) < val_ratio).reshape(n_rows, n_cols).cpu().numpy()
    u = torch.nan_to_num(u).reshape(n_rows, n_cols).cpu().numpy()
    v = torch.nan_to_num(v).reshape(n_rows, n_cols).cpu().numpy()
    return u, v, mask


def pass_first(fa: torch.Tensor, fb: torch.Tensor, w: int, o: int):
    n_rows, n_cols = O.get_field_shape(fa.shape, w, o)
    x, y = O.get_coordinates(fa.shape, w, o)
    aa, bb = windows(fa, w, o), windows(fb, w, o)
    aa = aa / aa.mean(dim=(-2, -1), dtype=torch.float64, keepdim=True)
    bb = bb / bb.mean(dim=(-2, -1), dtype=torch.float64, keepdim=True)
    corr = correlate(aa, bb)
    corr = corr - corr.amin(dim=(-2, -1), keepdim=True)
    u, v, mask = corr_to_disp(corr, n_rows, n_cols, True)
    return u, v, x, y, mask


def _bilinear(frame: torch.Tensor, grid: torch.Tensor, vx: torch.Tensor, vy: torch.Tensor) -> torch.Tensor:
    wf, n = frame.shape[-1], frame.numel()
    new_y = (grid // wf) + vy                    # int64 + float32 -> float32
    new_x = (grid % wf) + vx
    up_x, up_y = torch.ceil(new_x).long(), torch.ceil(new_y).long()
    dn_x, dn_y = torch.floor(new_x).long(), torch.floor(new_y).long()
    flat = frame.reshape(-1)

    def tap(yy, xx):
        return flat[(yy * wf + xx).clamp_(0, n - 1)]
    q11, q12, q21, q22 = tap(dn_y, dn_x), tap(up_y, dn_x), tap(dn_y, up_x), tap(up_y, up_x)
    wx1, wx0 = up_x - new_x, new_x - dn_x
    wy1, wy0 = up_y - new_y, new_y - dn_y
    out = q11 * wx1 * wy1 + q21 * wx0 * wy1 + q12 * wx1 * wy0 + q22 * wx0 * wy0
    exact = (up_x - dn_x) * (up_y - dn_y) == 0
    out[exact] = q11[exact].to(out.dtype)
    return out


class IterCWS:
    """One later CWS pass: host splines, device window shift + correlation, three D2H copies."""

    def __init__(self, frame_shape, w: int, o: int, device):
        self.w, self.o, self.device = w, o, device
        self.n_rows, self.n_cols = O.get_field_shape(frame_shape, w, o)
        self.x, self.y = O.get_coordinates(frame_shape, w, o)
        h, wf = frame_shape
        self.idx = windows(torch.arange(h * wf, dtype=torch.int64, device=device).reshape(h, wf), w, o).contiguous()

    def __call__(self, fa, fb, x0, y0, u0, v0, mask):
        sy, sx = self.y[:, 0], self.x[0, :]
        u0 = O.resample_predictor(x0, y0, u0, sy, sx)
        v0 = O.resample_predictor(x0, y0, v0, sy, sx)
        u2, v2 = u0 / 2, v0 / 2
        if mask is not None:
            bad = O.resample_predictor(x0, y0, mask, sy, sx) >= .5
            u0[bad] = 0.0
            v0[bad] = 0.0
        u2t = torch.tensor(u2, dtype=torch.float32, device=self.device).reshape(-1, 1, 1)
        v2t = torch.tensor(v2, dtype=torch.float32, device=self.device).reshape(-1, 1, 1)
        aa = _bilinear(fa, self.idx, -u2t, -v2t)
        bb = _bilinear(fb, self.idx, u2t, v2t)
        corr = correlate(aa, bb)
        corr = corr - corr.amin(dim=(-2, -1), keepdim=True)
        du, dv, val = corr_to_disp(corr, self.n_rows, self.n_cols, mask is not None)
        u, v = 2 * u2 + du, 2 * v2 + dv
        mu, mv = (du > u0) * (np.rint(u0) > 0), (dv > v0) * (np.rint(v0) > 0)
        if val is not None:
            mu[val] = True
            mv[val] = True
        u[mu], v[mv] = u0[mu], v0[mv]
        return u, v, self.x, self.y, val


def two_pass_cws(fa: torch.Tensor, fb: torch.Tensor, w: int = 64, o: int = 32, second: IterCWS = None):
    """Pass 1 at (w, o) and one CWS pass at (w/2, o/2) on uint8 frames already on their device."""
    u, v, x, y, mask = pass_first(fa, fb, w, o)
    if second is None:
        second = IterCWS(tuple(fa.shape), w // 2, o // 2, fa.device)
    return second(fa, fb, x, y, u, v, mask)
