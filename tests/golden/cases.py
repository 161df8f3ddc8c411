"""Deterministic inputs shared by the golden-vector generator and the tests.

Everything here is regenerated from seeds, so the committed fixtures hold only the
reference's OUTPUTS plus a SHA-256 of the inputs they were computed from."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from torchpiv_b200 import synth  # noqa: E402

GOLDEN_DIR = os.path.dirname(os.path.abspath(__file__))
SMALL_SHAPE = (288, 352)
PASS1_GEOMS = [(64, 32), (32, 16), (16, 8), (32, 8), (64, 48)]
GENERAL_GEOMS = [(48, 24), (24, 12), (128, 64), (20, 6), (42, 21), (96, 48), (160, 80), (192, 64), (256, 200),
                 (25, 12), (35, 17), (31, 15), (63, 31)]   # not 16/32/64: general kernel; odd sizes: [w, w-1] maps


def sha(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode() + str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def small_pair(seed: int = 1, kind: str = "uniform", zero_patch: bool = False):
    """288x352 pair with a noise patch (frame b), a constant patch (both) and, optionally,
    an all-black patch (both; pass-1 mean normalisation divides by zero there)."""
    shape = SMALL_SHAPE
    if kind == "uniform":
        field = synth.uniform_shift(3.3, -2.2)
    elif kind == "vortex":
        field = synth.rankine_vortex(shape[1] / 2, shape[0] / 2, 60.0, 5.0)
    elif kind == "zero":
        field = synth.uniform_shift(0.0, 0.0)
    else:
        raise ValueError(kind)
    a, b = synth.particle_pair(shape, field, seed=seed, noise_patch=(20, 100, 200, 300),
                               blank_patch=(160, 260, 40, 150))
    if zero_patch:
        a[100:200, 230:340] = 0
        b[100:200, 230:340] = 0
    return a, b


def shift_case(seed: int = 0, w: int = 32, ovl: int = 16):
    """Per-window shifts that leave the frame, hit exact integers and sit one float32 ulp
    below an integer (the reference adds them to the absolute pixel coordinate in float32)."""
    a, _ = small_pair(seed=3)
    h, wf = a.shape
    n_r = (h - w) // (w - ovl) + 1
    n_c = (wf - w) // (w - ovl) + 1
    n = n_r * n_c
    rng = np.random.default_rng(seed)
    vx = rng.uniform(-40, 40, n).astype(np.float32)
    vy = rng.uniform(-40, 40, n).astype(np.float32)
    vx[::7] = np.rint(vx[::7])
    vy[::5] = np.rint(vy[::5])
    vx[3::11] = np.nextafter(np.rint(vx[3::11]), np.float32(-100)).astype(np.float32)
    vy[4::13] = np.nextafter(np.rint(vy[4::13]), np.float32(100)).astype(np.float32)
    vx[0], vy[0] = 0.0, 0.0
    vx[1], vy[1] = -400.0, -400.0      # far outside: every tap clamps to flat index 0
    vx[2], vy[2] = 400.0, 400.0        # ... or to the last pixel
    return a, vx, vy


def adversarial_maps(seed: int, c: int, w: int, dtype):
    """Positive random maps with a planted main peak (corners, row ends, interior) and a
    planted second peak inside / on the edge of / outside the 7x7 flat-index patch."""
    rng = np.random.default_rng(seed)
    maps = rng.uniform(0.0, 1.0, (c, w, w))
    n2 = w * w
    special = [0, 1, w - 1, w, n2 - 1, n2 - 2, n2 - w, n2 - w - 1, w * (w // 2) + w // 2,
               w * 3 + w - 1, w * 4, n2 - w + 1, 2 * w - 2, 3 + 3 * w, 4 + 3 * w, n2 - 4 - 3 * w]
    flat = maps.reshape(c, n2)
    for i in range(c):
        m = special[i % len(special)] if i % 3 else int(rng.integers(0, n2))
        flat[i, m] = rng.uniform(4.0, 8.0)
        for nb in (m - 1, m + 1, m - w, m + w):      # a plausible peak shape
            if 0 <= nb < n2:
                flat[i, nb] = max(flat[i, nb], flat[i, m] * rng.uniform(0.3, 0.8))
        off_i, off_j = int(rng.integers(-5, 6)), int(rng.integers(-5, 6))
        m2 = min(max(m + off_i + w * off_j, 0), n2 - 1)
        if m2 != m and m2 not in (m - 1, m + 1, m - w, m + w):
            ratio = (1.05, 1.15, 1.3, 1.6)[i % 4]
            flat[i, m2] = flat[i, m] / ratio
    if seed % 2:
        flat[0, :] = 0.5           # featureless map: every value ties
    maps -= maps.min(axis=(1, 2), keepdims=True)     # what the callers do before (PB:518)
    return maps.astype(dtype)
