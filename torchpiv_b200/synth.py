"""Synthetic particle-image pairs (host side, NumPy only).

Used by ``bench.py`` and the tests to build the workloads BASELINE.json names
(uniform-shift and Rankine-vortex particle pairs) -- there are no datasets offline and
the reference's ``test_images`` are stripped from the checkout.  Not part of the hot path.

Image model: Gaussian particles, sigma 1 px, 0.03 particles/px,
amplitude U(0.5, 1) * 200 over a background of 10, clipped to uint8.  Every image can carry
one uniform-noise patch in frame *b* and one constant patch in both frames, so that some
interrogation windows are invalid (the reference silently skips pairs without any
invalid vector, PB:299-304 / 891-892) and exact-tie windows are exercised.
"""
from __future__ import annotations

import os
import struct
import numpy as np

__all__ = ["uniform_shift", "rankine_vortex", "particle_pair", "write_bmp", "write_pair_folder"]


def uniform_shift(dx: float, dy: float):
    def field(px, py):
        return np.full_like(px, dx), np.full_like(py, dy)
    return field


def rankine_vortex(cx: float, cy: float, core_radius: float, peak_disp: float):
    """Solid-body core, 1/r outside; ``peak_disp`` px tangential displacement at the core edge."""
    def field(px, py):
        rx, ry = px - cx, py - cy
        r = np.hypot(rx, ry)
        r_safe = np.maximum(r, 1e-9)
        vt = np.where(r < core_radius, peak_disp * r / core_radius,
                      peak_disp * core_radius / r_safe)
        return -vt * ry / r_safe, vt * rx / r_safe
    return field


def _render(shape, px, py, amp, sigma, radius=4):
    h, w = shape
    ix = np.rint(px).astype(np.int64)
    iy = np.rint(py).astype(np.int64)
    img = np.zeros(h * w, dtype=np.float64)
    inv = 1.0 / (2.0 * sigma * sigma)
    for oy in range(-radius, radius + 1):
        yy = iy + oy
        wy = np.exp(-((yy - py) ** 2) * inv)
        for ox in range(-radius, radius + 1):
            xx = ix + ox
            ok = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
            wgt = amp * wy * np.exp(-((xx - px) ** 2) * inv)
            img += np.bincount((yy[ok] * w + xx[ok]), weights=wgt[ok], minlength=h * w)
    return img.reshape(h, w)


def particle_pair(shape, field, seed: int, density: float = 0.03, sigma: float = 1.0,
                  background: float = 10.0, noise_patch=None, blank_patch=None,
                  margin: int = 24):
    """Return (frame_a, frame_b) uint8.  ``field(px, py) -> (dx, dy)`` displaces the particles
    of frame *a* to make frame *b*.  Patches are ``(r0, r1, c0, c1)`` or None."""
    h, w = shape
    rng = np.random.default_rng(seed)
    n = int(round(density * (h + 2 * margin) * (w + 2 * margin)))
    px = rng.uniform(-margin, w + margin, n)
    py = rng.uniform(-margin, h + margin, n)
    amp = rng.uniform(0.5, 1.0, n) * 200.0
    dx, dy = field(px, py)
    a = _render(shape, px, py, amp, sigma) + background
    b = _render(shape, px + dx, py + dy, amp, sigma) + background
    a = np.clip(np.rint(a), 0, 255).astype(np.uint8)
    b = np.clip(np.rint(b), 0, 255).astype(np.uint8)
    if noise_patch is not None:
        r0, r1, c0, c1 = noise_patch
        b[r0:r1, c0:c1] = rng.integers(0, 256, (r1 - r0, c1 - c0), dtype=np.uint8)
    if blank_patch is not None:
        r0, r1, c0, c1 = blank_patch
        a[r0:r1, c0:c1] = int(background)
        b[r0:r1, c0:c1] = int(background)
    return a, b


def default_patches(shape):
    """The patch placement of SURVEY.md 8d, scaled with the frame (2048^2 -> the probe's boxes)."""
    h, w = shape
    s_r, s_c = h / 2048.0, w / 2048.0
    noise = (int(300 * s_r), int(500 * s_r), int(900 * s_c), int(1200 * s_c))
    blank = (int(1500 * s_r), int(1700 * s_r), int(200 * s_c), int(420 * s_c))
    return noise, blank


def write_bmp(path: str, img: np.ndarray) -> None:
    """8-bit palettised grayscale BMP (bottom-up rows, 4-byte row padding)."""
    h, w = img.shape
    pad = (-w) % 4
    rows = np.zeros((h, w + pad), dtype=np.uint8)
    rows[:, :w] = img[::-1]
    palette = b"".join(struct.pack("<BBBB", i, i, i, 0) for i in range(256))
    off = 14 + 40 + len(palette)
    size = off + rows.size
    header = struct.pack("<2sIHHI", b"BM", size, 0, 0, off)
    info = struct.pack("<IiiHHIIiiII", 40, w, h, 1, 8, 0, rows.size, 2835, 2835, 256, 0)
    with open(path, "wb") as fh:
        fh.write(header + info + palette + rows.tobytes())


def write_pair_folder(folder: str, pairs, start: int = 1300, fmt: str = "bmp") -> list:
    """Write ``image<NNNN>_a.bmp`` / ``_b.bmp`` like the reference's bundled example set."""
    os.makedirs(folder, exist_ok=True)
    names = []
    for i, (a, b) in enumerate(pairs):
        for tag, img in (("a", a), ("b", b)):
            name = os.path.join(folder, f"image{start + i}_{tag}.{fmt}")
            write_bmp(name, img)
            names.append(name)
    return names
