"""Image-pair enumeration and decoding (host side).

Same file semantics as the reference's ``PIVDataset`` / ``natural_keys`` (PIVbackend.py:114-144,
PlotterFunctions.py:27-37): files of the folder whose name ends with ``file_fmt``, natural
(human) sort, ``pairs`` -> (0,1),(2,3),...  ``sequential`` -> (i, i+1); 8-bit grayscale decode
with OpenCV; an unreadable file yields ``(None, None)`` and the pair is skipped by the caller."""
from __future__ import annotations

import os
import re
from typing import List, Tuple

import numpy as np

__all__ = ["natural_keys", "list_pairs", "read_gray", "read_gray_into", "PIVDataset", "ToTensor", "shard_range",
           "FrameBatch", "plan_batches"]

_DIGITS = re.compile(r"(\d+)")


def natural_keys(text: str):
    """Sort key that orders embedded integers numerically ("img2" < "img10")."""
    return [int(tok) if tok.isdigit() else tok for tok in _DIGITS.split(text)]


def list_pairs(folder: str, file_fmt: str, folder_mode: str) -> List[Tuple[str, str]]:
    names = sorted((os.path.join(folder, n) for n in os.listdir(folder) if n.endswith(file_fmt)),
                   key=natural_keys)
    if folder_mode == "pairs":
        return list(zip(names[::2], names[1::2]))
    if folder_mode == "sequential":
        return list(zip(names[:-1], names[1:]))
    return []


def read_gray(path: str):
    """uint8 [H, W] array or None.  np.fromfile + imdecode handles non-ASCII paths."""
    import cv2
    try:
        raw = np.fromfile(path, dtype=np.uint8)
    except OSError:
        return None
    if raw.size == 0:
        return None
    return cv2.imdecode(raw, cv2.IMREAD_GRAYSCALE)


_GRAY_PALETTE = bytes(b for i in range(256) for b in (i, i, i, 0))


def _bmp8_view(raw: np.ndarray):
    """View of the pixel rows (top-down) of an UNCOMPRESSED 8-bit BMP whose palette is the identity grey
    ramp -- the format PIV cameras and the reference's example set use -- or None for anything else.
    For such a file ``cv2.imdecode(..., IMREAD_GRAYSCALE)`` returns exactly these bytes, so the generic
    decoder (palette lookup + colour conversion, several ms per 4 MP frame) can be skipped."""
    if raw.size < 54 + 1024 or raw[0] != 0x42 or raw[1] != 0x4D:
        return None
    hdr = raw[:54].tobytes()
    off = int.from_bytes(hdr[10:14], "little")
    if int.from_bytes(hdr[14:18], "little") != 40:                       # BITMAPINFOHEADER only
        return None
    w = int.from_bytes(hdr[18:22], "little", signed=True)
    h = int.from_bytes(hdr[22:26], "little", signed=True)
    planes, bpp = int.from_bytes(hdr[26:28], "little"), int.from_bytes(hdr[28:30], "little")
    compression, used = int.from_bytes(hdr[30:34], "little"), int.from_bytes(hdr[46:50], "little")
    if planes != 1 or bpp != 8 or compression != 0 or used not in (0, 256) or w <= 0 or h == 0:
        return None
    if off != 54 + 1024 or raw[54:54 + 1024].tobytes() != _GRAY_PALETTE:
        return None
    stride = (w + 3) & ~3
    rows = abs(h)
    if raw.size < off + stride * rows:
        return None
    px = raw[off:off + stride * rows].reshape(rows, stride)[:, :w]
    return px[::-1] if h > 0 else px                                     # positive height = bottom-up


def read_gray_into(path: str, dst: np.ndarray) -> bool:
    """Decode ``path`` straight into ``dst`` (uint8 ``[H, W]``, e.g. pinned staging memory).  False when
    the file is unreadable or its shape differs (the caller skips the pair)."""
    try:
        raw = np.fromfile(path, dtype=np.uint8)
    except OSError:
        return False
    if raw.size == 0:
        return False
    img = _bmp8_view(raw)
    if img is None:
        import cv2
        img = cv2.imdecode(raw, cv2.IMREAD_GRAYSCALE)
        if img is None:
            return False
    if img.shape != dst.shape:
        print(f"Warning! {path}: frame shape {img.shape} != {dst.shape}, pair skipped")
        return False
    np.copyto(dst, img)
    return True


class ToTensor:
    """numpy -> torch tensor of a fixed dtype (None passes through)."""

    def __init__(self, dtype) -> None:
        self.dtype = dtype

    def __call__(self, data):
        if data is None:
            return None
        import torch
        return torch.tensor(data, dtype=self.dtype)


class PIVDataset:
    def __init__(self, folder, file_fmt, folder_mode, transform=None):
        self.transform = transform
        self.img_pairs = list_pairs(folder, file_fmt, folder_mode)

    def __len__(self):
        return len(self.img_pairs)

    def __getitem__(self, index):
        name_a, name_b = self.img_pairs[int(index)]
        frame_b = read_gray(name_b)
        frame_a = read_gray(name_a)
        if frame_a is None or frame_b is None:
            return None, None
        if self.transform:
            return self.transform(frame_a), self.transform(frame_b)
        return frame_a, frame_b


# --------------------------------------------------------------------------------------------
# Batching and sharding of the pair list (no reference counterpart:
# the reference decodes and processes one pair at a time on one device, PIVbackend.py:862-903)
# --------------------------------------------------------------------------------------------
def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous block of ``range(n_items)`` owned by ``rank`` of ``world`` (block sizes differ by
    at most one).  Contiguous blocks keep consecutive pairs on one GPU, so ``sequential`` folders
    decode and upload every frame once per shard instead of twice."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(int(n_items), world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


class FrameBatch:
    """One batch of pairs as the frames to decode and where each pair finds its two frames.

    ``files``   unique frame files of the batch, in upload order
    ``index_a`` / ``index_b``  position in ``files`` of frame a / b of every pair
    ``chained`` True when pair i's frame b is pair i+1's frame a for the whole batch (sequential
                folders): the frames are then uploaded ONCE as ``[K+1, H, W]`` and the two frame
                stacks the kernels read are the overlapping views ``[0:K]`` and ``[1:K+1]``."""

    def __init__(self, first_pair: int, pairs):
        self.first_pair = int(first_pair)
        self.pairs = list(pairs)
        k = len(self.pairs)
        self.chained = k > 0 and all(self.pairs[i][1] == self.pairs[i + 1][0] for i in range(k - 1))
        if self.chained:
            self.files = [p[0] for p in self.pairs] + [self.pairs[-1][1]]
            self.index_a = list(range(k))
            self.index_b = list(range(1, k + 1))
        else:
            # frames a first, then frames b: both stacks are contiguous blocks of the upload
            self.files = [p[0] for p in self.pairs] + [p[1] for p in self.pairs]
            self.index_a = list(range(k))
            self.index_b = list(range(k, 2 * k))

    def __len__(self) -> int:
        return len(self.pairs)


def plan_batches(pairs, batch_pairs: int, indices=None):
    """Split the pair list (or the sub-range ``indices`` of it, e.g. a :func:`shard_range`) into
    :class:`FrameBatch` es of at most ``batch_pairs`` consecutive pairs."""
    if batch_pairs < 1:
        raise ValueError("batch_pairs must be >= 1")
    idx = range(len(pairs)) if indices is None else indices
    out = []
    for s in range(idx.start, idx.stop, batch_pairs):
        e = min(s + batch_pairs, idx.stop)
        out.append(FrameBatch(s, [pairs[i] for i in range(s, e)]))
    return out
