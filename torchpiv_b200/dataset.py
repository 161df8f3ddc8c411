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

__all__ = ["natural_keys", "list_pairs", "read_gray", "PIVDataset", "ToTensor"]

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
