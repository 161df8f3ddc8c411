"""Result files in the reference's formats.

``save_table`` / ``save_binary`` write what the reference's worker writes per pair and for the
statistics table (PlotterFunctions.py:16-24, 48-65; called from workers.py:64-70, 122): a dict of
equally shaped arrays becomes either a ``%.6f`` CSV with a ``", "``-joined header line (columns =
dict order, arrays flattened row-major) or one ``.npy`` holding the arrays stacked along axis 0.
Existing files are never overwritten: ``name (1).ext``, ``name (2).ext`` ... like the reference.
Pinned byte for byte by ``tests/golden/output_formats.npz``."""
from __future__ import annotations

import os
from typing import Dict

import numpy as np

__all__ = ["uniquify", "save_table", "save_binary", "pair_output", "PairWriter"]


def uniquify(path: str) -> str:
    """First of ``path``, ``stem (1).ext``, ``stem (2).ext`` ... that does not exist yet."""
    stem, ext = os.path.splitext(path)
    candidate, n = path, 0
    while os.path.exists(candidate):
        n += 1
        candidate = f"{stem} ({n}){ext}"
    return candidate


def _target(name: str, path: str) -> str:
    if not os.path.exists(path):
        os.mkdir(path)              # like the reference: one level only, a missing parent raises
    return uniquify(os.path.join(path, name))


def save_binary(name: str, path: str, data: Dict[str, np.ndarray], sep: str = ", ") -> str:
    """``np.save`` of the dict's arrays stacked on a new leading axis (``np.save`` appends ``.npy``
    when the name lacks it, after the uniqueness check -- same as the reference).  Returns the path
    handed to ``np.save``."""
    target = _target(name, path)
    np.save(target, np.stack(list(data.values()), axis=0))
    return target


def save_table(name: str, path: str, data: Dict[str, np.ndarray], sep: str = ", ") -> str:
    """Text table: header ``sep.join(keys)``, one row per vector, ``%.6f``.  Unlike the reference the
    caller's dict is left untouched (the reference flattens its arrays in place, which is why its
    callers pass copies)."""
    columns = np.stack([np.asarray(v).reshape(-1) for v in data.values()], axis=1)
    target = _target(name, path)
    np.savetxt(target, columns, delimiter=sep, header=sep.join(data.keys()), comments="", fmt="%.6f")
    return target


def pair_output(x, y, u, v) -> Dict[str, np.ndarray]:
    """The per-pair dict of workers.py:58-63 (keys and order)."""
    return {"x[mm]": x, "y[mm]": y, "Vx[m/s]": u, "Vy[m/s]": v}


class PairWriter:
    """Per-pair saving as the reference's worker does it (workers.py:64-70): ``<folder name>_pair.npy``
    or ``_pair.txt`` in ``save_dir``, numbered by :func:`uniquify`; ``statistics`` writes
    ``<folder name>_statistics.txt`` (workers.py:120-122)."""

    MODES = ("Save all binary", "Save all text", "Dont save")

    def __init__(self, folder: str, save_dir: str, save_opt: str = "Save all binary"):
        if save_opt not in self.MODES:
            raise KeyError(save_opt)
        self.name = os.path.basename(os.path.normpath(folder))
        self.save_dir, self.save_opt = save_dir, save_opt

    def pair(self, x, y, u, v):
        if self.save_opt == "Save all binary":
            return save_binary(f"{self.name}_pair.npy", self.save_dir, pair_output(x, y, u, v))
        if self.save_opt == "Save all text":
            return save_table(f"{self.name}_pair.txt", self.save_dir, pair_output(x, y, u, v))
        return None

    def statistics(self, table: Dict[str, np.ndarray]):
        if self.save_opt == "Dont save":
            return None
        return save_table(f"{self.name}_statistics.txt", self.save_dir, table)
