"""ctypes binding of libpivb200.so (C ABI declared in include/pivb200.h).

The shared library is the product: if it is missing, loading fails loudly -- there is no
CPU or PyTorch fallback for any entry point."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_int, c_longlong, c_void_p, POINTER

_HERE = os.path.dirname(os.path.abspath(__file__))
# PIVB200_LIB selects another build of the same library (A/B experiments); default: the in-tree build
LIB_PATH = os.environ.get("PIVB200_LIB") or os.path.join(_HERE, "libpivb200.so")

E_WINDOW, E_OVERLAP, E_FRAME, E_ARG, E_DRIVER, E_SIZE = -1, -2, -3, -4, -5, -6
MODE_DWS, MODE_CWS = 0, 1
MODES = {"DWS": MODE_DWS, "CWS": MODE_CWS}

_lib = None

# name -> (restype, argtypes); every symbol of include/pivb200.h
SIGNATURES = {
    "pivb200_version": (c_int, []),
    "pivb200_error_string": (c_char_p, [c_int]),
    "pivb200_field_shape": (c_int, [c_int, c_int, c_int, c_int, POINTER(c_int), POINTER(c_int)]),
    "pivb200_pass_first": (c_int, [c_void_p, c_void_p, c_int, c_longlong, c_int, c_int, c_int, c_int,
                                   c_int, c_int, c_double, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p]),
    "pivb200_pass_next": (c_int, [c_void_p, c_void_p, c_int, c_longlong, c_int, c_int, c_int, c_int,
                                  c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_int, c_double, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p]),
    "pivb200_predictor": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                  c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p]),
    "pivb200_correlate": (c_int, [c_void_p, c_void_p, c_int, c_longlong, c_int, c_void_p, c_void_p]),
    "pivb200_corr_to_disp": (c_int, [c_void_p, c_int, c_longlong, c_int, c_int, c_int, c_double, c_int,
                                     c_void_p, c_void_p, c_void_p, c_void_p]),
    "pivb200_windows": (c_int, [c_void_p, c_void_p, c_int, c_longlong, c_int, c_int, c_int, c_int,
                                c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pivb200_bilinear_cws": (c_int, [c_void_p, c_int, c_int, c_void_p, c_longlong, c_int, c_void_p,
                                     c_void_p, c_void_p, c_void_p]),
    "pivb200_shift_dws": (c_int, [c_void_p, c_int, c_int, c_void_p, c_longlong, c_int, c_void_p,
                                  c_void_p, c_void_p, c_void_p]),
    "pivb200_nmt": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_double, c_void_p,
                            c_void_p]),
    "pivb200_replace_workspace_bytes": (c_longlong, [c_int, c_int, c_int]),
    "pivb200_replace": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "pivb200_stats_accumulate": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_longlong, c_void_p, c_void_p]),
    "pivb200_measure_fp32_peak": (c_int, [c_int, POINTER(c_double), c_void_p]),
    "pivb200_launch_count": (c_longlong, []),
}


def lib() -> ctypes.CDLL:
    """Load (once) and return the library with typed prototypes."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is not built. Build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (or `make -C torchpiv_b200/csrc -j4`). torchpiv_b200 has no CPU or "
                "PyTorch fallback: the hand-written sm_100a kernels are the only compute path.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def error_string(code: int) -> str:
    return lib().pivb200_error_string(int(code)).decode()


def check(code: int) -> None:
    """Raise like the reference does: geometry problems are ValueError (PB:503-507)."""
    if code == 0:
        return
    msg = error_string(code)
    if code in (E_WINDOW, E_OVERLAP, E_FRAME):
        raise ValueError(msg)
    raise RuntimeError(f"pivb200 error {code}: {msg}")


def launch_count() -> int:
    return int(lib().pivb200_launch_count())
