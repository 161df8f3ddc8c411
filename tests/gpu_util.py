"""Helpers for the GPU parity tests: thin wrappers that call the C ABI (through ctypes) on
torch-allocated device buffers.  No numerics live here."""
import numpy as np
import torch

from torchpiv_b200 import _lib

DEV = torch.device("cuda", 0)


def dev_u8(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def stream():
    return torch.cuda.current_stream(DEV).cuda_stream


def field_shape(shape, w, o):
    return (shape[0] - w) // (w - o) + 1, (shape[1] - w) // (w - o) + 1


def windows(a, b, w, o, mode="DWS", sx=None, sy=None):
    """pivb200_windows on [H,W] (or [B,H,W]) uint8 numpy frames -> float32 numpy [N,w,w] x2."""
    L = _lib.lib()
    fa, fb = dev_u8(a), dev_u8(b)
    if fa.dim() == 2:
        fa, fb = fa[None], fb[None]
    B, H, W = fa.shape
    nr, nc = field_shape((H, W), w, o)
    n = B * nr * nc
    wa = torch.empty((n, w, w), dtype=torch.float32, device=DEV)
    wb = torch.empty_like(wa)
    if sx is not None:
        dt = torch.float32 if mode == "CWS" else torch.int32
        tsx = torch.from_numpy(np.ascontiguousarray(sx)).to(DEV, dt)
        tsy = torch.from_numpy(np.ascontiguousarray(sy)).to(DEV, dt)
        px, py = tsx.data_ptr(), tsy.data_ptr()
    else:
        px = py = None
    _lib.check(L.pivb200_windows(fa.data_ptr(), fb.data_ptr(), B, fa.stride(0), H, W, fa.stride(1), w, o,
                                 _lib.MODES[mode], px, py, wa.data_ptr(), wb.data_ptr(), stream()))
    torch.cuda.synchronize()
    return wa.cpu().numpy(), wb.cpu().numpy()


def pass_first(a, b, w, o, validate=True, val_ratio=1.2, want_ratio=False):
    L = _lib.lib()
    fa, fb = dev_u8(a), dev_u8(b)
    if fa.dim() == 2:
        fa, fb = fa[None], fb[None]
    B, H, W = fa.shape
    nr, nc = field_shape((H, W), w, o)
    u = torch.empty((B, nr, nc), dtype=torch.float64, device=DEV)
    v = torch.empty_like(u)
    m = torch.empty((B, nr, nc), dtype=torch.uint8, device=DEV)
    r = torch.empty((B, nr, nc), dtype=torch.float32, device=DEV)
    _lib.check(L.pivb200_pass_first(fa.data_ptr(), fb.data_ptr(), B, fa.stride(0), H, W, fa.stride(1), w, o,
                                    1 if validate else 0, val_ratio, u.data_ptr(), v.data_ptr(),
                                    m.data_ptr() if validate else None,
                                    r.data_ptr() if want_ratio else None, stream()))
    torch.cuda.synchronize()
    out = [u.cpu().numpy(), v.cpu().numpy(), m.cpu().numpy().astype(bool) if validate else None]
    if want_ratio:
        out.append(r.cpu().numpy())
    return out
