"""Micro-benchmark (not a test, not product code): the DFT-as-GEMM alternative that BASELINE.json's
north_star allows "only if measured faster within tolerance".

The 2-D circular cross-correlation of N 64x64 windows is evaluated as four dense DFT-matrix products
(library GEMMs through torch.matmul = cuBLAS, i.e. the best case for tensor-core throughput, with every
intermediate making an HBM round trip) in several precisions, and compared with

  * accuracy: the float64 FFT correlation of the same windows -- relative map error and the error of the
    3-point log-Gaussian sub-pixel estimate (the quantity the 1e-3 px tolerance is about);
  * speed: the fused FP32 in-register FFT kernel of this repo on the same number of windows
    (pivb200_correlate + the pass kernel time is printed by bench.py; here the plain correlate entry).

Run on the GPU box:  python tests/micro/dft_gemm.py [pairs]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

W = 64


def dft_mats(dtype, device):
    k = np.arange(W)
    ang = 2 * np.pi * np.outer(k, k) / W
    C, S = np.cos(ang), np.sin(ang)
    t = lambda a: torch.tensor(a, dtype=torch.float64, device=device).to(dtype)   # noqa: E731
    # forward along columns:  [Yr | Yi] = x @ [C | -S];  forward along rows: [[C, S], [-S, C]] @ [Yr; Yi]
    right = t(np.concatenate([C, -S], axis=1))                      # [64, 128]
    left = t(np.block([[C, S], [-S, C]]))                           # [128, 128]
    left_inv = t(np.block([[C, -S], [S, C]]))                       # conjugate
    right_inv = t(np.concatenate([C, -S], axis=0))                  # [Pr | Pi] @ [C; -S] = real part of the inverse
    return right, left, left_inv, right_inv


def corr_gemm(a, b, mats):
    """a, b: [N, 64, 64] in the compute dtype.  Returns the (unshifted, unnormalised) correlation, float32."""
    right, left, left_inv, right_inv = mats

    def fwd(x):
        y = torch.matmul(x, right)                                  # [N, 64, 128] = [Yr | Yi]
        y = torch.cat([y[..., :W], y[..., W:]], dim=1)              # [N, 128, 64] = [Yr; Yi]
        return torch.matmul(left, y)                                # [N, 128, 64] = [Zr; Zi]
    A, B = fwd(a).float(), fwd(b).float()
    Ar, Ai, Br, Bi = A[:, :W], A[:, W:], B[:, :W], B[:, W:]
    Pr, Pi = Ar * Br + Ai * Bi, Ar * Bi - Ai * Br                   # conj(A) * B
    P = torch.cat([Pr, Pi], dim=1).to(a.dtype)                      # [N, 128, 64]
    Q = torch.matmul(left_inv, P)                                   # inverse along rows
    Q = torch.cat([Q[:, :W], Q[:, W:]], dim=2)                      # [N, 64, 128] = [Qr | Qi]
    return torch.matmul(Q, right_inv).float()                       # real part, [N, 64, 64]


def subpixel(c):
    """3-point log fit around the peak of fft-shifted float64 maps (interior peaks only)."""
    c = torch.fft.fftshift(c.double(), dim=(-2, -1))
    c = c - c.amin(dim=(-2, -1), keepdim=True) + 1e-7 * c.amax(dim=(-2, -1), keepdim=True)
    n = c.shape[0]
    m = c.reshape(n, -1).argmax(dim=1)
    r, col = (m // W).clamp(1, W - 2), (m % W).clamp(1, W - 2)
    idx = torch.arange(n, device=c.device)
    l0 = torch.log(c[idx, r, col])
    lx0, lx1 = torch.log(c[idx, r, col - 1]), torch.log(c[idx, r, col + 1])
    ly0, ly1 = torch.log(c[idx, r - 1, col]), torch.log(c[idx, r + 1, col])
    u = col + (lx0 - lx1) / (2 * (lx0 + lx1) - 4 * l0)
    v = r + (ly0 - ly1) / (2 * (ly0 + ly1) - 4 * l0)
    return u, v, m


def timeit(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    dev = torch.device("cuda:0")
    from torchpiv_b200 import synth
    import torchpiv_b200 as T
    shape = (2048, 2048)
    a_img, b_img = synth.particle_pair(shape, synth.uniform_shift(3.3, -2.2), seed=0)
    fa, fb = torch.from_numpy(a_img).to(dev), torch.from_numpy(b_img).to(dev)
    wa = T.moving_window_array(fa, W, 32).float()
    wb = T.moving_window_array(fb, W, 32).float()
    # every variant gets mean-free windows (what the fused kernel does by dropping the DC bin): the best
    # case for the low-precision GEMMs
    wa = wa - wa.mean(dim=(1, 2), keepdim=True)
    wb = wb - wb.mean(dim=(1, 2), keepdim=True)
    wa, wb = wa.repeat(pairs, 1, 1).contiguous(), wb.repeat(pairs, 1, 1).contiguous()
    n = wa.shape[0]
    print(f"{n} windows of {W}x{W} ({pairs} 4 MP pairs, 50 % overlap)")

    ref = torch.fft.irfft2(torch.conj(torch.fft.rfft2(wa[:3969].double())) * torch.fft.rfft2(wb[:3969].double()))
    ru, rv, rm = subpixel(ref)

    def report(name, fn, flops_scale):
        c = fn()[:3969] / (W * W)
        if not torch.isfinite(c).all():
            print(f"{name:34s} overflow: spectra of 64x64 uint8 windows exceed the fp16 range")
            return
        rel = float((c.double() - ref).abs().amax() / ref.abs().amax())
        u, v, m = subpixel(c)
        same = (m == rm)
        err = float(torch.maximum((u - ru).abs(), (v - rv).abs())[same].max())
        ms = timeit(fn)
        # four GEMM stages, real arithmetic: 2*64*64*128 + 2*128*128*64 per forward transform (x2 frames), same inverse
        flops = n * (3 * (2 * 64 * 64 * 128 + 2 * 128 * 128 * 64)) * flops_scale
        print(f"{name:34s} {ms:8.2f} ms = {ms / pairs * 1e3:7.1f} us/pair  {flops / ms / 1e9:7.1f} TFLOP/s(dense)  "
              f"map rel err {rel:.1e}  peak mismatches {int((~same).sum()):4d}/3969  max |d| {err:.1e} px")

    for name, dtype, tf32, scale in (("fp32 SGEMM (no tensor cores)", torch.float32, False, 1),
                                     ("tf32 tensor cores, 1 term", torch.float32, True, 1),
                                     ("bf16 tensor cores, 1 term", torch.bfloat16, False, 1),
                                     ("fp16 tensor cores, 1 term", torch.float16, False, 1)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        mats = dft_mats(dtype, dev)
        xa, xb = wa.to(dtype), wb.to(dtype)
        try:
            report(name, lambda: corr_gemm(xa, xb, mats), scale)
        except RuntimeError as exc:
            print(f"{name:34s} failed: {str(exc)[:80]}")
    torch.backends.cuda.matmul.allow_tf32 = False

    # 3xTF32 error-compensated product (hi*hi + hi*lo + lo*hi), the cheapest variant with FP32-like accuracy
    def split(x):
        hi = (x.view(torch.int32) & ~0x1fff).view(torch.float32)
        return hi, x - hi
    mats32 = dft_mats(torch.float32, dev)

    def mm3(x, y):
        torch.backends.cuda.matmul.allow_tf32 = True
        xh, xl = split(x.contiguous())
        yh, yl = split(y.contiguous())
        return torch.matmul(xh, yh) + (torch.matmul(xh, yl) + torch.matmul(xl, yh))

    def corr_3xtf32():
        right, left, left_inv, right_inv = mats32

        def fwd(x):
            y = mm3(x, right)
            y = torch.cat([y[..., :W], y[..., W:]], dim=1)
            return mm3(left, y)
        A, B = fwd(wa), fwd(wb)
        Ar, Ai, Br, Bi = A[:, :W], A[:, W:], B[:, :W], B[:, W:]
        P = torch.cat([Ar * Br + Ai * Bi, Ar * Bi - Ai * Br], dim=1)
        Q = mm3(left_inv, P)
        Q = torch.cat([Q[:, :W], Q[:, W:]], dim=2)
        return mm3(Q, right_inv)
    report("3xTF32 split (compensated)", corr_3xtf32, 3)
    torch.backends.cuda.matmul.allow_tf32 = False

    # the fused in-register FFT path of this repo on the same windows (correlation maps written to HBM)
    ms = timeit(lambda: T.correalte_fft(wa, wb))
    c = T.correalte_fft(wa[:3969], wb[:3969])
    c = torch.fft.ifftshift(c, dim=(-2, -1))
    rel = float((c.double() - ref).abs().amax() / ref.abs().amax())
    u, v, m = subpixel(c)
    same = (m == rm)
    err = float(torch.maximum((u - ru).abs(), (v - rv).abs())[same].max())
    print(f"{'fused FP32 FFT kernel (this repo)':34s} {ms:8.2f} ms = {ms / pairs * 1e3:7.1f} us/pair  (maps to HBM)          "
          f"map rel err {rel:.1e}  peak mismatches {int((~same).sum()):4d}/3969  max |d| {err:.1e} px")


if __name__ == "__main__":
    main()
