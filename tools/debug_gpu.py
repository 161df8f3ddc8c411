"""Developer diagnostic (not a test): prints CUDA-vs-oracle/golden deviations on the GPU box."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases, gpu_util as G
from oracle import piv_oracle as O
import torchpiv_b200 as T
from torchpiv_b200 import _lib

print(torch.cuda.get_device_name(0))
gold = lambda n: np.load(os.path.join(ROOT, "tests", "golden", n))

def stat(name, got, ref):
    d = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    print(f"  {name}: max {d.max():.3e}  mismatch {(got != ref).sum()}/{got.size}")

# 1. loader parity (windows)
a, b = cases.small_pair(seed=1)
for w, o in [(64, 32), (32, 16), (16, 8), (32, 8)]:
    wa, wb = G.windows(a, b, w, o)
    ra = O.moving_window_array(a, w, o).astype(np.float32); rb = O.moving_window_array(b, w, o).astype(np.float32)
    print("windows INT", w, o); stat("a", wa, ra); stat("b", wb, rb)
for w, o in [(32, 16), (16, 8), (64, 32)]:
    fr, vx, vy = cases.shift_case(seed=w, w=w, ovl=o)
    idx = O.window_index_grid(fr.shape, w, o)
    wa, wb = G.windows(fr, fr, w, o, "CWS", vx, vy)
    ra = O.bilinear_interpolation_cws(fr, idx, -vx[:, None, None], -vy[:, None, None])
    rb = O.bilinear_interpolation_cws(fr, idx, vx[:, None, None], vy[:, None, None])
    print("windows CWS", w, o); stat("a", wa, ra); stat("b", wb, rb)
    bad = np.unique(np.argwhere(wb != rb)[:, 0])
    if bad.size: print("   bad windows b:", bad[:20], vx[bad[:5]], vy[bad[:5]])
    ix, iy = np.rint(vx).astype(np.int64), np.rint(vy).astype(np.int64)
    wa, wb = G.windows(fr, fr, w, o, "DWS", ix, iy)
    ra = O.interpolation_dws(fr, idx, -ix[:, None, None], -iy[:, None, None]).astype(np.float32)
    rb = O.interpolation_dws(fr, idx, ix[:, None, None], iy[:, None, None]).astype(np.float32)
    print("windows DWS", w, o); stat("a", wa, ra); stat("b", wb, rb)

# 2. correlate
for w in (64, 32, 16):
    aa = O.moving_window_array(a, w, w // 2); bb = O.moving_window_array(b, w, w // 2)
    ref = O.correlate_fft(aa, bb)
    got = T.correalte_fft(torch.from_numpy(aa.copy()).cuda(), torch.from_numpy(bb.copy()).cuda()).cpu().numpy()
    print("correlate", w, "rel err", np.abs(got - ref).max() / np.abs(ref).max())

# 3. pass 1 vs golden
g = gold("pass1.npz")
for kind, zero in (("uniform", False), ("vortex", True)):
    a, b = cases.small_pair(seed=1, kind=kind, zero_patch=zero)
    for w, o in cases.PASS1_GEOMS:
        u, v, m = G.pass_first(a, b, w, o)
        ru, rv, rm = g[f"{kind}_{w}_{o}_u"], g[f"{kind}_{w}_{o}_v"], g[f"{kind}_{w}_{o}_mask"]
        ok = ~rm
        print(f"pass1 {kind} {w}/{o}: mask mismatches {(m[0] != rm).sum()}/{rm.size} (invalid {rm.sum()}), "
              f"du valid max {np.abs(u[0] - ru)[ok].max():.2e} dv {np.abs(v[0] - rv)[ok].max():.2e} "
              f"all max {np.abs(u[0] - ru).max():.2e} {np.abs(v[0] - rv).max():.2e}")

# 4. later passes vs golden (function boundary)
for mode in ("CWS", "DWS"):
    g = gold(f"multipass_{mode}.npz")
    for kind in ("uniform", "vortex"):
        a, b = cases.small_pair(seed=2, kind=kind)
        fa, fb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
        x, y = O.get_coordinates(a.shape, 64, 32)
        w, o = 64, 32
        for it in (1, 2):
            u0, v0, m0 = (g[f"{kind}_p{it-1}_{k}"].copy() for k in ("u", "v", "mask"))
            w, o = w // 2, o // 2
            fn = T.IterModMap.functions[mode](a.shape, w, o, "cuda:0")
            u, v, x1, y1, m = fn(fa, fb, x, y, u0, v0, m0)
            ru, rv, rm = g[f"{kind}_p{it}_u"], g[f"{kind}_p{it}_v"], g[f"{kind}_p{it}_mask"]
            eu, ev = np.abs(u - ru), np.abs(v - rv)
            print(f"pass{it+1} {mode} {kind} w={w}: mask mism {(m != rm).sum()}/{rm.size} (invalid {rm.sum()}) "
                  f"err max {eu.max():.2e} {ev.max():.2e} q99 {np.quantile(eu, .99):.2e} n>1e-3: {(eu > 1e-3).sum() + (ev > 1e-3).sum()}")
            x, y = x1, y1

# 5. full plan 2048^2
from torchpiv_b200 import synth
shape = (2048, 2048)
noise, blank = synth.default_patches(shape)
t0 = time.time(); a, b = synth.particle_pair(shape, synth.uniform_shift(3.3, -2.2), seed=0, noise_patch=noise, blank_patch=blank); print("synth", time.time() - t0)
plan = T.PIVPlan(shape, 64, 32, 2, "CWS", 2.0, device="cuda:0")
fa, fb = torch.from_numpy(a).cuda()[None], torch.from_numpy(b).cuda()[None]
u, v, m = plan.run(fa, fb); torch.cuda.synchronize()
u, v, m = u.cpu().numpy()[0], v.cpu().numpy()[0], m.cpu().numpy()[0].astype(bool)
print("2048 2-pass CWS: median", np.median(u), np.median(v), "invalid", m.sum())
t0 = time.time(); ou, ov, ox, oy, oval, hist = O.piv_passes(a, b, 64, 32, 2, "CWS"); print("oracle s", time.time() - t0)
ok = ~oval
print("  vs oracle: mask mism", (m != oval).sum(), "err max(valid)", np.abs(u - ou)[ok].max(), np.abs(v - ov)[ok].max(), "q999", np.quantile(np.abs(u - ou), .999))
for B in (1, 8):
    fa8, fb8 = fa.expand(B, -1, -1).contiguous(), fb.expand(B, -1, -1).contiguous()
    plan.run(fa8, fb8); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): plan.run(fa8, fb8)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"B={B}: {ms:.3f} ms/batch -> {B / ms * 1e3:.1f} pairs/s")
import ctypes
tf = ctypes.c_double()
_lib.check(_lib.lib().pivb200_measure_fp32_peak(10, ctypes.byref(tf), None)); print("FFMA peak TFLOP/s", tf.value)
