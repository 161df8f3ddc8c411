"""Developer helper (not a test): per-pass kernel time of the 2-pass CWS plan, measured (a) inside plan.run with events
around each C-ABI call and (b) as bench.py's isolated back-to-back launches of the same call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchpiv_b200 as T
from torchpiv_b200 import synth, _lib
shape = (2048, 2048)
noise, blank = synth.default_patches(shape)
a, b = synth.particle_pair(shape, synth.uniform_shift(3.3, -2.2), seed=0, noise_patch=noise, blank_patch=blank)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
fa = torch.from_numpy(a).cuda()[None].expand(B, -1, -1).contiguous()
fb = torch.from_numpy(b).cuda()[None].expand(B, -1, -1).contiguous()
plan = T.PIVPlan(shape, 64, 32, 2, "CWS", 2.0, device="cuda:0")
for _ in range(3): plan.run(fa, fb)
torch.cuda.synchronize()
stream = torch.cuda.current_stream()

class Proxy:
    def __init__(self, lib): self._lib, self.ev = lib, []
    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if name not in ("pivb200_pass_first", "pivb200_pass_next"): return fn
        def wrapped(*args):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); rc = fn(*args); e1.record(stream)
            self.ev.append((name, e0, e1)); return rc
        return wrapped
real = plan.lib
px = Proxy(real); plan.lib = px
for _ in range(8): plan.run(fa, fb)
torch.cuda.synchronize()
plan.lib = real
import collections
acc = collections.defaultdict(list)
for name, e0, e1 in px.ev: acc[name].append(e0.elapsed_time(e1))
inplan = {k: sum(v[2:]) / len(v[2:]) for k, v in acc.items()}
# isolated, as bench.py does
ws = plan._ws; w0, w1 = ws[0], ws[1]; g0, g1 = plan.passes[0], plan.passes[1]
L = _lib.lib(); st = stream.cuda_stream; ps, pitch = fa.stride(0), fa.stride(1)
def k_first():
    _lib.check(L.pivb200_pass_first(fa.data_ptr(), fb.data_ptr(), B, ps, shape[0], shape[1], pitch, g0.wind, g0.overlap, 1, 1.2,
                                    w0["u"].data_ptr(), w0["v"].data_ptr(), w0["mask"].data_ptr(), None, st))
def k_next():
    _lib.check(L.pivb200_pass_next(fa.data_ptr(), fb.data_ptr(), B, ps, shape[0], shape[1], pitch, g1.wind, g1.overlap, plan.mode,
                                   w1["sx"].data_ptr(), w1["sy"].data_ptr(), w1["base_u"].data_ptr(), w1["base_v"].data_ptr(),
                                   w1["pred_u"].data_ptr(), w1["pred_v"].data_ptr(), 1, 1.2, w1["u"].data_ptr(), w1["v"].data_ptr(),
                                   w1["mask"].data_ptr(), None, st))
def time_kernel(fn, reps=5):
    fn(); torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record(stream)
    for _ in range(reps): fn()
    k1.record(stream); torch.cuda.synchronize()
    return k0.elapsed_time(k1) / reps
print(f"B={B} in-plan: first {inplan['pivb200_pass_first']:.3f} ms, next {inplan['pivb200_pass_next']:.3f} ms | isolated: first {time_kernel(k_first):.3f}, next {time_kernel(k_next):.3f} | isolated again: next {time_kernel(k_next):.3f} first {time_kernel(k_first):.3f}")
