"""Developer timing helper (not a test): 64 px first pass and the 2-pass CWS plan only, us per 4 MP pair."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchpiv_b200 as T
from torchpiv_b200 import synth
shape = (2048, 2048)
noise, blank = synth.default_patches(shape)
a, b = synth.particle_pair(shape, synth.uniform_shift(3.3, -2.2), seed=0, noise_patch=noise, blank_patch=blank)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
fa = torch.from_numpy(a).cuda()[None].expand(B, -1, -1).contiguous()
fb = torch.from_numpy(b).cuda()[None].expand(B, -1, -1).contiguous()
def timeit(fn, n=8):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
p1 = T.PIVPlan(shape, 64, 32, 1, "CWS", 2.0, device="cuda:0")
p2 = T.PIVPlan(shape, 64, 32, 2, "CWS", 2.0, device="cuda:0")
t1 = timeit(lambda: p1.run(fa, fb)) / B * 1e3
t2 = timeit(lambda: p2.run(fa, fb)) / B * 1e3
print(f"w64 {t1:.1f} | cws64+32 {t2:.1f} | pass2 {t2 - t1:.1f}")
