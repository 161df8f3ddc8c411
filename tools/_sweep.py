"""Developer timing helper (not a test): per-pass kernel times for one configuration."""
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
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
out = []
for (w, o, mp, mode, name) in [(64, 32, 1, "CWS", "w64"), (32, 16, 1, "CWS", "w32int"), (16, 8, 1, "CWS", "w16int"),
                               (64, 32, 2, "CWS", "cws64+32"), (64, 32, 2, "DWS", "dws64+32"), (64, 32, 3, "CWS", "cws64+32+16")]:
    plan = T.PIVPlan(shape, w, o, mp, mode, 2.0, device="cuda:0")
    ms = timeit(lambda: plan.run(fa, fb))
    out.append(f"{name} {ms / B * 1e3:.1f}")
print(f"nwarps={os.environ.get('PIVB200_NWARPS')} sync={os.environ.get('PIVB200_SYNC_MASK')} B={B} us/pair: " + " | ".join(out))
