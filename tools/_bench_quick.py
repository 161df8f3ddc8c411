import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import torchpiv_b200 as T
from torchpiv_b200 import synth
shape = (2048, 2048)
noise, blank = synth.default_patches(shape)
a, b = synth.particle_pair(shape, synth.uniform_shift(3.3, -2.2), seed=0, noise_patch=noise, blank_patch=blank)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
modes = sys.argv[2:] or ["CWS2"]
fa = torch.from_numpy(a).cuda()[None].expand(B, -1, -1).contiguous()
fb = torch.from_numpy(b).cuda()[None].expand(B, -1, -1).contiguous()
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for m in modes:
    mp = int(m[3:]); mode = m[:3]
    plan = T.PIVPlan(shape, 64, 32, mp, mode, 2.0, device="cuda:0")
    ms = timeit(lambda: plan.run(fa, fb))
    p1 = T.PIVPlan(shape, 64, 32, 1, mode, 2.0, device="cuda:0")
    ms1 = timeit(lambda: p1.run(fa, fb))
    print(f"sync={os.environ.get('PIVB200_SYNC_MASK')} {m} B={B}: {ms/B*1e3:.1f} us/pair ({B/ms*1e3:.0f} pairs/s); pass1 only {ms1/B*1e3:.1f} us/pair")
