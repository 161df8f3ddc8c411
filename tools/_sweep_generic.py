"""Developer timing helper (not a test): general-size plans (csrc/generic_pass.cuh), us per 4 MP pair."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchpiv_b200 as T
from torchpiv_b200 import synth
shape = (2048, 2048)
a, b = synth.particle_pair(shape, synth.uniform_shift(3.3, -2.2), seed=0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
fa = torch.from_numpy(a).cuda()[None].expand(B, -1, -1).contiguous()
fb = torch.from_numpy(b).cuda()[None].expand(B, -1, -1).contiguous()
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
out = []
for (w, o, mp, mode, sc, name) in [(48, 24, 1, "CWS", 2.0, "w48"), (128, 64, 1, "CWS", 2.0, "w128"), (160, 80, 1, "CWS", 2.0, "w160"), (256, 128, 1, "CWS", 2.0, "w256"), (200, 100, 2, "CWS", 2.0, "cws200+100"),
                                   (24, 12, 1, "CWS", 2.0, "w24"), (48, 24, 2, "CWS", 2.0, "cws48+24"),
                                   (64, 32, 2, "CWS", 1.5, "cws64+42"), (128, 64, 3, "CWS", 2.0, "cws128+64+32"),
                                   (68, 34, 1, "CWS", 2.0, "w68(direct)")]:
    plan = T.PIVPlan(shape, w, o, mp, mode, sc, device="cuda:0")
    ms = timeit(lambda: plan.run(fa, fb))
    out.append(f"{name} {ms / B * 1e3:.0f}")
print(f"direct={os.environ.get('PIVB200_GENERIC_DIRECT')} B={B} us/pair: " + " | ".join(out))
