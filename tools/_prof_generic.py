import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchpiv_b200 as T
from torchpiv_b200 import synth
shape = (2048, 2048)
a, b = synth.particle_pair(shape, synth.uniform_shift(3.3, -2.2), seed=0)
B = 2
fa = torch.from_numpy(a).cuda()[None].expand(B, -1, -1).contiguous()
fb = torch.from_numpy(b).cuda()[None].expand(B, -1, -1).contiguous()
w = int(sys.argv[1]); mp = int(sys.argv[2]) if len(sys.argv) > 2 else 1
plan = T.PIVPlan(shape, w, w // 2, mp, "CWS", 2.0, device="cuda:0")
for _ in range(2):
    plan.run(fa, fb)
torch.cuda.synchronize()
