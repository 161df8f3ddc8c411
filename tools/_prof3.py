import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import torchpiv_b200 as T
from torchpiv_b200 import synth
shape = (2048, 2048)
a, b = synth.particle_pair(shape, synth.rankine_vortex(1024, 1024, 256, 6.0), seed=0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
plan = T.PIVPlan(shape, 64, 32, 3, "CWS", 2.0, device="cuda:0")
fa = torch.from_numpy(a).cuda()[None].expand(B, -1, -1).contiguous()
fb = torch.from_numpy(b).cuda()[None].expand(B, -1, -1).contiguous()
for _ in range(3):
    plan.run(fa, fb)
torch.cuda.synchronize()
