import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo/tests/golden")
import numpy as np, cases, gpu_util as G
from oracle import piv_oracle as O
a, b = cases.small_pair(seed=1)
w, o = int(sys.argv[1]), int(sys.argv[2])
wa, wb = G.windows(a, b, w, o)
ra = O.moving_window_array(a, w, o).astype(np.float32)
print("W", w, o, "mismatch", (wa != ra).sum())
