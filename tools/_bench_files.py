"""Developer benchmark (not a test): OfflinePIV end to end FROM IMAGE FILES (BASELINE config 5 style).

K unique 4 MP synthetic frames are written once as bmp and hard-linked cyclically to N files, so the decode
cost is paid per file while the disk footprint stays small.  Reports pairs/s of the whole generator
(decode threads -> pinned staging -> H2D -> fused passes -> D2H -> host post-processing).

    python tests/_bench_files.py [n_files] [folder_mode] [decode_threads] [replace]
"""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import torchpiv_b200 as T  # noqa: E402
from torchpiv_b200 import synth  # noqa: E402

n_files = int(sys.argv[1]) if len(sys.argv) > 1 else 257
mode = sys.argv[2] if len(sys.argv) > 2 else "sequential"
threads = int(sys.argv[3]) if len(sys.argv) > 3 else 8
replace = sys.argv[4] if len(sys.argv) > 4 else "reference"
shape = (2048, 2048)
K = 16
noise, blank = synth.default_patches(shape)
tmp = tempfile.mkdtemp(prefix="pivfiles_")
uniq = []
rng_frames = []
# a sequence of frames of one particle field advected by (+3.3, -2.2) px per frame
a, b = synth.particle_pair(shape, synth.uniform_shift(3.3, -2.2), seed=0, noise_patch=noise, blank_patch=blank)
for k in range(K):
    fr = np.roll(a if k % 2 == 0 else b, (7 * (k // 2), 11 * (k // 2)), axis=(0, 1))
    path = os.path.join(tmp, f"uniq{k}.bin")
    synth.write_bmp(path, fr)
    uniq.append(path)
folder = os.path.join(tmp, "seq")
os.makedirs(folder)
for i in range(n_files):
    os.link(uniq[i % K], os.path.join(folder, f"frame{i:05d}.bmp"))
for bp in (16, 32):
    piv = T.OfflinePIV(folder=folder, device="cuda:0", file_fmt="bmp", wind_size=64, overlap=32, multipass=2,
                       multipass_mode="CWS", multipass_scale=2.0, dt=12, scale=0.02, folder_mode=mode,
                       batch_pairs=bp, decode_threads=threads, replace=replace)
    n = 0
    t0 = time.perf_counter()
    for out in piv():
        n += 1
    dt = time.perf_counter() - t0
    print(f"files={n_files} mode={mode} decode_threads={threads} batch_pairs={bp} replace={replace}: "
          f"{len(piv)} pairs, {n} yielded, {dt:.2f} s -> {len(piv) / dt:.1f} pairs/s (cores: {os.cpu_count()})")
import shutil  # noqa: E402
shutil.rmtree(tmp)
