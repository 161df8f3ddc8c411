"""Developer sweep (not a test): OfflinePIV from bmp files, reference-exact mode, vs decode threads / fill workers."""
import os, sys, tempfile, time, shutil, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torchpiv_b200 as T
from torchpiv_b200 import synth
shape = (2048, 2048); K = 16; n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1536
noise, blank = synth.default_patches(shape)
tmp = tempfile.mkdtemp(prefix="pivfiles_")
a, b = synth.particle_pair(shape, synth.uniform_shift(3.3, -2.2), seed=7, noise_patch=noise, blank_patch=blank)
uniq = []
for k in range(K):
    fr = np.roll(a if k % 2 == 0 else b, (7 * (k // 2), 11 * (k // 2)), axis=(0, 1))
    path = os.path.join(tmp, f"uniq{k}.bin"); synth.write_bmp(path, fr); uniq.append(path)
folder = os.path.join(tmp, "seq"); os.makedirs(folder)
for i in range(n_pairs + 1):
    os.link(uniq[i % K], os.path.join(folder, f"frame{i:05d}.bmp"))
def run(dec, fill, prof=False):
    piv = T.OfflinePIV(folder=folder, device="cuda:0", file_fmt="bmp", wind_size=64, overlap=32, multipass=2,
                       multipass_mode="CWS", multipass_scale=2.0, dt=12, scale=0.02, folder_mode="sequential",
                       batch_pairs=32, decode_threads=dec, replace="reference", fill_workers=fill)
    pr = cProfile.Profile() if prof else None
    t0 = time.perf_counter()
    if pr: pr.enable()
    n = sum(1 for _ in piv())
    if pr: pr.disable()
    dt = time.perf_counter() - t0
    piv.close()
    print(f"decode={dec} fill={fill}: {n} pairs in {dt:.2f} s -> {n / dt:.0f} pairs/s", flush=True)
    if pr: pstats.Stats(pr).sort_stats("tottime").print_stats(14)
for dec, fill in ((4, 12), (4, 8), (4, 6), (2, 12), (6, 9), (3, 13)):
    run(dec, fill)
run(4, 12, prof=True)
shutil.rmtree(tmp)
