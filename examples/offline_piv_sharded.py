#!/usr/bin/env python
"""OfflinePIV over a folder on N GPUs of one box: one process per GPU, pairs sharded by index, no collective on
the data path; the small result fields and the per-rank statistics are gathered on the host of rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
        examples/offline_piv_sharded.py --folder DIR [--fmt bmp] [--mode sequential] [--out OUTDIR]

Without --folder a small synthetic sequence is rendered first (demo / smoke test).  Single process works too."""
import argparse
import os
import sys
import tempfile
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchpiv_b200 as T  # noqa: E402
from torchpiv_b200 import output, sharding, synth  # noqa: E402
from torchpiv_b200.postprocess_device import merge_states  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--folder")
    ap.add_argument("--fmt", default="bmp")
    ap.add_argument("--mode", default="sequential", choices=["pairs", "sequential"])
    ap.add_argument("--out")
    ap.add_argument("--frames", type=int, default=17, help="synthetic demo: number of frames")
    args = ap.parse_args()
    rank, world, local = sharding.dist_env()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    folder = args.folder
    if folder is None:
        folder = os.path.join(tempfile.gettempdir(), "pivb200_demo")
        if rank == 0:
            os.makedirs(folder, exist_ok=True)
            a, _ = synth.particle_pair((512, 512), synth.uniform_shift(0.0, 0.0), seed=0)
            for i in range(args.frames):       # a particle field advected by (+3, -2) px per frame (periodic)
                synth.write_bmp(os.path.join(folder, f"frame{i:04d}.bmp"), np.roll(a, (-2 * i, 3 * i), axis=(0, 1)))
        sharding.barrier()

    piv = T.OfflinePIV(folder=folder, device=f"cuda:{local}", file_fmt=args.fmt, wind_size=64, overlap=32, multipass=2,
                       multipass_mode="CWS", multipass_scale=2.0, dt=12, scale=0.02, folder_mode=args.mode,
                       shard=(rank, world), replace="stencil", statistics=True, batch_pairs=8)
    t0 = time.perf_counter()
    # pairs can be skipped (unreadable frame, ...): take the pair number from the generator, do not count yields
    local_results = [(piv.last_pair_index, x, y, u, v) for (x, y, u, v) in piv()]
    dt = time.perf_counter() - t0
    results = sharding.gather_results(local_results)                       # pair order, rank 0 only
    state = piv.statistics.state() if piv.statistics is not None else (0, None)
    states = [state]
    if world > 1:
        bucket = [None] * world if rank == 0 else None
        dist.gather_object(state, bucket, dst=0)
        states = bucket
    if rank == 0:
        total = sum(s[0] for s in states)
        print(f"{len(results)} pairs from {world} rank(s) in {dt:.2f} s; statistics over {total} fields")
        geo = piv._plan.out_geometry
        table = piv.statistics.table(geo.x, geo.y, 0.02, 12, state=merge_states(states))
        print("mean Vx, Vy [m/s]:", float(table["Vx[m/s]"].mean()), float(table["Vy[m/s]"].mean()))
        if args.out:
            w = output.PairWriter(folder, args.out, "Save all binary")
            for _, x, y, u, v in results:
                w.pair(x, y, u, v)
            output.PairWriter(folder, args.out, "Save all text").statistics(table)
            print("wrote", len(results), "pair files and the statistics table to", args.out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
