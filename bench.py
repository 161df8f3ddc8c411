#!/usr/bin/env python
"""Benchmark of the PIV cross-correlation hot path (BASELINE.json metric:
4 MP pairs/s, 64 px windows, 50 % overlap, 2-pass CWS).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path (pass 1 at 64/32 + predictor + CWS pass 2 at 32/16) over 16 batches of 32
synthetic 2048x2048 particle-image pairs per GPU (512 pairs; 20 steps last about a second).

  value : whole-job pairs/s with the frames already resident in HBM (CUDA events on the launch stream, barrier +
          synchronize on both sides, max over ranks).  A batch (2 x 32 x 4 MB = 268 MB of frames) is larger than
          the 126 MB L2, so every launch re-reads it from HBM.
  e2e   : the same metric through the public host API (HostPipeline: pinned host frames -> H2D -> fused kernels
          -> D2H of u, v, mask), copies inside the timed region for every batch; next to it the ceiling a plain
          pinned cudaMemcpyAsync of the same buffers reaches in the same run (all ranks at once).
  e2e_files : BASELINE config 5 -- a sequential-mode folder of bmp files through OfflinePIV(...)(), sharded over
          the ranks, reference-exact hole filling and the device stencil.
  roofline     : dominant kernel (CWS pass at 32 px) against the FFMA peak measured in this run.
  cpu_baseline : the UNMODIFIED reference (baseline/_ref) with device=cpu on the host cores, bounded sample,
                 rank 0 only (kind "reference"; the NumPy oracle port, kind "port", when baseline/_ref is absent).
  torch_eager_baseline : the UNMODIFIED reference with device=cuda on the same GPU (its own torch-CUDA path),
                 bounded sample, rank 0 only -- the comparator north_star names.

Multi-GPU (torchrun, one rank per GPU): pairs are independent, so each rank processes its own shard of pairs
(weak scaling); torch.distributed is used for the barrier and the max-over-ranks of the device time only --
there is no collective on the data path.

--impl reference times the reference's own CPU path (baseline/_ref, device=cpu, all host threads).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SHAPE = (2048, 2048)
WIND, OVERLAP, PASSES, MODE, SCALE = 64, 32, 2, "CWS", 2.0
PAIRS_PER_STEP = 32          # per GPU; 268 MB of frames > L2
UNIQUE_PAIRS = 4             # rendered on the host, then rolled into PAIRS_PER_STEP distinct pairs
METRIC = "4MP pairs/sec (64px, 50% ovl, 2-pass CWS)"
WORKLOAD = ("synthetic 2048x2048 uniform-shift particle pairs (+3.3,-2.2 px, noise + blank patch), "
            "64->32 px windows, 50% overlap, 2-pass CWS")


def flops_per_window(w: int) -> float:
    """three real 2-D transforms + half-spectrum conjugate multiply."""
    return 15.0 * w * w * np.log2(w) + 6.0 * w * (w / 2 + 1)


def pass_geometry():
    from torchpiv_b200.engine import pass_schedule
    out = []
    for w, o in pass_schedule(WIND, OVERLAP, PASSES, SCALE):
        n = ((SHAPE[0] - w) // (w - o) + 1) * ((SHAPE[1] - w) // (w - o) + 1)
        out.append((w, o, n))
    return out


def make_pairs(n_unique: int, seed0: int = 0):
    from torchpiv_b200 import synth
    noise, blank = synth.default_patches(SHAPE)
    return [synth.particle_pair(SHAPE, synth.uniform_shift(3.3, -2.2), seed=seed0 + i,
                                noise_patch=noise, blank_patch=blank) for i in range(n_unique)]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(names, parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax,
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    return rank, world, local


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline / torch-CUDA comparator
# ------------------------------------------------------------------------------------------------
# The UNMODIFIED reference (baseline/_ref, pip-installed by baseline/install_ref.sh and shipped with the working
# tree) is used whenever it is present: kind = "reference".  Without it the legs fall back to the restatements
# under oracle/ (kind = "port") and say so.
def load_reference():
    """PIVbackend of the unmodified reference, or None."""
    try:
        from oracle import ref_loader
        if ref_loader.available():
            return ref_loader.load_ref()
    except Exception as exc:                       # noqa: BLE001 - report and fall back to the ports
        print(f"# reference not importable: {exc}", file=sys.stderr)
    return None


def reference_pass_rate(PB, device, pairs, n_pairs: int, warmup: int):
    """pairs/s of the reference's own pass functions (PB:874-882: extended_search_area_piv, then
    piv_iteration_CWS.__call__) on `device`, frames already decoded (and resident on `device`)."""
    import torch
    frames = [(torch.from_numpy(a).to(device), torch.from_numpy(b).to(device)) for a, b in pairs]
    w, o = WIND, OVERLAP
    iters = []
    for _ in range(PASSES - 1):
        w, o = int(w // SCALE), int(o // SCALE)
        iters.append(PB.IterModMap.functions[MODE](SHAPE, w, o, device))

    def one(i):
        a, b = frames[i % len(frames)]
        u, v, x, y, val = PB.extended_search_area_piv(a, b, window_size=WIND, overlap=OVERLAP, validate=True)
        for it in iters:
            u, v, x, y, val = it(a, b, x, y, u, v, val)
        return u

    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):      # the reference prints a line per pass
        for i in range(warmup):
            one(i)
        if device.type == "cuda":
            torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        for i in range(n_pairs):
            one(i)
        if device.type == "cuda":
            torch.cuda.synchronize(device)
        dt = time.perf_counter() - t0
    return n_pairs / dt, dt


def reference_offline_rate(PB, device_key: str, n_pairs: int, warmup: int):
    """pairs/s of the reference's public generator OfflinePIV(...)() from bmp files (decode, H2D, passes,
    host hole filling: PB:862-903), pairs = files 2i, 2i+1."""
    import contextlib, io, shutil, tempfile
    from torchpiv_b200 import synth
    pairs = make_pairs(2, seed0=100)
    tmp = tempfile.mkdtemp(prefix="pivref_")
    try:
        synth.write_pair_folder(tmp, [pairs[i % 2] for i in range(n_pairs + warmup)])
        gen = PB.OfflinePIV(folder=tmp, device=device_key, file_fmt="bmp", wind_size=WIND, overlap=OVERLAP,
                            multipass=PASSES, multipass_mode=MODE, dt=1, scale=1.0, multipass_scale=SCALE)
        n = 0
        t0 = None
        with contextlib.redirect_stdout(io.StringIO()):
            for _ in gen():
                n += 1
                if n == warmup:
                    t0 = time.perf_counter()
        dt = time.perf_counter() - (t0 if t0 is not None else 0.0)
        return (n - warmup) / dt if t0 is not None and n > warmup else 0.0, n - warmup
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def cpu_oracle_rate(n_pairs: int, warmup: int = 0):
    """pairs/s of the CPU oracle port (2-pass CWS, all host threads) on n_pairs 4 MP pairs."""
    from oracle import piv_oracle as O
    pairs = make_pairs(min(n_pairs, 2), seed0=100)
    iters = [O.ITER_MODES[MODE](SHAPE, w, o, workers=-1) for (w, o, _) in pass_geometry()[1:]]
    for i in range(warmup):
        a, b = pairs[i % len(pairs)]
        O.piv_passes(a, b, WIND, OVERLAP, PASSES, MODE, SCALE, workers=-1, iter_objs=iters)
    t0 = time.perf_counter()
    for i in range(n_pairs):
        a, b = pairs[i % len(pairs)]
        O.piv_passes(a, b, WIND, OVERLAP, PASSES, MODE, SCALE, workers=-1, iter_objs=iters)
    dt = time.perf_counter() - t0
    return n_pairs / dt, dt


def cpu_reference_rate(n_pairs: int, warmup: int):
    """(pairs/s, seconds, kind, what) of the reference path on the host cores."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)        # torchrun exports OMP_NUM_THREADS=1
    PB = load_reference()
    if PB is not None:
        rate, dt = reference_pass_rate(PB, torch.device("cpu"), make_pairs(min(n_pairs, 2), seed0=100), n_pairs, warmup)
        return rate, dt, "reference", ("unmodified reference (baseline/_ref), device=cpu: extended_search_area_piv + "
                                       "piv_iteration_CWS.__call__ (PB:874-882), torch CPU kernels, all host threads")
    rate, dt = cpu_oracle_rate(n_pairs, warmup)
    return rate, dt, "port", "CPU oracle port (NumPy / scipy.fft, all host threads); baseline/_ref is absent"


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    pairs_per_step = 1
    rate, dt, kind, what = cpu_reference_rate(args.steps * pairs_per_step, warmup=min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 (pass 1) / f32 (pass 2), as the reference computes", "data": "synthetic",
        # same keys as the GPU arm's config (the sample is bounded: one pair per step on the host cores)
        "config": {"workload": WORKLOAD, "pairs_per_step_per_gpu": pairs_per_step, "batches_per_step": 1,
                   "pairs_per_batch": pairs_per_step, "frame_bytes_per_batch_per_gpu": 2 * pairs_per_step * SHAPE[0] * SHAPE[1],
                   "l2_policy": "n/a (host cores)", "sharding": "rank 0 only"},
        "reference": {"device": "cpu", "what": what},
        "cpu_baseline": {"value": rate, "unit": "pairs/s", "cores": cores, "kind": kind,
                         "sample": f"{args.steps} x 1 4MP pair, 2-pass CWS, pass functions only (no image decode)"},
        "e2e": {"value": rate, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def files_leg(T, synth, dev, rank, world, n_pairs_total, decode_threads, fill_workers, per_rank):
    """BASELINE config 5: a sequential-mode folder of 4 MP bmp files through the public generator
    OfflinePIV(...)() (decode threads -> pinned staging -> H2D -> fused passes -> D2H -> host post-processing),
    sharded by pair index over the ranks (shard=(rank, world), no collective).  K unique frames are written once
    and hard-linked cyclically, so every file is decoded while the disk footprint stays small.  Returns
    {mode: (pairs done by this rank, seconds)} for the reference-exact hole filling and the device stencil."""
    import shutil
    import tempfile
    K = 16
    tmp = tempfile.mkdtemp(prefix=f"pivfiles_r{rank}_")
    out = {}
    try:
        noise, blank = synth.default_patches(SHAPE)
        a, b = synth.particle_pair(SHAPE, synth.uniform_shift(3.3, -2.2), seed=7, noise_patch=noise, blank_patch=blank)
        uniq = []
        for k in range(K):
            fr = np.roll(a if k % 2 == 0 else b, (7 * (k // 2), 11 * (k // 2)), axis=(0, 1))
            path = os.path.join(tmp, f"uniq{k}.bin")
            synth.write_bmp(path, fr)
            uniq.append(path)
        folder = os.path.join(tmp, "seq")
        os.makedirs(folder)
        for i in range(n_pairs_total + 1):
            os.link(uniq[i % K], os.path.join(folder, f"frame{i:05d}.bmp"))
        # reference mode: a few decode threads + worker processes for Qhull; stencil mode: every core decodes
        for mode, kw in (("reference", dict(replace="reference", fill_workers=fill_workers, decode_threads=decode_threads)),
                         ("stencil", dict(replace="stencil", decode_threads=max(2, per_rank - 4)))):
            piv = T.OfflinePIV(folder=folder, device=f"cuda:{dev.index}", file_fmt="bmp", wind_size=WIND, overlap=OVERLAP,
                               multipass=PASSES, multipass_mode=MODE, multipass_scale=SCALE, dt=12, scale=0.02,
                               folder_mode="sequential", batch_pairs=32, shard=(rank, world), **kw)
            n = 0
            t0 = time.perf_counter()
            for _ in piv():
                n += 1
            out[mode] = (n, len(piv), time.perf_counter() - t0)
            piv.close()
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    import torchpiv_b200 as T
    from torchpiv_b200 import _lib, synth
    from torchpiv_b200.engine import FramePipeline, HostPipeline

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: torchpiv_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def reduce(x: float, op) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(ms: float) -> float:
        return reduce(ms, dist.ReduceOp.MAX) if world > 1 else ms

    B, R = args.pairs_per_batch, args.batches_per_step
    # ---- synthetic data: a few rendered pairs, rolled into B distinct pairs per rank -----------
    base = make_pairs(UNIQUE_PAIRS, seed0=1000 * rank)
    host_a = torch.empty((B,) + SHAPE, dtype=torch.uint8).pin_memory()
    host_b = torch.empty((B,) + SHAPE, dtype=torch.uint8).pin_memory()
    for i in range(B):
        a, b = base[i % UNIQUE_PAIRS]
        sh = (37 * (i // UNIQUE_PAIRS), 53 * (i // UNIQUE_PAIRS))
        host_a[i].copy_(torch.from_numpy(np.roll(a, sh, axis=(0, 1))))
        host_b[i].copy_(torch.from_numpy(np.roll(b, sh, axis=(0, 1))))
    fa, fb = host_a.to(dev), host_b.to(dev)

    plan = T.PIVPlan(SHAPE, WIND, OVERLAP, PASSES, MODE, SCALE, device=dev)
    stream = torch.cuda.current_stream(dev)

    # ---- device-resident timing: a step = R batches of B pairs (the timed region lasts ~1 s) -----
    for _ in range(max(args.warmup, 3)):
        plan.run(fa, fb)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps * R):
        plan.run(fa, fb)
    e1.record(stream)
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * R * args.steps / (ms_total * 1e-3)

    # sanity: the benchmarked computation produced the imposed displacement
    u, v, m = plan.run(fa[:1], fb[:1])
    ok = ~m[0].bool()
    med_u, med_v = float(u[0][ok].median()), float(v[0][ok].median())
    if abs(med_u - 3.3) > 0.1 or abs(med_v + 2.2) > 0.1:
        raise SystemExit(f"benchmark output is wrong: median displacement {med_u:.3f}, {med_v:.3f}")

    # ---- plain pinned H2D copies, all ranks at once: the ceiling of the end-to-end number ---------------
    cstream = torch.cuda.Stream(dev)
    dst = torch.empty_like(fa)
    with torch.cuda.stream(cstream):
        dst.copy_(host_a, non_blocking=True)
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_copy = 8
    with torch.cuda.stream(cstream):
        c0.record(cstream)
        for i in range(n_copy):
            dst.copy_(host_a if i % 2 == 0 else host_b, non_blocking=True)
        c1.record(cstream)
    cstream.synchronize()
    barrier()
    copy_ms = max_over_ranks(c0.elapsed_time(c1))
    h2d_ceiling_gbs = world * n_copy * host_a.numel() / (copy_ms * 1e-3) / 1e9
    del dst

    # ---- end to end through the host API: every batch pays its H2D and D2H -------------------------------
    pipe = HostPipeline(plan, B)
    for _ in range(2):
        pipe.result(pipe.submit(host_a, host_b))
    barrier()
    h0, d0 = pipe.h2d_bytes, pipe.d2h_bytes
    t0 = time.perf_counter()
    pending = None
    checksum = 0.0
    for _ in range(args.steps * R):
        sid = pipe.submit(host_a, host_b)
        if pending is not None:
            ru, rv, rm = pipe.result(pending)
            checksum += float(ru[0, 0, 0])
        pending = sid
    ru, rv, rm = pipe.result(pending)
    checksum += float(ru[0, 0, 0])
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_ms = max_over_ranks(e2e_s * 1e3)
    e2e_value = world * B * R * args.steps / (e2e_ms * 1e-3)
    h2d_step = (pipe.h2d_bytes - h0) // args.steps
    d2h_step = (pipe.d2h_bytes - d0) // args.steps
    e2e_h2d_gbs = world * h2d_step * args.steps / (e2e_ms * 1e-3) / 1e9

    # ---- the same, sequential-folder style: pair i = frames (i, i+1), a batch of B pairs uploads B + 1 frames ----
    del pipe
    fpipe = FramePipeline(plan, B)
    for sid in (0, 1):
        stack = fpipe.host_frames(sid)
        stack[:B] = host_a.numpy()
        stack[B] = host_b[B - 1].numpy()
    for sid in (0, 1):
        fpipe.submit(sid, B, chained=True)
        fpipe.result(sid)
    barrier()
    sh0 = fpipe.h2d_bytes
    seq_batches = max(args.steps * R // 4, 4)
    t0 = time.perf_counter()
    prev = None
    for i in range(seq_batches):
        fpipe.submit(i & 1, B, chained=True)
        if prev is not None:
            checksum += float(fpipe.result(prev)[0][0, 0, 0])
        prev = i & 1
    checksum += float(fpipe.result(prev)[0][0, 0, 0])
    barrier()
    seq_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    seq_value = world * B * seq_batches / (seq_ms * 1e-3)
    seq_h2d_batch = (fpipe.h2d_bytes - sh0) // seq_batches
    del fpipe

    # ---- BASELINE config 5: the sequential batch FROM FILES through OfflinePIV, sharded over the ranks ----------
    files = None
    if args.files_pairs > 0:
        cores = os.cpu_count() or 1
        per_rank = max(1, cores // world)
        decode_threads = max(2, min(8, per_rank // 4))
        # measured on the 16-vCPU box (tools/_files_sweep.py): 4 decode threads + 6 worker processes; more workers
        # oversubscribe the cores (main thread, CUDA and executor threads need theirs) and the rate drops again
        fill_workers = max(2, (3 * per_rank) // 8)
        barrier()
        t0 = time.perf_counter()
        res = files_leg(T, synth, dev, rank, world, args.files_pairs, decode_threads, fill_workers, per_rank)
        files = {}
        for mode, (n_yield, n_pairs, secs) in res.items():
            t_max = max_over_ranks(secs * 1e3) * 1e-3
            tot = reduce(float(n_pairs), dist.ReduceOp.SUM) if world > 1 else float(n_pairs)
            files[mode] = {"value": tot / t_max, "unit": "pairs/s", "pairs": int(tot), "seconds": t_max,
                           "yielded_rank0": n_yield}
        files["config"] = {"files": args.files_pairs + 1, "folder_mode": "sequential", "unique_frames": 16,
                           "decode_threads_per_rank": decode_threads, "fill_workers_per_rank": fill_workers,
                           "host_cores": cores, "api": "torchpiv_b200.OfflinePIV(...)() with shard=(rank, world)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (rank 0, timed alone on its stream) --------------------
    import ctypes
    geo = pass_geometry()
    ws = plan._ws

    def time_kernel(fn, reps=5):
        fn()
        torch.cuda.synchronize(dev)
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(stream)
        for _ in range(reps):
            fn()
        k1.record(stream)
        torch.cuda.synchronize(dev)
        return k0.elapsed_time(k1) / reps

    L = _lib.lib()
    st = stream.cuda_stream
    w0, w1 = ws[0], ws[1]
    g0, g1 = plan.passes[0], plan.passes[1]
    ps, pitch = fa.stride(0), fa.stride(1)

    def k_first():
        _lib.check(L.pivb200_pass_first(fa.data_ptr(), fb.data_ptr(), B, ps, SHAPE[0], SHAPE[1], pitch, g0.wind,
                                        g0.overlap, 1, 1.2, w0["u"].data_ptr(), w0["v"].data_ptr(),
                                        w0["mask"].data_ptr(), None, st))

    def k_next():
        _lib.check(L.pivb200_pass_next(fa.data_ptr(), fb.data_ptr(), B, ps, SHAPE[0], SHAPE[1], pitch, g1.wind,
                                       g1.overlap, plan.mode, w1["sx"].data_ptr(), w1["sy"].data_ptr(),
                                       w1["base_u"].data_ptr(), w1["base_v"].data_ptr(), w1["pred_u"].data_ptr(),
                                       w1["pred_v"].data_ptr(), 1, 1.2, w1["u"].data_ptr(), w1["v"].data_ptr(),
                                       w1["mask"].data_ptr(), None, st))

    # the legs above ran the same plan on other inputs (the sequential-folder leg pairs unrelated frames): restore the
    # benchmark's own predictor field in the workspace, so that the second pass is timed on the shifts of the timed region
    plan.run(fa, fb)
    torch.cuda.synchronize(dev)
    ms_first, ms_next = time_kernel(k_first), time_kernel(k_next)
    peak = ctypes.c_double()
    _lib.check(L.pivb200_measure_fp32_peak(10, ctypes.byref(peak), st))
    flops_first = B * geo[0][2] * flops_per_window(geo[0][0])
    flops_next = B * geo[1][2] * flops_per_window(geo[1][0])
    achieved = flops_next / (ms_next * 1e-3) / 1e12
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj.get("pass_next_cws_w32_dram_bytes_per_pair")
            traffic = traffic * B if traffic is not None else None
            traffic_src = tj.get("source")
        except (OSError, ValueError):
            traffic = None
    peaks = {}
    ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(ppath):
        peaks = json.load(open(ppath))
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    bytes_next = B * (2 * SHAPE[0] * SHAPE[1] + 18 * geo[1][2])
    ms_step = ms_total / (args.steps * R)            # one batch of B pairs
    roofline = {"bound": "fp32", "kernel": "piv_soa_kernel<32, CWS> (pass 2)", "achieved": achieved,
                "peak": peak.value, "peak_source": "FFMA micro-benchmark run by this bench.py (MEASURED_PEAKS.json has no FP32 figure)",
                "unit": "TFLOP/s", "frac": achieved / peak.value, "traffic": traffic, "traffic_source": traffic_src,
                "ms_per_launch": ms_next, "algorithmic_flops_per_launch": flops_next,
                "share_of_step": ms_next / ms_step,
                "pass_first": {"kernel": "piv_soa_kernel<64, ALN> (pass 1)", "ms_per_launch": ms_first,
                               "achieved": flops_first / (ms_first * 1e-3) / 1e12,
                               "frac": flops_first / (ms_first * 1e-3) / 1e12 / peak.value},
                "whole_step_frac": (flops_first + flops_next) / (ms_step * 1e-3) / 1e12 / peak.value,
                "hbm": {"achieved": bytes_next / (ms_next * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": bytes_next / (ms_next * 1e-3) / 1e9 / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"}}

    # ---- CPU baseline (bounded sample): the reference on the host cores --------------------------
    cores = os.cpu_count() or 1
    n_cpu = args.cpu_pairs
    if n_cpu > 0:
        cpu_rate, cpu_dt, cpu_kind, cpu_what = cpu_reference_rate(n_cpu, warmup=1)
    else:
        cpu_rate, cpu_dt, cpu_kind, cpu_what = None, 0.0, None, "skipped (--cpu-pairs 0)"
    # ---- the reference's own torch-CUDA path on this GPU (bounded sample) -------------------------
    torch_cuda = None
    if args.torch_pairs > 0:
        PB = load_reference()
        if PB is not None:
            rate, te = reference_pass_rate(PB, dev, [(host_a[i].numpy(), host_b[i].numpy()) for i in range(4)],
                                           args.torch_pairs, warmup=3)
            key = next((k for k, d in PB.DeviceMap.devicies.items() if d == dev or str(d) == str(dev)), None)
            off_rate, off_n = (reference_offline_rate(PB, key, 12, 2) if key is not None else (None, 0))
            torch_cuda = {"value": rate, "unit": "pairs/s", "kind": "reference",
                          "what": "UNMODIFIED reference (baseline/_ref) on this GPU: extended_search_area_piv + "
                                  "piv_iteration_CWS.__call__ (PB:874-882) with device=cuda, frames resident",
                          "sample": f"{args.torch_pairs} 4MP pairs after 3 warm-up pairs, 2-pass CWS, {te:.2f} s",
                          "offline_piv_from_files": {"value": off_rate, "unit": "pairs/s", "pairs": off_n,
                                                     "what": "reference OfflinePIV(...)() generator from bmp files "
                                                             "(decode, H2D, passes, host hole filling), device=cuda"}}
        else:
            from oracle import torch_eager as E
            second = E.IterCWS(SHAPE, WIND // 2, OVERLAP // 2, dev)
            E.two_pass_cws(fa[0], fb[0], WIND, OVERLAP, second)              # warm-up (cuFFT plans, allocator)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for i in range(args.torch_pairs):
                E.two_pass_cws(fa[i % B], fb[i % B], WIND, OVERLAP, second)
            torch.cuda.synchronize(dev)
            te = time.perf_counter() - t0
            torch_cuda = {"value": args.torch_pairs / te, "unit": "pairs/s", "kind": "port",
                          "what": "oracle/torch_eager.py (restatement with stock eager PyTorch); baseline/_ref is absent",
                          "sample": f"{args.torch_pairs} 4MP pairs, 2-pass CWS, {te:.2f} s"}
            del second
        torch.cuda.empty_cache()

    line = {
        "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_step_per_gpu": B * R, "batches_per_step": R, "pairs_per_batch": B,
                   "frame_bytes_per_batch_per_gpu": 2 * B * SHAPE[0] * SHAPE[1],
                   "l2_policy": "inputs larger than L2 (268 MB of frames per batch vs 126 MB L2)",
                   "sharding": "pairs split across ranks, no data-path collective"},
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d_step),
                "d2h_bytes_per_step": int(d2h_step), "ms_per_step": e2e_ms / args.steps,
                "h2d_gbs": e2e_h2d_gbs, "h2d_ceiling_gbs": h2d_ceiling_gbs,
                "h2d_ceiling_what": "plain pinned cudaMemcpyAsync of the same buffers, all ranks at once, measured in this run",
                "api": "torchpiv_b200.engine.HostPipeline (pinned host frames in, u/v/mask on the host out)"},
        "e2e_sequential": {"value": seq_value, "unit": "pairs/s", "h2d_bytes_per_batch": int(seq_h2d_batch),
                           "batches": seq_batches,
                           "note": "informative: sequential-folder pairing (pair i = frames i, i+1), every frame "
                                   "uploaded once; same kernels, torchpiv_b200.engine.FramePipeline"},
        "e2e_files": files,
        "gpu_launches": int(launches), "launches_per_batch": plan.launches_per_batch,
        "clocks": clocks, "roofline": roofline,
        "cpu_baseline": {"value": cpu_rate, "unit": "pairs/s", "cores": cores, "kind": cpu_kind,
                         "sample": f"{n_cpu} 4MP pairs, 2-pass CWS pass functions, {cpu_dt:.1f} s; {cpu_what}"},
        "torch_eager_baseline": torch_cuda,
        "check": {"median_u_px": med_u, "median_v_px": med_v, "imposed": [3.3, -2.2]},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs-per-batch", type=int, default=PAIRS_PER_STEP, help="pairs per kernel launch (per GPU)")
    ap.add_argument("--batches-per-step", type=int, default=16,
                    help="batches per step: a step is 16 x 32 = 512 pairs per GPU, so that 20 steps last ~1 s")
    ap.add_argument("--cpu-pairs", type=int, default=3, help="size of the bounded CPU-baseline sample")
    ap.add_argument("--torch-pairs", type=int, default=24,
                    help="size of the bounded reference-on-GPU sample (0 = skip)")
    ap.add_argument("--files-pairs", type=int, default=4000,
                    help="pairs of the from-files leg (BASELINE config 5: 4 000 pairs = 4 001 files; 0 = skip)")
    args = ap.parse_args()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
