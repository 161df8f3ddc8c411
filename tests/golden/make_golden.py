"""Generate the golden vectors in tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The reference backend is imported as is through oracle/ref_loader.py (four stub modules
for its unused GUI / imageio imports) and run with device="cpu".  Inputs are regenerated
from seeds (tests/golden/cases.py); the fixtures store the reference's outputs and a
SHA-256 of the inputs.  Nothing under /root/reference is copied.
"""
import contextlib
import io
import json
import os
import shutil
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import cases  # noqa: E402
from oracle import ref_loader  # noqa: E402
from torchpiv_b200 import synth  # noqa: E402

PB = ref_loader.load_ref()
CPU = torch.device("cpu")


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def save(name, **arrays):
    np.savez_compressed(os.path.join(HERE, name), **arrays)
    print("wrote", name, {k: getattr(v, "shape", None) for k, v in arrays.items()})


def gen_pass1():
    out = {}
    for kind, zero in (("uniform", False), ("vortex", True)):
        a, b = cases.small_pair(seed=1, kind=kind, zero_patch=zero)
        out[f"{kind}_sha"] = np.array(cases.sha(a, b))
        for w, o in cases.PASS1_GEOMS:
            u, v, x, y, m = PB.extended_search_area_piv(torch.tensor(a), torch.tensor(b),
                                                        window_size=w, overlap=o, validate=True)
            out[f"{kind}_{w}_{o}_u"], out[f"{kind}_{w}_{o}_v"] = u, v
            out[f"{kind}_{w}_{o}_x"], out[f"{kind}_{w}_{o}_y"] = x, y
            out[f"{kind}_{w}_{o}_mask"] = m
    # validate=False path (mask None)
    a, b = cases.small_pair(seed=1)
    u, v, x, y, m = PB.extended_search_area_piv(torch.tensor(a), torch.tensor(b),
                                                window_size=32, overlap=16, validate=False)
    assert m is None
    out["noval_32_16_u"], out["noval_32_16_v"] = u, v
    save("pass1.npz", **out)


def gen_multipass():
    for mode in ("CWS", "DWS"):
        out = {}
        for kind in ("uniform", "vortex"):
            a, b = cases.small_pair(seed=2, kind=kind)
            ta, tb = torch.tensor(a), torch.tensor(b)
            out[f"{kind}_sha"] = np.array(cases.sha(a, b))
            u, v, x, y, m = PB.extended_search_area_piv(ta, tb, window_size=64, overlap=32,
                                                        validate=True)
            out[f"{kind}_p0_u"], out[f"{kind}_p0_v"], out[f"{kind}_p0_mask"] = u, v, m
            w, o = 64, 32
            for it in (1, 2):
                w, o = int(w // 2.0), int(o // 2.0)
                fn = PB.IterModMap.functions[mode](ta.shape, w, o, CPU)
                u, v, x, y, m = quiet(fn, ta, tb, x, y, u.copy(), v.copy(), m.copy())
                out[f"{kind}_p{it}_u"], out[f"{kind}_p{it}_v"] = u, v
                out[f"{kind}_p{it}_mask"] = m
                out[f"{kind}_p{it}_x"], out[f"{kind}_p{it}_y"] = x, y
        # validation_mask=None branch (validate False inside the pass)
        a, b = cases.small_pair(seed=2)
        ta, tb = torch.tensor(a), torch.tensor(b)
        u, v, x, y, m = PB.extended_search_area_piv(ta, tb, window_size=64, overlap=32, validate=True)
        fn = PB.IterModMap.functions[mode](ta.shape, 32, 16, CPU)
        u, v, x, y, m = quiet(fn, ta, tb, x, y, u.copy(), v.copy(), None)
        assert m is None
        out["noval_p1_u"], out["noval_p1_v"] = u, v
        save(f"multipass_{mode}.npz", **out)


def gen_shift():
    out = {}
    for w, o in ((32, 16), (16, 8), (64, 32)):
        a, vx, vy = cases.shift_case(seed=w, w=w, ovl=o)
        ta = torch.tensor(a)
        idx = PB.moving_window_array(
            torch.arange(0, a.size, dtype=torch.int64).reshape(a.shape), w, o)
        r = PB.biliniar_interpolation_CWS(ta, idx, torch.from_numpy(vx)[:, None, None],
                                          torch.from_numpy(vy)[:, None, None]).numpy()
        out[f"cws_{w}_sha"] = np.array(cases.sha(r))
        out[f"cws_{w}_head"] = r[:6].copy()
        out[f"cws_{w}_sum"] = r.sum(axis=(1, 2), dtype=np.float64)
        ix, iy = np.rint(vx).astype(np.int64), np.rint(vy).astype(np.int64)
        r = PB.interpolation_DWS(ta, idx, torch.from_numpy(ix)[:, None, None],
                                 torch.from_numpy(iy)[:, None, None]).numpy()
        assert r.dtype == np.uint8
        out[f"dws_{w}_sha"] = np.array(cases.sha(r))
        out[f"dws_{w}_head"] = r[:6].copy()
        out[f"dws_{w}_sum"] = r.sum(axis=(1, 2), dtype=np.int64)
        out[f"in_{w}_sha"] = np.array(cases.sha(a, vx, vy))
    save("shift.npz", **out)


def gen_corr2disp():
    out = {}
    for w, dt, c in ((16, np.float32, 400), (64, np.float32, 400), (16, np.float64, 400),
                     (32, np.float64, 400)):
        tag = f"{w}_{np.dtype(dt).name}"
        maps = cases.adversarial_maps(seed=w + (dt == np.float64), c=c, w=w, dtype=dt)
        out[f"{tag}_sha"] = np.array(cases.sha(maps))
        u, v, m = PB.correlation_to_displacement(torch.from_numpy(maps.copy()), 20, 20, True)
        out[f"{tag}_u"], out[f"{tag}_v"], out[f"{tag}_mask"] = u, v, m
    # real correlation maps through the reference's correalte_fft (float32 and uint8 inputs)
    a, b = cases.small_pair(seed=4)
    ta, tb = torch.tensor(a), torch.tensor(b)
    aa = PB.moving_window_array(ta, 32, 16)
    bb = PB.moving_window_array(tb, 32, 16)
    corr = PB.correalte_fft(aa, bb)
    assert corr.dtype == torch.float32
    out["corr_u8_32_head"] = corr[:4].numpy().copy()
    out["corr_u8_32_sum"] = corr.sum(dim=(1, 2), dtype=torch.float64).numpy()
    save("corr2disp.npz", **out)


def gen_offline():
    tmp = tempfile.mkdtemp(prefix="pivgold_")
    try:
        pairs = [cases.small_pair(seed=10 + i, kind="uniform" if i % 2 == 0 else "vortex")
                 for i in range(3)]
        synth.write_pair_folder(tmp, pairs)
        out = {"in_sha": np.array(cases.sha(*[f for p in pairs for f in p]))}
        runs = {
            "cws2": dict(wind_size=64, overlap=32, multipass=2, multipass_mode="CWS",
                         multipass_scale=2.0, dt=12, scale=0.02, folder_mode="pairs"),
            "dws2": dict(wind_size=64, overlap=32, multipass=2, multipass_mode="DWS",
                         multipass_scale=2.0, dt=12, scale=0.02, folder_mode="pairs"),
            "cws3": dict(wind_size=64, overlap=32, multipass=3, multipass_mode="CWS",
                         multipass_scale=2.0, dt=1, scale=1.0, folder_mode="pairs"),
            "single": dict(wind_size=32, overlap=16, multipass=1, dt=1, scale=1.0,
                           folder_mode="pairs"),
            "seq_cws2": dict(wind_size=64, overlap=32, multipass=2, multipass_mode="CWS",
                             multipass_scale=2.0, dt=5, scale=0.5, folder_mode="sequential"),
        }
        for tag, kw in runs.items():
            gen = PB.OfflinePIV(folder=tmp, device="cpu", file_fmt="bmp", **kw)
            res = quiet(lambda: list(gen()))
            out[f"{tag}_n"] = np.array([len(gen), len(res)])
            for i, (x, y, u, v) in enumerate(res):
                out[f"{tag}_{i}_x"], out[f"{tag}_{i}_y"] = x, y
                out[f"{tag}_{i}_u"], out[f"{tag}_{i}_v"] = u, v
        save("offline.npz", **out)
    finally:
        shutil.rmtree(tmp)


def _worker_statistics_source():
    """The statistics block of the reference's PIVWorker.run (workers.py: from `if u_inst:` to the end
    of the `table = {...}` literal), read from /root/reference at generation time and dedented.  The
    worker module itself cannot be imported (Qt thread classes); its numerics are plain NumPy."""
    import textwrap
    path = os.path.join(ref_loader.REF_ROOT, "workers.py")
    lines = open(path).read().splitlines()
    start = next(i for i, ln in enumerate(lines) if ln.strip() == "if u_inst:")
    end = next(i for i in range(start, len(lines)) if lines[i].strip() == "}" and "table" in "".join(lines[start:i]))
    return textwrap.dedent("\n".join(lines[start:end + 1]))


def gen_statistics():
    import types
    rng = np.random.default_rng(7)
    out = {}
    for tag, (nr, nc, n, scale) in {"a": (9, 13, 6, 0.02), "b": (17, 21, 3, 1.0)}.items():
        xs, ys = np.meshgrid((np.arange(nc) * 16.0 + 16.0) * scale, (np.arange(nr) * 16.0 + 16.0) * scale)
        u_list = [rng.normal(3.0, 1.0, (nr, nc)) + 0.05 * xs for _ in range(n)]
        v_list = [rng.normal(-2.0, 0.5, (nr, nc)) - 0.03 * ys * xs for _ in range(n)]
        sink = types.SimpleNamespace(emit=lambda *_: None)
        ns = {"np": np, "x": xs, "y": ys, "u_inst": list(u_list), "v_inst": list(v_list),
              "self": types.SimpleNamespace(progress=sink, avg_u=None, avg_v=None)}
        exec(_worker_statistics_source(), ns)
        out[f"{tag}_sha"] = np.array(cases.sha(xs, ys, *u_list, *v_list))
        out[f"{tag}_x"], out[f"{tag}_y"] = xs, ys
        out[f"{tag}_u"], out[f"{tag}_v"] = np.stack(u_list), np.stack(v_list)
        out[f"{tag}_keys"] = np.array(list(ns["table"].keys()))
        out[f"{tag}_table"] = np.stack(list(ns["table"].values()))
    save("statistics.npz", **out)


def gen_output_formats():
    PF = sys.modules["torchPIV.PlotterFunctions"]
    rng = np.random.default_rng(11)
    data = {"x[mm]": rng.uniform(0, 40, (5, 7)), "y[mm]": rng.uniform(0, 40, (5, 7)),
            "Vx[m/s]": rng.normal(0, 3, (5, 7)), "Vy[m/s]": rng.normal(0, 1e-4, (5, 7))}
    tmp = tempfile.mkdtemp(prefix="pivgold_")
    try:
        d = os.path.join(tmp, "out")
        for _ in range(3):      # the second and third call exercise the " (n)" numbering
            PF.save_table("run_pair.txt", d, {k: v.copy() for k, v in data.items()})
            PF.save_binary("run_pair.npy", d, {k: v.copy() for k, v in data.items()})
        names = sorted(os.listdir(d))
        out = {"names": np.array(names), "keys": np.array(list(data.keys())),
               "values": np.stack(list(data.values()))}
        for n in names:
            out["file_" + n] = np.frombuffer(open(os.path.join(d, n), "rb").read(), dtype=np.uint8)
        save("output_formats.npz", **out)
    finally:
        shutil.rmtree(tmp)


def gen_general_sizes():
    """Window sizes outside 16/32/64 (general direct-DFT kernel): first passes, a 64 -> 42 -> 28 chain
    (what multipass_scale = 1.5 produces, PB:855-857) at the function boundary, and the generator."""
    out = {}
    a, b = cases.small_pair(seed=3, kind="vortex", zero_patch=True)
    ta, tb = torch.tensor(a), torch.tensor(b)
    out["sha"] = np.array(cases.sha(a, b))
    for w, o in cases.GENERAL_GEOMS:
        u, v, x, y, m = PB.extended_search_area_piv(ta, tb, window_size=w, overlap=o, validate=True)
        out[f"p1_{w}_{o}_u"], out[f"p1_{w}_{o}_v"], out[f"p1_{w}_{o}_mask"] = u, v, m
        out[f"p1_{w}_{o}_x"], out[f"p1_{w}_{o}_y"] = x, y
    for mode in ("CWS", "DWS"):
        u, v, x, y, m = PB.extended_search_area_piv(ta, tb, window_size=64, overlap=32, validate=True)
        out[f"{mode}_p0_u"], out[f"{mode}_p0_v"], out[f"{mode}_p0_mask"] = u, v, m
        w, o = 64, 32
        for it in (1, 2):
            w, o = int(w // 1.5), int(o // 1.5)
            fn = PB.IterModMap.functions[mode](ta.shape, w, o, CPU)
            u, v, x, y, m = quiet(fn, ta, tb, x, y, u.copy(), v.copy(), m.copy())
            out[f"{mode}_p{it}_u"], out[f"{mode}_p{it}_v"], out[f"{mode}_p{it}_mask"] = u, v, m
            out[f"{mode}_p{it}_x"], out[f"{mode}_p{it}_y"] = x, y
    # halving into an ODD window: 50 -> 25 px (the reference's maps are then [25, 24], PB:255)
    for mode in ("CWS", "DWS"):
        u, v, x, y, m = PB.extended_search_area_piv(ta, tb, window_size=50, overlap=25, validate=True)
        out[f"odd{mode}_p0_u"], out[f"odd{mode}_p0_v"], out[f"odd{mode}_p0_mask"] = u, v, m
        fn = PB.IterModMap.functions[mode](ta.shape, 25, 12, CPU)
        u, v, x, y, m = quiet(fn, ta, tb, x, y, u.copy(), v.copy(), m.copy())
        out[f"odd{mode}_p1_u"], out[f"odd{mode}_p1_v"], out[f"odd{mode}_p1_mask"] = u, v, m
        out[f"odd{mode}_p1_x"], out[f"odd{mode}_p1_y"] = x, y
    tmp = tempfile.mkdtemp(prefix="pivgold_")
    try:
        pairs = [cases.small_pair(seed=10 + i, kind="uniform" if i % 2 == 0 else "vortex") for i in range(2)]
        synth.write_pair_folder(tmp, pairs)
        gen = PB.OfflinePIV(folder=tmp, device="cpu", file_fmt="bmp", wind_size=64, overlap=32, multipass=2,
                            multipass_mode="CWS", multipass_scale=1.5, dt=12, scale=0.02, folder_mode="pairs")
        res = quiet(lambda: list(gen()))
        out["offline_n"] = np.array([len(gen), len(res)])
        for i, (x, y, u, v) in enumerate(res):
            out[f"offline_{i}_x"], out[f"offline_{i}_y"] = x, y
            out[f"offline_{i}_u"], out[f"offline_{i}_v"] = u, v
    finally:
        shutil.rmtree(tmp)
    save("general_sizes.npz", **out)


GENERATORS = {"pass1": gen_pass1, "multipass": gen_multipass, "shift": gen_shift, "corr2disp": gen_corr2disp,
              "offline": gen_offline, "statistics": gen_statistics, "output_formats": gen_output_formats,
              "general_sizes": gen_general_sizes}

if __name__ == "__main__":
    # no arguments: regenerate everything; otherwise only the named fixtures
    for name in (sys.argv[1:] or list(GENERATORS)):
        GENERATORS[name]()
    import scipy
    meta = {"torch": torch.__version__, "numpy": np.__version__, "scipy": scipy.__version__,
            "reference": "NikNazarov/TorchPIV src/torchPIV/PIVbackend.py (unmodified, device=cpu)",
            "threads": torch.get_num_threads()}
    with open(os.path.join(HERE, "VERSIONS.json"), "w") as fh:
        json.dump(meta, fh, indent=1)
    print(meta)
