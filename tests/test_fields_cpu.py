"""CPU tests of what follows the correlation passes: the statistics oracle against the reference
worker's own statistics block (tests/golden/statistics.npz), the result-file formats against files
written by the reference's save_table / save_binary (tests/golden/output_formats.npz), and sanity
checks of the (parity-unpinned) normalised-median / stencil-replacement oracle."""
import os

import numpy as np
import pytest

import cases
from oracle import field_oracle as F


@pytest.mark.parametrize("tag", ["a", "b"])
def test_statistics_oracle_matches_reference_worker(golden, tag):
    g = golden("statistics.npz")
    x, y, u, v = g[f"{tag}_x"], g[f"{tag}_y"], g[f"{tag}_u"], g[f"{tag}_v"]
    assert cases.sha(x, y, *u, *v) == str(g[f"{tag}_sha"])
    table = F.statistics_table(x, y, list(u), list(v))
    assert list(table.keys()) == [str(k) for k in g[f"{tag}_keys"]]
    for k, (name, val) in enumerate(table.items()):
        ref = g[f"{tag}_table"][k]
        assert np.allclose(val, ref, rtol=1e-12, atol=1e-12), name


def test_output_formats_byte_exact(golden, tmp_path):
    from torchpiv_b200 import output
    g = golden("output_formats.npz")
    data = {str(k): v for k, v in zip(g["keys"], g["values"])}
    d = str(tmp_path / "out")
    for _ in range(3):
        output.save_table("run_pair.txt", d, data)
        output.save_binary("run_pair.npy", d, data)
    names = sorted(os.listdir(d))
    assert names == [str(n) for n in g["names"]]
    for n in names:
        got = np.frombuffer(open(os.path.join(d, n), "rb").read(), dtype=np.uint8)
        assert np.array_equal(got, g["file_" + n]), n
    # the caller's arrays are not flattened in place
    assert data["x[mm]"].shape == (5, 7)


def test_pair_writer(tmp_path):
    from torchpiv_b200 import output
    x, y = np.meshgrid(np.arange(4.0), np.arange(3.0))
    w = output.PairWriter("/data/run7/", str(tmp_path / "res"), "Save all text")
    p0, p1 = w.pair(x, y, x * 2, y * 3), w.pair(x, y, x, y)
    assert os.path.basename(p0) == "run7_pair.txt" and os.path.basename(p1) == "run7_pair (1).txt"
    rows = open(p0).read().splitlines()
    assert rows[0] == "x[mm], y[mm], Vx[m/s], Vy[m/s]" and len(rows) == 13
    assert rows[6] == "1.000000, 1.000000, 2.000000, 3.000000"
    wb = output.PairWriter("/data/run7", str(tmp_path / "res"), "Save all binary")
    pb = wb.pair(x, y, x, y)
    assert np.load(pb).shape == (4, 3, 4)
    assert output.PairWriter("/data/run7", str(tmp_path / "res"), "Dont save").pair(x, y, x, y) is None
    with pytest.raises(KeyError):
        output.PairWriter("/data/run7", str(tmp_path), "save")


def test_nmt_oracle_flags_a_planted_outlier():
    rng = np.random.default_rng(0)
    u = 3.0 + 0.05 * rng.standard_normal((12, 15))
    v = -2.0 + 0.05 * rng.standard_normal((12, 15))
    u[5, 7] += 4.0
    v[0, 0] -= 3.0          # corner: 3 neighbours
    mask = np.zeros_like(u, dtype=bool)
    mask[9, 9] = True
    out = F.normalized_median_test(u, v, mask)
    assert out[5, 7] and out[0, 0] and out[9, 9]
    assert out.sum() == 3


def test_stencil_replace_oracle_fills_from_the_rim():
    u = np.fromfunction(lambda r, c: 1.0 + 0 * r + 0 * c, (9, 9))
    v = np.fromfunction(lambda r, c: c * 1.0, (9, 9))
    bad = np.zeros((9, 9), dtype=bool)
    bad[2:7, 2:7] = True            # 5x5 hole: three sweeps reach the centre
    fu, fv, left = F.stencil_replace(u, v, bad, 2)
    assert left[4, 4] and left.sum() == 1 and fu[4, 4] == 0.0
    fu, fv, left = F.stencil_replace(u, v, bad, 3)          # rounded up to 4 sweeps
    assert not left.any() and np.all(fu == 1.0)
    assert np.abs(fv - v).max() <= 2.0       # medians of a ramp lag behind it inside the hole
    # nothing usable at all: stays flagged, becomes zero
    fu, fv, left = F.stencil_replace(u, v, np.ones((9, 9), dtype=bool), 4)
    assert left.all() and not fu.any() and not fv.any()


def test_merge_states_equals_one_accumulator(golden):
    """Host-side merge of per-shard moments (multi-GPU statistics) == moments of the whole sequence."""
    from torchpiv_b200.postprocess_device import merge_states
    g = golden("statistics.npz")
    u, v = g["a_u"], g["a_v"]

    def state(lo, hi):
        a, b = u[lo:hi], v[lo:hi]
        ma, mb = a.mean(0), b.mean(0)
        return hi - lo, np.stack([ma, mb, ((a - ma) ** 2).sum(0), ((b - mb) ** 2).sum(0), ((a - ma) * (b - mb)).sum(0)])

    n, mom = merge_states([state(0, 1), (0, None), state(1, 4), state(4, 6)])
    wn, want = state(0, 6)
    assert n == wn and np.allclose(mom, want, rtol=1e-12, atol=1e-12)
    with pytest.raises(ValueError):
        merge_states([(0, None)])
