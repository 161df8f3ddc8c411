"""GPU tests of the on-device field post-processing (pivb200_nmt / pivb200_replace /
pivb200_stats_accumulate, csrc/field_ops.cu) against oracle/field_oracle.py, and of
OfflinePIV(replace="stencil", statistics=True).  Medians of identical float64 values are exact, so
the stencil kernels must agree with the oracle bit for bit; the statistics (running sums vs the
reference worker's two-pass means) within 1e-9 relative."""
import numpy as np
import pytest

import cases
from oracle import field_oracle as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    import torch
    assert torch.cuda.is_available()
    from torchpiv_b200 import postprocess_device
    return postprocess_device


def _fields(seed, B, nr, nc, holes=True):
    rng = np.random.default_rng(seed)
    u = 3.0 + 0.1 * rng.standard_normal((B, nr, nc))
    v = -2.0 + 0.1 * rng.standard_normal((B, nr, nc))
    spikes = rng.random((B, nr, nc)) < 0.03
    u[spikes] += rng.choice([-5.0, 5.0], size=int(spikes.sum()))
    mask = rng.random((B, nr, nc)) < (0.08 if holes else 0.0)
    if holes:
        mask[0, 2:8, 3:10] = True               # a hole several vectors deep
        mask[-1, :, :] = B > 1                  # one field entirely invalid
    return u, v, mask


def _dev(a, dtype=None):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


@pytest.mark.parametrize("shape", [(3, 17, 21), (1, 1, 1), (2, 2, 9), (2, 63, 63)])
@pytest.mark.parametrize("with_mask", [True, False])
def test_nmt_matches_oracle(P, shape, with_mask):
    import torch
    u, v, mask = _fields(1, *shape)
    got = P.normalized_median_test(_dev(u), _dev(v), _dev(mask, torch.uint8) if with_mask else None)
    torch.cuda.synchronize()
    got = got.cpu().numpy().astype(bool)
    for b in range(shape[0]):
        want = F.normalized_median_test(u[b], v[b], mask[b] if with_mask else None)
        assert np.array_equal(got[b], want), b


@pytest.mark.parametrize("shape,sweeps", [((3, 17, 21), 2), ((3, 17, 21), 7), ((1, 1, 1), 1), ((2, 33, 5), 40)])
def test_replace_matches_oracle(P, shape, sweeps):
    import torch
    u, v, mask = _fields(2, *shape)
    du, dv, dm = _dev(u), _dev(v), _dev(mask, torch.uint8)
    P.replace_invalid(du, dv, dm, sweeps)
    torch.cuda.synchronize()
    for b in range(shape[0]):
        wu, wv, wm = F.stencil_replace(u[b], v[b], mask[b], sweeps)
        assert np.array_equal(dm[b].cpu().numpy().astype(bool), wm)
        assert np.array_equal(du[b].cpu().numpy(), wu) and np.array_equal(dv[b].cpu().numpy(), wv)


def test_replace_rejects_bad_arguments(P):
    import torch
    u, v, mask = _fields(3, 2, 9, 9)
    with pytest.raises(RuntimeError):
        P.replace_invalid(torch.from_numpy(u), torch.from_numpy(v), torch.from_numpy(mask.astype(np.uint8)))
    with pytest.raises(TypeError):
        P.replace_invalid(_dev(u).float(), _dev(v).float(), _dev(mask, torch.uint8))
    with pytest.raises(RuntimeError):
        P.replace_invalid(_dev(u), _dev(v), _dev(mask, torch.uint8), max_sweeps=0)


def test_statistics_match_the_reference_table(P, golden):
    """Fields in px, un-flipped, in two batches -> the table the reference's worker computes from the
    finished (flipped, scaled) fields."""
    g = golden("statistics.npz")
    scale, dt = 0.02, 12.0
    k = scale / dt * 1000
    u_fin, v_fin = g["b_u"], g["b_v"]                   # what the generator would have yielded
    nr, nc = u_fin.shape[1:]
    xs, ys = np.meshgrid(np.arange(nc) * 16.0 + 16.0, np.arange(nr) * 16.0 + 16.0)     # px
    u_px, v_px = np.flip(u_fin, axis=1) / k, -np.flip(v_fin, axis=1) / k
    st = P.FieldStatistics(nr, nc, "cuda:0")
    st.add(_dev(u_px[:2]), _dev(v_px[:2]))
    st.add(u_px[2], v_px[2])                            # NumPy input, single field
    assert st.count == 3
    got = st.table(xs, ys, scale, dt)
    want = F.statistics_table(xs * scale, ys * scale, list(u_fin), list(v_fin))
    assert list(got.keys()) == list(want.keys())
    for name in want:
        ref = want[name]
        assert np.allclose(got[name], ref, rtol=1e-9, atol=1e-9 * max(1.0, np.abs(ref).max())), name


def test_offline_piv_stencil_mode_and_statistics(tmp_path):
    import torchpiv_b200 as T
    from torchpiv_b200 import synth
    pairs = [cases.small_pair(seed=10 + i, kind="uniform" if i % 2 == 0 else "vortex") for i in range(3)]
    synth.write_pair_folder(str(tmp_path), pairs)
    kw = dict(folder=str(tmp_path), device="cuda:0", file_fmt="bmp", wind_size=64, overlap=32, multipass=2,
              multipass_mode="CWS", dt=12, scale=0.02, batch_pairs=2)
    ref = list(T.OfflinePIV(**kw)())
    gen = T.OfflinePIV(replace="stencil", statistics=True, **kw)
    res = list(gen())
    assert len(res) == 3                     # the stencil mode never skips a pair
    # the per-pass masks of pair 0, to know which vectors were replaced
    a, b = pairs[0]
    plan = T.PIVPlan(a.shape, 64, 32, 2, "CWS", 2.0, device="cuda:0")
    import torch
    _, _, m = plan.run(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
    valid = np.flip(~m[0].cpu().numpy().astype(bool), axis=0)
    for (x, y, u, v) in res:
        assert np.isfinite(u).all() and np.isfinite(v).all()
    if ref:                                  # untouched vectors are identical in both modes
        assert np.array_equal(res[0][2][valid], ref[0][2][valid])
        assert np.array_equal(res[0][0], ref[0][0])
    table = gen.statistics_table()
    want = F.statistics_table(res[0][0], res[0][1], [r[2] for r in res], [r[3] for r in res])
    for name in want:
        assert np.allclose(table[name], want[name], rtol=1e-9, atol=1e-9 * max(1.0, np.abs(want[name]).max())), name
    with pytest.raises(ValueError):
        T.OfflinePIV(statistics=True, **kw)
    gen2 = T.OfflinePIV(replace="stencil+nmt", **kw)
    assert len(list(gen2())) == 3
