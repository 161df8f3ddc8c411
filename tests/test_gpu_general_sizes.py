"""GPU tests of the general-size path (csrc/generic_pass.cuh): interrogation windows that are not
16/32/64 px -- e.g. 48 px, 128 px, or the 42 / 28 px that multipass_scale = 1.5 produces -- against
the golden vectors of the unmodified reference (tests/golden/general_sizes.npz) and the oracle.
Same bars as tests/test_gpu_parity.py: windows bit-exact (this path evaluates the CWS taps in the
reference's own four-term order), masks identical on well-conditioned vectors, displacements within
1e-3 px."""
import numpy as np
import pytest
import torch

import cases
from oracle import piv_oracle as O
from test_gpu_parity import check_field

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import gpu_util
    return gpu_util


@pytest.fixture(scope="module")
def T():
    import torchpiv_b200
    return torchpiv_b200


@pytest.mark.parametrize("w,o", cases.GENERAL_GEOMS)
def test_windows_any_size_bit_exact(G, w, o):
    a, b = cases.small_pair(seed=3, kind="vortex", zero_patch=True)
    wa, wb = G.windows(a, b, w, o)
    assert np.array_equal(wa, O.moving_window_array(a, w, o).astype(np.float32))
    assert np.array_equal(wb, O.moving_window_array(b, w, o).astype(np.float32))
    n = wa.shape[0]
    rng = np.random.default_rng(w)
    vx = rng.uniform(-12, 12, n).astype(np.float32)
    vy = rng.uniform(-12, 12, n).astype(np.float32)
    vx[::7] = np.rint(vx[::7])              # exact-integer coordinates (PB:170, 193)
    vy[::5] = 400.0                          # far outside the frame: flat-index clamp
    idx = O.window_index_grid(a.shape, w, o)
    wa, wb = G.windows(a, b, w, o, "CWS", vx, vy)
    assert np.array_equal(wa, O.bilinear_interpolation_cws(a, idx, -vx[:, None, None], -vy[:, None, None]))
    assert np.array_equal(wb, O.bilinear_interpolation_cws(b, idx, vx[:, None, None], vy[:, None, None]))
    ix, iy = np.rint(vx).astype(np.int64), np.rint(vy).astype(np.int64)
    wa, wb = G.windows(a, b, w, o, "DWS", ix, iy)
    assert np.array_equal(wa, O.interpolation_dws(a, idx, -ix[:, None, None], -iy[:, None, None]).astype(np.float32))
    assert np.array_equal(wb, O.interpolation_dws(b, idx, ix[:, None, None], iy[:, None, None]).astype(np.float32))


@pytest.mark.parametrize("w", [48, 24, 128, 20, 6, 42, 28, 66, 104, 100, 160, 34, 4, 192, 200, 256, 25, 35, 31, 9, 5, 63, 165, 131, 37])   # 34 = 2 x 17 and 31: direct sums; > 160: global scratch; odd: [w, w-1]
def test_correlate_any_size(T, w):
    rng = np.random.default_rng(w)
    a = rng.integers(0, 256, (5, w, w), dtype=np.uint8)
    b = np.roll(a, (2, -1), axis=(1, 2))
    ref = O.correlate_fft(a.astype(np.float64), b.astype(np.float64))
    for arr_a, arr_b in ((a, b), (a.astype(np.float32), b.astype(np.float32))):
        got = T.correalte_fft(torch.from_numpy(arr_a).cuda(), torch.from_numpy(arr_b).cuda())
        assert got.dtype == torch.float32 and tuple(got.shape) == ref.shape
        got = got.cpu().numpy().astype(np.float64)
        # sizes with a prime factor above 13 take FP32 direct sums (O(w) terms per bin instead of O(log w) stages)
        n, big = w, 1
        for q in range(2, w + 1):
            while n % q == 0:
                n, big = n // q, q
        assert np.abs(got - ref).max() <= (5e-6 if big > 13 else 2e-6) * np.abs(ref).max()
        assert np.array_equal(got.reshape(5, -1).argmax(1), ref.reshape(5, -1).argmax(1))


@pytest.mark.parametrize("w,o", cases.GENERAL_GEOMS)
def test_pass_first_any_size_vs_reference(T, golden, w, o):
    g = golden("general_sizes.npz")
    a, b = cases.small_pair(seed=3, kind="vortex", zero_patch=True)
    assert cases.sha(a, b) == str(g["sha"])
    u, v, x, y, m = T.extended_search_area_piv(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(),
                                               window_size=w, overlap=o, validate=True)
    assert np.array_equal(x, g[f"p1_{w}_{o}_x"]) and np.array_equal(y, g[f"p1_{w}_{o}_y"])
    stash = {}
    O.extended_search_area_piv(a, b, w, o, validate=True, stash=stash)
    # odd windows: the reference's [w, w-1] maps are a warped correlation with flatter peaks -- a few more vectors are
    # ill-conditioned by the oracle-side rule (35 px: 4 of the 226 valid ones)
    check_field(u, v, m, g[f"p1_{w}_{o}_u"], g[f"p1_{w}_{o}_v"], g[f"p1_{w}_{o}_mask"], stash["corr"],
                max_ill_valid=0.05 if w % 2 else 0.01)


@pytest.mark.parametrize("mode", ["CWS", "DWS"])
def test_scale_1p5_chain_at_the_function_boundary(T, golden, mode):
    """64 -> 42 -> 28 px, each pass fed with the REFERENCE's previous field."""
    g = golden("general_sizes.npz")
    a, b = cases.small_pair(seed=3, kind="vortex", zero_patch=True)
    fa, fb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    x, y = O.get_coordinates(a.shape, 64, 32)
    w, o = 64, 32
    for it in (1, 2):
        prev = [g[f"{mode}_p{it - 1}_{k}"] for k in ("u", "v", "mask")]
        w, o = int(w // 1.5), int(o // 1.5)
        fn = T.IterModMap.functions[mode](a.shape, w, o, "cuda:0")
        u, v, x1, y1, m = fn(fa, fb, x, y, prev[0].copy(), prev[1].copy(), prev[2].copy())
        assert np.array_equal(x1, g[f"{mode}_p{it}_x"]) and np.array_equal(y1, g[f"{mode}_p{it}_y"])
        orc = O.ITER_MODES[mode](a.shape, w, o)
        orc(a, b, x, y, prev[0].copy(), prev[1].copy(), prev[2].copy())
        # the noise / blank / black patches of this pair cover a larger share of the coarse 42 px grid
        check_field(u, v, m, g[f"{mode}_p{it}_u"], g[f"{mode}_p{it}_v"], g[f"{mode}_p{it}_mask"], orc.last_corr,
                    max_ill=0.2)
        x, y = x1, y1


@pytest.mark.parametrize("mode", ["CWS", "DWS"])
def test_halving_into_an_odd_window(T, golden, mode):
    """50 -> 25 px: the second pass runs on ODD windows ([25, 24] correlation maps like the reference's), fed with the
    REFERENCE's first-pass field."""
    g = golden("general_sizes.npz")
    a, b = cases.small_pair(seed=3, kind="vortex", zero_patch=True)
    fa, fb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    x, y = O.get_coordinates(a.shape, 50, 25)
    prev = [g[f"odd{mode}_p0_{k}"] for k in ("u", "v", "mask")]
    fn = T.IterModMap.functions[mode](a.shape, 25, 12, "cuda:0")
    u, v, x1, y1, m = fn(fa, fb, x, y, prev[0].copy(), prev[1].copy(), prev[2].copy())
    assert np.array_equal(x1, g[f"odd{mode}_p1_x"]) and np.array_equal(y1, g[f"odd{mode}_p1_y"])
    orc = O.ITER_MODES[mode](a.shape, 25, 12)
    orc(a, b, x, y, prev[0].copy(), prev[1].copy(), prev[2].copy())
    check_field(u, v, m, g[f"odd{mode}_p1_u"], g[f"odd{mode}_p1_v"], g[f"odd{mode}_p1_mask"], orc.last_corr, max_ill=0.2,
                max_ill_valid=0.05)


@pytest.mark.parametrize("mode", ["CWS", "DWS"])
def test_plan_with_an_odd_pass_vs_oracle(T, mode):
    """Device-resident 100 -> 50 -> 25 px plan (general kernel on every pass, the last one on odd windows) vs the
    oracle's chained passes."""
    a, b = cases.small_pair(seed=7, kind="vortex")
    plan = T.PIVPlan(a.shape, 100, 50, 3, mode, 2.0, device="cuda:0")
    assert [g.wind for g in plan.passes] == [100, 50, 25]
    plan.run(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
    ou, ov, _, _, om, hist = O.piv_passes(a, b, 100, 50, 3, mode)
    for k, (hu, hv, hm) in enumerate(hist):
        du, dv, dm = (t[0].cpu().numpy() for t in plan.pass_results(k, 1))
        assert du.shape == hu.shape
        assert (dm.astype(bool) == hm).mean() > 0.95
        assert np.quantile(np.abs(du - hu), 0.9) < 1e-4 and np.quantile(np.abs(dv - hv), 0.9) < 1e-4, (mode, k)


def test_offline_piv_with_scale_1p5(T, golden, tmp_path):
    from torchpiv_b200 import synth
    g = golden("general_sizes.npz")
    pairs = [cases.small_pair(seed=10 + i, kind="uniform" if i % 2 == 0 else "vortex") for i in range(2)]
    synth.write_pair_folder(str(tmp_path), pairs)
    gen = T.OfflinePIV(folder=str(tmp_path), device="cuda:0", file_fmt="bmp", wind_size=64, overlap=32,
                       multipass=2, multipass_mode="CWS", multipass_scale=1.5, dt=12, scale=0.02)
    res = list(gen())
    assert (len(gen), len(res)) == tuple(g["offline_n"])
    k = 0.02 / 12 * 1000
    for i, (x, y, u, v) in enumerate(res):
        assert np.array_equal(x, g[f"offline_{i}_x"]) and np.array_equal(y, g[f"offline_{i}_y"])
        du, dv = np.abs(u - g[f"offline_{i}_u"]) / k, np.abs(v - g[f"offline_{i}_v"]) / k
        assert np.quantile(du, 0.95) < 1e-4 and np.quantile(dv, 0.95) < 1e-4
        assert np.quantile(du, 0.995) < 5e-2 and np.quantile(dv, 0.995) < 5e-2


def test_mixed_plan_full_size(T):
    """4 MP, 48 -> 24 px DWS (both passes on the general path): imposed displacement recovered."""
    from torchpiv_b200 import synth
    shape = (2048, 2048)
    a, b = synth.particle_pair(shape, synth.uniform_shift(3.3, -2.2), seed=5)
    plan = T.PIVPlan(shape, 48, 24, 2, "DWS", 2.0, device="cuda:0")
    u, v, m = plan.run(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
    ok = ~m[0].bool()
    assert ok.float().mean() > 0.95
    assert abs(float(u[0][ok].median()) - 3.3) < 0.1 and abs(float(v[0][ok].median()) + 2.2) < 0.1


def test_general_path_batches_pitch_and_no_validation(T):
    """Batch of pairs, frames that are strided views (pitch > width), and validate=False all give what a
    single contiguous pair gives."""
    a, b = cases.small_pair(seed=3, kind="vortex", zero_patch=True)
    fa, fb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    plan = T.PIVPlan(a.shape, 48, 24, 2, "CWS", 2.0, device="cuda:0")          # 48 -> 24 px
    u1, v1, m1 = (t.clone() for t in plan.run(fa, fb))
    wide_a = torch.zeros((3, a.shape[0], a.shape[1] + 16), dtype=torch.uint8, device="cuda")
    wide_b = torch.zeros_like(wide_a)
    wide_a[:, :, :a.shape[1]] = torch.stack([fa, fb, fa])
    wide_b[:, :, :a.shape[1]] = torch.stack([fb, fa, fb])
    u3, v3, m3 = plan.run(wide_a[:, :, :a.shape[1]], wide_b[:, :, :a.shape[1]])
    for i in (0, 2):
        assert torch.equal(u3[i], u1[0]) and torch.equal(v3[i], v1[0]) and torch.equal(m3[i], m1[0])
    un, vn, mn = plan.run(fa, fb, validate=False)
    # without validation there is no mask-driven replacement, but vectors the first pass keeps are the same
    assert torch.isfinite(un).all() and torch.isfinite(vn).all()
    u, v, x, y, m = T.extended_search_area_piv(fa, fb, window_size=48, overlap=24, validate=False)
    assert m is None
    ou, ov, _, _, _ = O.extended_search_area_piv(a, b, 48, 24, validate=False)
    assert np.quantile(np.abs(u - ou), 0.9) < 2e-5 and np.quantile(np.abs(v - ov), 0.9) < 2e-5
