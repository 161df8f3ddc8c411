"""CPU-only tests: host logic of torchpiv_b200 (geometry, spline operator, dataset, post-processing,
error behaviour) and the C-ABI surface (library loads, exports every declared symbol, argument
checks that need no GPU)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import cases
from oracle import piv_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------------------------------
# C ABI surface
# ------------------------------------------------------------------------------------------
def _header_functions():
    text = open(os.path.join(ROOT, "include", "pivb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pivb200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from torchpiv_b200 import _lib
    names = _header_functions()
    assert len(names) >= 12
    handle = _lib.lib()
    for name in names:
        assert hasattr(handle, name), f"{name} declared in include/pivb200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes prototypes out of sync with the header"
    assert handle.pivb200_version() >= 100


def test_error_strings_and_geometry_checks_without_gpu():
    from torchpiv_b200 import _lib
    L = _lib.lib()
    r, c = ctypes.c_int(), ctypes.c_int()
    assert L.pivb200_field_shape(2048, 2048, 64, 32, ctypes.byref(r), ctypes.byref(c)) == 0
    assert (r.value, c.value) == (63, 63)
    assert L.pivb200_field_shape(288, 352, 32, 8, ctypes.byref(r), ctypes.byref(c)) == 0
    assert (r.value, c.value) == (11, 14)
    assert L.pivb200_field_shape(64, 64, 32, 32, ctypes.byref(r), ctypes.byref(c)) == _lib.E_OVERLAP
    assert L.pivb200_field_shape(16, 64, 32, 0, ctypes.byref(r), ctypes.byref(c)) == _lib.E_FRAME
    assert "smaller than the window_size" in _lib.error_string(_lib.E_OVERLAP)
    assert "larger than the image" in _lib.error_string(_lib.E_FRAME)
    # argument validation happens before any CUDA call, so it is testable here
    dummy = ctypes.c_void_p(0x1000)
    for bad_window in (258, 3, 2):      # above 256 / below 4 (47 or 48 would take the general kernel)
        rc = L.pivb200_pass_first(dummy, dummy, 1, 0, 256, 256, 256, bad_window, 0, 1, 1.2, dummy, dummy, dummy,
                                  None, None)
        assert rc == _lib.E_WINDOW
    rc = L.pivb200_pass_first(dummy, dummy, 1, 0, 64, 64, 64, 48, 50, 1, 1.2, dummy, dummy, dummy, None, None)
    assert rc == _lib.E_OVERLAP
    rc = L.pivb200_pass_first(dummy, dummy, 1, 0, 64, 64, 64, 32, 40, 1, 1.2, dummy, dummy, dummy, None, None)
    assert rc == _lib.E_OVERLAP
    rc = L.pivb200_pass_first(dummy, dummy, 1, 0, 16, 16, 16, 32, 16, 1, 1.2, dummy, dummy, dummy, None, None)
    assert rc == _lib.E_FRAME
    rc = L.pivb200_pass_first(dummy, dummy, 1, 0, 64, 64, 64, 32, 16, 1, 1.2, None, dummy, dummy, None, None)
    assert rc == _lib.E_ARG
    with pytest.raises(ValueError):
        _lib.check(_lib.E_WINDOW)
    with pytest.raises(RuntimeError):
        _lib.check(_lib.E_ARG)


def test_missing_library_fails_loudly(monkeypatch):
    """No CPU / PyTorch fallback: without the .so every entry point raises."""
    from torchpiv_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libpivb200.so")
    with pytest.raises(RuntimeError, match="no CPU"):
        _lib.lib()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "torchpiv_b200")
    for name in os.listdir(pkg):
        if name.endswith(".py"):
            src = open(os.path.join(pkg, name)).read()
            assert "oracle" not in src.replace("no oracle", ""), f"{name} mentions the oracle"
    out = subprocess.run([sys.executable, "-c",
                          "import sys; sys.path.insert(0, %r); import torchpiv_b200, sys as s; "
                          "print(any(m.startswith('oracle') for m in s.modules))" % ROOT],
                         capture_output=True, text=True, check=True)
    assert out.stdout.strip() == "False"


def test_cpu_device_is_refused(tmp_path):
    import torchpiv_b200 as T
    from torchpiv_b200 import synth
    synth.write_pair_folder(str(tmp_path), [cases.small_pair(seed=1)])
    with pytest.raises(RuntimeError, match="no CPU"):
        T.OfflinePIV(str(tmp_path), "cpu", "bmp", 64, 32)
    with pytest.raises(KeyError):
        T.OfflinePIV(str(tmp_path), "no such device", "bmp", 64, 32)
    with pytest.raises(KeyError):
        T.OfflinePIV(str(tmp_path), "cpu", "bmp", 64, 32, multipass_mode="XYZ")
    empty = tmp_path / "empty"
    empty.mkdir()
    gen = T.OfflinePIV(str(empty), "cpu", "bmp", 64, 32)       # empty folder: constructed, zero pairs
    assert len(gen) == 0 and list(gen()) == []


# ------------------------------------------------------------------------------------------
# geometry / predictor operator
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,w,o", [((288, 352), 64, 32), ((2048, 2048), 32, 16), ((300, 500), 64, 48),
                                       ((1000, 1216), 16, 8), ((65, 64), 64, 0)])
def test_field_shape_and_coordinates(shape, w, o):
    from torchpiv_b200 import geometry as Gm
    assert tuple(Gm.get_field_shape(shape, w, o)) == tuple(O.get_field_shape(shape, w, o))
    x, y = Gm.get_coordinates(shape, w, o)
    ox, oy = O.get_coordinates(shape, w, o)
    assert x.dtype == np.float64 and np.array_equal(x, ox) and np.array_equal(y, oy)


def test_coordinates_match_reference_golden(golden):
    from torchpiv_b200 import geometry as Gm
    g = golden("pass1.npz")
    for w, o in cases.PASS1_GEOMS:
        x, y = Gm.get_coordinates(cases.SMALL_SHAPE, w, o)
        assert np.array_equal(x, g[f"uniform_{w}_{o}_x"]) and np.array_equal(y, g[f"uniform_{w}_{o}_y"])


@pytest.mark.parametrize("shape,g0,g1", [((288, 352), (64, 32), (32, 16)), ((2048, 2048), (64, 32), (32, 16)),
                                         ((2048, 2048), (32, 16), (16, 8)), ((300, 500), (64, 48), (32, 24)),
                                         ((288, 352), (32, 8), (16, 4)), ((512, 512), (64, 32), (42, 21))])
def test_spline_operator_equals_fitpack(shape, g0, g1):
    """Ay @ U @ Ax.T reproduces scipy's RectBivariateSpline (what the reference calls) to 1e-12,
    including the clamped evaluation outside the old grid."""
    from scipy.interpolate import RectBivariateSpline
    from torchpiv_b200.geometry import get_coordinates, spline_operator
    rng = np.random.default_rng(0)
    x0, y0 = get_coordinates(shape, *g0)
    x1, y1 = get_coordinates(shape, *g1)
    for field in (rng.normal(size=x0.shape) * 5, (rng.random(x0.shape) < 0.1).astype(np.float64)):
        ref = RectBivariateSpline(y0[:, 0], x0[0, :], field)(y1[:, 0], x1[0, :])
        got = spline_operator(y0[:, 0], y1[:, 0]) @ field @ spline_operator(x0[0, :], x1[0, :]).T
        assert np.abs(ref - got).max() < 1e-12


def test_spline_operator_rejects_tiny_fields():
    from torchpiv_b200.geometry import spline_operator
    with pytest.raises(ValueError):
        spline_operator(np.arange(3.0), np.arange(5.0))


def test_pass_schedule():
    from torchpiv_b200.engine import pass_schedule
    assert pass_schedule(64, 32, 3, 2.0) == [(64, 32), (32, 16), (16, 8)]
    assert pass_schedule(64, 32, 2, 1.5) == [(64, 32), (42, 21)]       # PB:856-857: int(w // scale)
    assert pass_schedule(64, 48, 1, 2.0) == [(64, 48)]


# ------------------------------------------------------------------------------------------
# dataset
# ------------------------------------------------------------------------------------------
def test_dataset_pairing_and_natural_sort(tmp_path):
    from torchpiv_b200 import synth
    from torchpiv_b200.dataset import PIVDataset, list_pairs, natural_keys
    rng = np.random.default_rng(0)
    names = ["img10_a.bmp", "img2_a.bmp", "img2_b.bmp", "img10_b.bmp", "img1_a.bmp", "img1_b.bmp", "note.txt"]
    imgs = {}
    for n in names:
        if n.endswith(".bmp"):
            imgs[n] = rng.integers(0, 256, (37, 53), dtype=np.uint8)     # odd width: BMP row padding
            synth.write_bmp(str(tmp_path / n), imgs[n])
        else:
            (tmp_path / n).write_text("x")
    order = sorted(imgs, key=natural_keys)
    assert order == ["img1_a.bmp", "img1_b.bmp", "img2_a.bmp", "img2_b.bmp", "img10_a.bmp", "img10_b.bmp"]
    pairs = list_pairs(str(tmp_path), "bmp", "pairs")
    assert [tuple(os.path.basename(p) for p in pr) for pr in pairs] == list(zip(order[::2], order[1::2]))
    seq = list_pairs(str(tmp_path), "bmp", "sequential")
    assert len(seq) == 5 and os.path.basename(seq[1][0]) == "img1_b.bmp"
    assert list_pairs(str(tmp_path), "bmp", "bogus") == []
    ds = PIVDataset(str(tmp_path), "bmp", "pairs")
    a, b = ds[2]
    assert np.array_equal(a, imgs["img10_a.bmp"]) and np.array_equal(b, imgs["img10_b.bmp"])
    (tmp_path / "img11_a.bmp").write_bytes(b"not an image")
    (tmp_path / "img11_b.bmp").write_bytes(b"")
    ds = PIVDataset(str(tmp_path), "bmp", "pairs")
    assert ds[3] == (None, None)                                         # unreadable -> skipped by caller


# ------------------------------------------------------------------------------------------
# post-processing
# ------------------------------------------------------------------------------------------
def test_postprocess_matches_oracle():
    from torchpiv_b200.postprocess import finalize_field
    rng = np.random.default_rng(3)
    x, y = O.get_coordinates((288, 352), 32, 16)
    for trial in range(6):
        u = rng.normal(size=x.shape) + 3
        v = rng.normal(size=x.shape) - 2
        val = rng.random(x.shape) < (0.05 if trial < 4 else 0.0)
        if trial == 1:
            val[0, :] = True                       # a whole border row invalid
        if trial == 2:
            val[3:9, 4:12] = True                  # a block
        if trial == 3:
            val[:] = rng.random(x.shape) < 0.6     # too many false vectors -> None
        ref = O.postprocess(u.copy(), v.copy(), x, y, val.copy(), scale=0.02, dt=12)
        got = finalize_field(u.copy(), v.copy(), x, y, val.copy(), scale=0.02, dt=12)
        if ref is None:
            assert got is None                      # incl. the zero-invalid-vector skip (PB:299-304)
            continue
        for r, g_ in zip(ref, got):
            assert np.array_equal(np.isnan(r), np.isnan(g_))
            assert np.allclose(r, g_, rtol=0, atol=1e-12, equal_nan=True)
    ref = O.postprocess(u.copy(), v.copy(), x, y, None, 2.0, 4.0)
    got = finalize_field(u.copy(), v.copy(), x, y, None, 2.0, 4.0)
    assert all(np.array_equal(r, g_) for r, g_ in zip(ref, got))


def test_synthetic_images_are_deterministic():
    a1, b1 = cases.small_pair(seed=3)
    a2, b2 = cases.small_pair(seed=3)
    assert np.array_equal(a1, a2) and np.array_equal(b1, b2)
    assert a1.dtype == np.uint8 and 5 < a1.mean() < 60


def test_read_gray_into_fast_bmp_path_equals_opencv(tmp_path):
    """Uncompressed 8-bit grey-palette BMPs are copied straight out of the file; everything else goes
    through OpenCV.  Both must give what the reference's decode (cv2.imdecode GRAYSCALE, PB:136-137) gives."""
    import cv2
    from torchpiv_b200 import dataset, synth
    rng = np.random.default_rng(0)
    for shape in ((37, 53), (64, 62), (128, 256)):
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        files = {}
        files["synth"] = str(tmp_path / f"s{shape[1]}.bmp")
        synth.write_bmp(files["synth"], img)
        files["cv2"] = str(tmp_path / f"c{shape[1]}.bmp")
        cv2.imwrite(files["cv2"], img)
        files["png"] = str(tmp_path / f"p{shape[1]}.png")
        cv2.imwrite(files["png"], img)
        for kind, path in files.items():
            raw = np.fromfile(path, dtype=np.uint8)
            assert (dataset._bmp8_view(raw) is not None) == (kind != "png"), kind
            dst = np.full(shape, 7, dtype=np.uint8)
            assert dataset.read_gray_into(path, dst)
            assert np.array_equal(dst, cv2.imdecode(raw, cv2.IMREAD_GRAYSCALE)) and np.array_equal(dst, img)
    # colour BMP: not the fast format, OpenCV converts
    col = rng.integers(0, 256, (20, 24, 3), dtype=np.uint8)
    path = str(tmp_path / "col.bmp")
    cv2.imwrite(path, col)
    raw = np.fromfile(path, dtype=np.uint8)
    assert dataset._bmp8_view(raw) is None
    dst = np.empty((20, 24), dtype=np.uint8)
    assert dataset.read_gray_into(path, dst) and np.array_equal(dst, cv2.imdecode(raw, cv2.IMREAD_GRAYSCALE))
    # a grey BMP whose palette is NOT the identity ramp must not take the fast path
    raw = np.fromfile(files["synth"], dtype=np.uint8).copy()
    raw[54 + 4 * 10] = 200
    assert dataset._bmp8_view(raw) is None
    # truncated / missing / wrong shape
    trunc = str(tmp_path / "t.bmp")
    open(trunc, "wb").write(np.fromfile(files["synth"], dtype=np.uint8)[:2000].tobytes())
    assert not dataset.read_gray_into(trunc, np.empty((128, 256), dtype=np.uint8))
    assert not dataset.read_gray_into(str(tmp_path / "missing.bmp"), dst)
    assert not dataset.read_gray_into(files["synth"], np.empty((5, 5), dtype=np.uint8))


def test_hole_fill_pool_matches_inline_fill():
    """The worker-process post-processing (shared-memory slots, forked workers) returns exactly what the in-thread
    path returns, including pairs the reference would skip (no invalid vector at all -> None)."""
    from torchpiv_b200 import postprocess as P
    rng = np.random.default_rng(3)
    n, m, B = 37, 41, 5
    u, v = rng.normal(size=(B, n, m)), rng.normal(size=(B, n, m))
    inv = rng.random((B, n, m)) < 0.01
    inv[:, 10:15, 20:27] = True
    inv[3] = False                                  # the reference skips such a pair (PB:299-304)
    ref = [P.finalize_uv(u[i].copy(), v[i].copy(), inv[i], 0.5, 2.0) for i in range(B)]
    pool = P.HoleFillPool(2, B, n, m)
    try:
        handles = [pool.submit_batch(u, v, inv, 0.5, 2.0) for _ in range(pool.capacity)]
        for h in handles:
            got = pool.collect(h)
            assert [g is None for g in got] == [r is None for r in ref] and ref[3] is None
            for g, r in zip(got, ref):
                if r is not None:
                    assert np.array_equal(g[0], r[0]) and np.array_equal(g[1], r[1])
    finally:
        pool.shutdown()
