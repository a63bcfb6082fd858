"""Geometry fuzz for the pair kernels against the oracle (run with -m gpu).

The fixed fixtures are uniform random fields; the fast kernels, however, depend on geometry (Hilbert runs, tile
boxes, lag windows, groups of 8 pairs sharing one lag lookup, padding).  These cases push every kernel variant
(default choice, 16- and 32-slot ring kernels, full-histogram kernel) through awkward inputs the oracle can
still check: integer grids (many distances EXACTLY on a lag edge), clustered points, UTM-sized coordinate
offsets, a thin strip, duplicated coordinates, tiny coordinates, ragged sizes.  Bar as everywhere: counts
bit-exact, sums within 1e-9 relative.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
RTOL = 1e-9

VARIANTS = {
    "default": {},
    "ring16": {"PIB_FORCE_RING": "1", "PIB_RING": "16"},
    "ring32": {"PIB_FORCE_RING": "1", "PIB_RING": "32"},
    "ring32_four_lags": {"PIB_FORCE_RING": "1", "PIB_RING": "32", "PIB_LAGS_PER_GROUP": "4"},
    "ring16_four_lags": {"PIB_FORCE_RING": "1", "PIB_RING": "16", "PIB_LAGS_PER_GROUP": "4"},
    "ring16_v5": {"PIB_FORCE_RING": "1", "PIB_RING": "16", "PIB_PAIR_V": "5"},
    "full": {"PIB_FORCE_FULL": "1"},
    "ring16_pairwise": {"PIB_FORCE_RING": "1", "PIB_RING": "16", "PIB_GROUP_PATH": "0", "PIB_BALANCE": "0"},
    # two A points per thread, 10-slot sliding ring (the BASELINE config 2 kernel), forced onto every geometry: spans wider
    # than the ring take the per-window sweeps; with and without the high-word fast path (PIB_HIW=0: exact stages only)
    "apt2_ring10": {"PIB_FORCE_RING": "1", "PIB_V6": "2,8,512,10", "PIB_LAGS_PER_GROUP": "2"},
    # the same with the fp32 screening path switched off (high-word fast path instead)
    "apt2_ring10_hiw": {"PIB_FORCE_RING": "1", "PIB_V6": "2,8,512,10", "PIB_LAGS_PER_GROUP": "2", "PIB_F32": "0"},
    "apt2_ring10_exact": {"PIB_FORCE_RING": "1", "PIB_V6": "2,8,512,10", "PIB_LAGS_PER_GROUP": "2", "PIB_HIW": "0", "PIB_F32": "0"},
    "ring16_exact": {"PIB_FORCE_RING": "1", "PIB_RING": "16", "PIB_HIW": "0"},
}


def _case(name):
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
    if name == "grid":                       # distances 1, sqrt(2), 2, ..., 5 = hypot(3, 4): exact edge hits
        gx, gy = np.meshgrid(np.arange(70.0), np.arange(67.0))
        xy = np.column_stack([gx.ravel(), gy.ravel()])
        lags = np.arange(1.0, 21.0, 1.0)
    elif name == "grid_offset":              # the same grid shifted by a non-representable offset and scaled
        gx, gy = np.meshgrid(np.arange(64.0), np.arange(50.0))
        xy = np.column_stack([gx.ravel(), gy.ravel()]) * 0.1 + np.array([1234.567, -987.654])
        lags = np.arange(0.1, 3.05, 0.1)
    elif name == "clustered":
        centres = rng.uniform(0, 50_000, size=(12, 2))
        xy = np.vstack([c + rng.normal(0, 300.0, size=(500, 2)) for c in centres] + [rng.uniform(0, 50_000, size=(513, 2))])
        lags = np.arange(250.0, 20_000.0, 250.0)
    elif name == "utm":
        xy = rng.uniform(0, 30_000, size=(8191, 2)) + np.array([512_345.25, 4_198_765.5])
        lags = np.arange(300.0, 15_000.0, 300.0)
    elif name == "strip":
        xy = np.column_stack([rng.uniform(0, 100_000, 10_007), rng.uniform(0, 1_000, 10_007)])
        lags = np.arange(400.0, 30_000.0, 400.0)
    elif name == "duplicates":
        base = rng.uniform(0, 20_000, size=(3000, 2))
        xy = np.vstack([base, base[rng.choice(3000, 500, replace=False)]])
        lags = np.arange(200.0, 9_000.0, 200.0)
    elif name == "tiny":
        xy = rng.uniform(-5e-4, 5e-4, size=(4099, 2))
        lags = np.arange(2e-5, 4.1e-4, 2e-5)
    elif name == "custom_bins":              # irregular lag widths (custom_bins), wide and narrow mixed
        xy = rng.uniform(0, 40_000, size=(6001, 2))
        lags = np.array([50.0, 75.0, 400.0, 410.0, 2_000.0, 2_001.0, 9_000.0, 15_000.0, 15_500.0, 30_000.0])
    elif name in ("z_offset", "z_heavy_tail"):  # values that stress the centred / expanded sums of the ring kernel
        xy = rng.uniform(0, 40_000, size=(7001, 2))
        lags = np.arange(400.0, 16_000.0, 400.0)
    else:
        raise KeyError(name)
    z = 100.0 + 10.0 * np.sin(xy[:, 0] / (np.ptp(xy[:, 0]) + 1e-300) * 6.0) + rng.normal(0, 1.0, len(xy))
    if name == "z_offset":
        z = 1.0e9 + (z - 100.0)                 # huge mean, small spread: z - mean is exact, z^2 would not be
    elif name == "z_heavy_tail":
        z = np.exp(rng.normal(0.0, 2.5, len(xy)))   # log-normal over five orders of magnitude
    return np.column_stack([xy, z]), lags


CASES = ["grid", "grid_offset", "clustered", "utm", "strip", "duplicates", "tiny", "custom_bins", "z_offset", "z_heavy_tail"]


@pytest.fixture(scope="module")
def lib():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pyinterpolate_b200 import _lib
    return _lib


def _run(lib, monkeypatch, variant, *args, **kw):
    for k, v in VARIANTS[variant].items():
        monkeypatch.setenv(k, v)
    try:
        return lib.variogram_host(*args, **kw)
    finally:
        for k in VARIANTS[variant]:
            monkeypatch.delenv(k)


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("name", CASES)
def test_omni_fuzz(lib, oracle, monkeypatch, name, variant):
    pts, lags = _case(name)
    want = oracle.variogram_sums(pts, lags)
    got = _run(lib, monkeypatch, variant, pts, lags)
    np.testing.assert_array_equal(got["count"], want["count"])
    scale = np.abs(pts[:, 2]).max() ** 2 * np.maximum(want["count"], 1)
    np.testing.assert_allclose(got["sum_sq"], want["sum_sq"], rtol=RTOL, atol=0)
    np.testing.assert_allclose(got["sum_val"], want["sum_val"], rtol=RTOL, atol=0)
    np.testing.assert_allclose(got["sum_prod"], want["sum_prod"], rtol=0, atol=RTOL * scale.max())
    assert want["count"].sum() > 0


@pytest.mark.parametrize("variant", ["default", "ring16", "ring32", "full", "apt2_ring10"])
@pytest.mark.parametrize("name", ["grid", "clustered", "utm", "strip"])
def test_ellipse_fuzz(lib, oracle, monkeypatch, name, variant):
    from pyinterpolate_b200 import core
    pts, lags = _case(name)
    if name != "grid":
        lags = lags[::4] if len(lags) > 40 else lags[::2]       # an empty lag raises in the public layer; keep them coarse
    angles, tol = [0, 37, 90, 135], 0.3
    want = oracle.variogram_sums(pts, lags, angles, tol)
    W = np.stack([core.define_whitening_matrix(a, tol) for a in angles]).reshape(-1, 4)
    got = _run(lib, monkeypatch, variant, pts, lags, lower=core.ellipse_lower_limits(lags), W=W)
    np.testing.assert_array_equal(got["count"], want["count"])
    np.testing.assert_allclose(got["sum_sq"], want["sum_sq"], rtol=RTOL, atol=0)
    np.testing.assert_allclose(got["sum_val"], want["sum_val"], rtol=RTOL, atol=0)


@pytest.mark.parametrize("name", ["clustered", "strip"])
def test_sharded_ring_fuzz(lib, oracle, monkeypatch, name):
    """5 ranks' slices of the cost-ordered item list add up to the whole (counts exactly)."""
    pts, lags = _case(name)
    want = oracle.variogram_sums(pts, lags)
    parts = [_run(lib, monkeypatch, "ring16", pts, lags, shard=(r, 5)) for r in range(5)]
    np.testing.assert_array_equal(sum(p["count"] for p in parts), want["count"])
    np.testing.assert_allclose(sum(p["sum_sq"] for p in parts), want["sum_sq"], rtol=RTOL)
    whole = _run(lib, monkeypatch, "ring16", pts, lags)
    assert sum(p["pairs_evaluated"] for p in parts) == whole["pairs_evaluated"]


# ---------------------------------------------------------------------------------------------
# kriging: neighbour search on a uniform grid + solver, on geometry the grid index does not like
# ---------------------------------------------------------------------------------------------
def _krige_case(name):
    rng = np.random.default_rng(sum(map(ord, name)) + 7)
    if name == "clustered":                  # dense blobs and empty space: ring expansion over many empty cells
        centres = rng.uniform(0, 50_000, size=(10, 2))
        xy = np.vstack([c + rng.normal(0, 200.0, size=(400, 2)) for c in centres])
        unk = np.vstack([rng.uniform(-20_000, 70_000, size=(150, 2)), centres + 5.0])     # far outside the hull too
        rang = 8_000.0
    elif name == "grid":                     # exact distance ties everywhere: the (distance, index) order decides
        gx, gy = np.meshgrid(np.arange(40.0), np.arange(35.0))
        xy = np.column_stack([gx.ravel(), gy.ravel()])
        unk = np.vstack([rng.uniform(0, 39, size=(100, 2)), np.column_stack([np.arange(20) + 0.5, np.arange(20) + 0.5])])
        rang = 12.0
    elif name == "utm":
        xy = rng.uniform(0, 30_000, size=(5000, 2)) + np.array([512_345.25, 4_198_765.5])
        unk = rng.uniform(0, 30_000, size=(200, 2)) + np.array([512_345.25, 4_198_765.5])
        rang = 6_000.0
    elif name == "strip":
        xy = np.column_stack([rng.uniform(0, 100_000, 6000), rng.uniform(0, 500, 6000)])
        unk = np.column_stack([rng.uniform(0, 100_000, 200), rng.uniform(-2_000, 2_500, 200)])
        rang = 5_000.0
    else:
        raise KeyError(name)
    z = 50.0 + 5.0 * np.sin(xy[:, 0] / (np.ptp(xy[:, 0]) + 1e-300) * 5.0) + rng.normal(0, 0.5, len(xy))
    return np.column_stack([xy, z]), unk, rang


@pytest.mark.parametrize("k", [5, 16, 33, 64, 100])
@pytest.mark.parametrize("name", ["clustered", "grid", "utm", "strip"])
def test_kriging_fuzz(lib, oracle, name, k):
    known, unk, rang = _krige_case(name)
    var = float(np.var(known[:, 2]))
    for mode, mid in (("ok", lib.KRIGE_ORDINARY), ("sk", lib.KRIGE_SIMPLE)):
        pm = float(known[:, 2].mean())
        want, widx, wst = oracle.krige(known, unk, "exponential", 0.05 * var, var, rang, k, mode, pm)
        got, gidx, gst = lib.krige_host(known, unk, 2, [[0.05 * var, var, rang]], mid, k, process_mean=pm,
                                        want_neighbors=True)
        np.testing.assert_array_equal(gidx, widx)                       # same (distance, index) order as the oracle
        np.testing.assert_array_equal(gst[0], wst)
        ok = wst == 0
        np.testing.assert_allclose(got[0][ok, :2], want[ok, :2], rtol=1e-8 if k >= 64 else RTOL, atol=1e-9 * var)


@pytest.mark.parametrize("values", ["offset", "heavy_tail"])
def test_kriging_value_stress(lib, oracle, values):
    """Values with a huge offset / five orders of magnitude of spread through the neighbour search and the solver."""
    known, unk, rang = _krige_case("utm")
    rng = np.random.default_rng(11)
    if values == "offset":
        known[:, 2] = 1.0e9 + (known[:, 2] - 50.0)
    else:
        known[:, 2] = np.exp(rng.normal(0.0, 2.5, len(known)))
    var = float(np.var(known[:, 2]))
    scale = np.abs(known[:, 2]).max()
    for mode, mid in (("ok", lib.KRIGE_ORDINARY), ("sk", lib.KRIGE_SIMPLE)):
        pm = float(known[:, 2].mean())
        want, widx, wst = oracle.krige(known, unk, "spherical", 0.1 * var, var, rang, 24, mode, pm)
        got, gidx, gst = lib.krige_host(known, unk, 6, [[0.1 * var, var, rang]], mid, 24, process_mean=pm, want_neighbors=True)
        np.testing.assert_array_equal(gidx, widx)
        np.testing.assert_array_equal(gst[0], wst)
        ok = wst == 0
        np.testing.assert_allclose(got[0][ok, 0], want[ok, 0], rtol=0, atol=1e-9 * scale)
        np.testing.assert_allclose(got[0][ok, 1], want[ok, 1], rtol=1e-8, atol=1e-9 * var)
