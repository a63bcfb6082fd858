"""GPU parity tests for the experimental variogram path (run with -m gpu on the B200 box).

Every call goes through the public API -> ctypes -> the C ABI of libpyinterp_b200.so.
Bar: pair counts bit-exact; semivariances within 1e-9 relative; covariances within
1e-9 * (mean(z)^2 + var(z)) absolute (a covariance is the difference of two terms of that size).
"""
import numpy as np
import pytest

from pyinterpolate_b200.synthetic import dem_like_field

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.fixture(scope="module")
def pib():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pyinterpolate_b200 as pib
    return pib


def _cov_scale(points):
    z = np.asarray(points, dtype=float)[:, -1]
    return float(np.mean(z) ** 2 + np.var(z))


def _check(got, want, cov_scale=None):
    assert got.shape == want.shape
    np.testing.assert_array_equal(got[:, 0], want[:, 0])
    np.testing.assert_array_equal(got[:, 2], want[:, 2])                    # counts bit-exact (nan == nan)
    if cov_scale is None:
        np.testing.assert_allclose(got[:, 1], want[:, 1], rtol=RTOL, atol=0, equal_nan=True)
    else:
        np.testing.assert_allclose(got[:, 1], want[:, 1], rtol=0, atol=RTOL * cov_scale, equal_nan=True)


def test_line13_golden(pib, golden):
    pts = golden["line13_points"]
    _check(pib.calculate_semivariance(pts, 1, 4), golden["line13_sem"])
    _check(pib.calculate_covariance(pts, 1, 4), golden["line13_cov"], _cov_scale(pts))


def test_armstrong_golden(pib, golden):
    pts = golden["armstrong_points"]
    _check(pib.calculate_semivariance(pts, 1, 6), golden["armstrong_omni_sem"])
    _check(pib.calculate_covariance(pts, 1, 6), golden["armstrong_omni_cov"], _cov_scale(pts))
    kw = dict(direction=15, tolerance=0.25, dir_neighbors_selection_method="e")
    _check(pib.calculate_semivariance(pts, 1.5, 6, **kw), golden["armstrong_ell_sem"])
    _check(pib.calculate_covariance(pts, 1.5, 6, **kw), golden["armstrong_ell_cov"], _cov_scale(pts))
    bins = golden["armstrong_custom_bins"]
    _check(pib.calculate_semivariance(pts, custom_bins=bins), golden["armstrong_custom_sem"])
    _check(pib.calculate_covariance(pts, custom_bins=bins), golden["armstrong_custom_cov"], _cov_scale(pts))
    _check(pib.calculate_semivariance(pts, custom_bins=golden["armstrong_custom_ell_bins"], direction=45,
                                      tolerance=0.5, dir_neighbors_selection_method="ellipse"),
           golden["armstrong_custom_ell_sem"])
    with pytest.raises(RuntimeError, match="There are no neighbors for a lag"):
        pib.calculate_semivariance(pts, custom_bins=bins, direction=45, tolerance=0.5,
                                   dir_neighbors_selection_method="e")


def test_experimental_variogram_class_golden(pib, golden):
    ev = pib.ExperimentalVariogram(golden["armstrong_points"], step_size=1, max_range=6)
    np.testing.assert_array_equal(ev.lags, golden["armstrong_ev_lags"])
    np.testing.assert_array_equal(ev.points_per_lag, golden["armstrong_ev_ppl"])
    np.testing.assert_allclose(ev.semivariances, golden["armstrong_ev_sem"], rtol=RTOL)
    np.testing.assert_allclose(ev.covariances, golden["armstrong_ev_cov"], rtol=0,
                               atol=RTOL * _cov_scale(golden["armstrong_points"]))
    assert ev.variance == float(golden["armstrong_ev_variance"])


def test_sparse_empty_lags_golden(pib, golden):
    pts = dem_like_field(40, extent=1000.0, seed=11)
    _check(pib.calculate_semivariance(pts, 2.0, 60.0), golden["sparse_sem"])
    _check(pib.calculate_covariance(pts, 2.0, 60.0), golden["sparse_cov"], _cov_scale(pts))


@pytest.mark.parametrize("name,n,seed,extent,step,maxr", [
    ("omni2k", 2000, 21, 100_000.0, 500, 40000),
    ("omni10k", 10000, 0, 100_000.0, 500, 40000),         # BASELINE config 1
    ("omni600", 600, 23, 10.0, 0.1, 3.05),
])
def test_synthetic_omni_golden(pib, golden, name, n, seed, extent, step, maxr):
    pts = dem_like_field(n, extent=extent, seed=seed)
    ev = pib.ExperimentalVariogram(pts, step_size=step, max_range=maxr)
    _check(np.column_stack([ev.lags, ev.semivariances, ev.points_per_lag]), golden[f"{name}_sem"])
    _check(np.column_stack([ev.lags, ev.covariances, ev.points_per_lag]), golden[f"{name}_cov"], _cov_scale(pts))


def test_synthetic_ellipse_golden(pib, golden):
    pts = dem_like_field(1500, seed=22)
    for ang in (90, 0, 45, 135, 15):
        kw = dict(direction=ang, tolerance=0.2, dir_neighbors_selection_method="e")
        _check(pib.calculate_semivariance(pts, 2500, 40000, **kw), golden[f"ell1500_dir{ang}_sem"])
        _check(pib.calculate_covariance(pts, 2500, 40000, **kw), golden[f"ell1500_dir{ang}_cov"], _cov_scale(pts))
    pts = dem_like_field(600, extent=10.0, seed=23)
    kw = dict(direction=30, tolerance=0.35, dir_neighbors_selection_method="e")
    _check(pib.calculate_semivariance(pts, 0.1, 3.05, **kw), golden["ell600_sem"])
    _check(pib.calculate_covariance(pts, 0.1, 3.05, **kw), golden["ell600_cov"], _cov_scale(pts))


def test_directional_variogram_fused_matches_golden(pib, golden):
    pts = dem_like_field(1500, seed=22)
    dv = pib.DirectionalVariogram(pts, step_size=2500, max_range=40000, tolerance=0.2,
                                  dir_neighbors_selection_method="e")
    for name, ang in (("NS", 90), ("WE", 0), ("NE-SW", 45), ("NW-SE", 135)):
        ev = dv.get(name)
        _check(np.column_stack([ev.lags, ev.semivariances, ev.points_per_lag]), golden[f"ell1500_dir{ang}_sem"])
        assert ev.direction == ang
    assert dv.get("ISO").direction is None


@pytest.mark.parametrize("n,seed,step,maxr", [(3000, 3, 500, 40500), (7777, 4, 750, 30000), (129, 5, 5000, 60000)])
def test_omni_vs_oracle(pib, oracle, n, seed, step, maxr):
    """Seeded inputs the golden file does not hold; ragged sizes (not multiples of the 128-point tile)."""
    pts = dem_like_field(n, seed=seed)
    ev = pib.ExperimentalVariogram(pts, step_size=step, max_range=maxr)
    lags = oracle.get_lags(step, maxr)
    sem, cov, npairs = oracle.finalize(oracle.variogram_sums(pts, lags))
    np.testing.assert_array_equal(ev.points_per_lag, npairs)
    np.testing.assert_allclose(ev.semivariances, sem, rtol=RTOL, equal_nan=True)
    np.testing.assert_allclose(ev.covariances, cov, rtol=0, atol=RTOL * _cov_scale(pts), equal_nan=True)


def test_ellipse_vs_oracle_4_directions(pib, oracle):
    pts = dem_like_field(6000, seed=9)
    lags = oracle.get_lags(1000, 40500)
    want = oracle.variogram_sums(pts, lags, [90, 0, 45, 135], 0.2)
    sem, cov, npairs = oracle.finalize(want)
    dv = pib.DirectionalVariogram(pts, step_size=1000, max_range=40500, tolerance=0.2,
                                  dir_neighbors_selection_method="e")
    for t, name in enumerate(("NS", "WE", "NE-SW", "NW-SE")):
        ev = dv.get(name)
        np.testing.assert_array_equal(ev.points_per_lag, npairs[t])
        np.testing.assert_allclose(ev.semivariances, sem[t], rtol=RTOL)
        np.testing.assert_allclose(ev.covariances, cov[t], rtol=0, atol=RTOL * _cov_scale(pts))


def test_generic_kernel_matches_tiled_kernel(pib, monkeypatch):
    """The fallback kernel (many lags / irregular ellipse limits) and the tiled kernel agree."""
    from pyinterpolate_b200 import _lib
    pts = dem_like_field(2500, seed=12)
    lags = np.arange(500.0, 40000.0, 500.0)
    fast = _lib.variogram_host(pts, lags)
    monkeypatch.setenv("PIB_FORCE_GENERIC", "1")
    slow = _lib.variogram_host(pts, lags)
    monkeypatch.delenv("PIB_FORCE_GENERIC")
    np.testing.assert_array_equal(fast["count"], slow["count"])
    for key in ("sum_sq", "sum_prod", "sum_val"):
        np.testing.assert_allclose(fast[key], slow[key], rtol=1e-12)
    # more lags than the shared-memory histograms hold at 128 threads: narrower CTAs / generic kernel
    for n_lags in (100, 200, 400):
        many = np.linspace(100.0, 40000.0, n_lags)
        a = _lib.variogram_host(pts, many)
        monkeypatch.setenv("PIB_FORCE_GENERIC", "1")
        b = _lib.variogram_host(pts, many)
        monkeypatch.delenv("PIB_FORCE_GENERIC")
        np.testing.assert_array_equal(a["count"], b["count"])
        np.testing.assert_allclose(a["sum_sq"], b["sum_sq"], rtol=1e-12)


@pytest.mark.parametrize("version", ["6", "5"])
@pytest.mark.parametrize("lags_per_group", ["2", "4"])
@pytest.mark.parametrize("ring", ["32", "16"])
@pytest.mark.parametrize("n,step,maxr", [(30000, 2000.0, 40500.0), (5000, 250.0, 30000.0), (1000, 500.0, 40500.0)])
def test_ring_kernel_vs_oracle(pib, oracle, monkeypatch, n, step, maxr, ring, lags_per_group, version):
    """The ring-window kernels (the 1M-point fast path) forced onto sizes the oracle can check:
    coarse lags (ring slots only), fine lags (tile pairs wider than the ring -> overflow path); the two-lag and the
    four-lag group scheme of pair_ring6_kernel and the round-1 kernel (PIB_PAIR_V=5, the dz-form)."""
    from pyinterpolate_b200 import _lib
    if version == "5" and lags_per_group == "4":
        pytest.skip("the round-1 kernel has no four-lag scheme")
    pts = dem_like_field(n, seed=31)
    lags = np.arange(step, maxr, step)
    monkeypatch.setenv("PIB_FORCE_RING", "1")
    monkeypatch.setenv("PIB_RING", ring)                    # 32-slot ring x 256 threads or 16-slot ring x 512 threads
    monkeypatch.setenv("PIB_LAGS_PER_GROUP", lags_per_group)
    monkeypatch.setenv("PIB_PAIR_V", version)
    got = _lib.variogram_host(pts, lags)
    assert ("ring6" in _lib.last_kernel_name(_lib.KERNEL_PAIR)) == (version == "6")
    monkeypatch.delenv("PIB_FORCE_RING")
    monkeypatch.delenv("PIB_RING")
    monkeypatch.delenv("PIB_LAGS_PER_GROUP")
    monkeypatch.delenv("PIB_PAIR_V")
    want = oracle.variogram_sums(pts, lags)
    np.testing.assert_array_equal(got["count"], want["count"])
    np.testing.assert_allclose(got["sum_sq"], want["sum_sq"], rtol=RTOL)
    np.testing.assert_allclose(got["sum_prod"], want["sum_prod"], rtol=RTOL)
    np.testing.assert_allclose(got["sum_val"], want["sum_val"], rtol=RTOL)
    assert got["pairs_evaluated"] <= n * (n - 1) // 2


@pytest.mark.parametrize("lags_per_group", ["2", "4"])
@pytest.mark.parametrize("ring", ["32", "16"])
def test_ring_kernel_ellipse_vs_oracle(pib, oracle, monkeypatch, ring, lags_per_group):
    """Ellipse directions through the ring kernel (whitened bounding boxes for culling and lag windows), two-lag and
    four-lag groups."""
    from pyinterpolate_b200 import _lib, core
    pts = dem_like_field(12000, seed=35)
    lags = np.arange(3000.0, 40500.0, 3000.0) if lags_per_group == "2" else np.arange(700.0, 40500.0, 700.0)
    angles = [90, 0, 45, 135, 15]
    W = np.stack([core.define_whitening_matrix(a, 0.2) for a in angles]).reshape(-1, 4)
    monkeypatch.setenv("PIB_FORCE_RING", "1")
    monkeypatch.setenv("PIB_RING", ring)
    monkeypatch.setenv("PIB_LAGS_PER_GROUP", lags_per_group)
    got = _lib.variogram_host(pts, lags, lower=core.ellipse_lower_limits(lags), W=W)
    monkeypatch.delenv("PIB_FORCE_RING")
    monkeypatch.delenv("PIB_RING")
    monkeypatch.delenv("PIB_LAGS_PER_GROUP")
    want = oracle.variogram_sums(pts, lags, angles, 0.2)
    np.testing.assert_array_equal(got["count"], want["count"])
    np.testing.assert_allclose(got["sum_sq"], want["sum_sq"], rtol=RTOL)
    np.testing.assert_allclose(got["sum_prod"], want["sum_prod"], rtol=RTOL)
    full = _lib.variogram_host(pts, lags, lower=core.ellipse_lower_limits(lags), W=W)     # default kernel choice
    np.testing.assert_array_equal(full["count"], want["count"])
    assert got["pairs_evaluated"] < 5 * len(pts) * (len(pts) - 1) // 2


def test_ring_and_full_kernels_agree_300k(pib, monkeypatch):
    from pyinterpolate_b200 import _lib
    pts = dem_like_field(300_000, seed=33)
    lags = np.arange(500.0, 40500.0, 500.0)
    monkeypatch.setenv("PIB_FORCE_RING", "1")
    ring = _lib.variogram_host(pts, lags)
    monkeypatch.setenv("PIB_RING", "16")
    ring16 = _lib.variogram_host(pts, lags)
    monkeypatch.delenv("PIB_RING")
    monkeypatch.delenv("PIB_FORCE_RING")
    np.testing.assert_array_equal(ring16["count"], ring["count"])
    np.testing.assert_allclose(ring16["sum_sq"], ring["sum_sq"], rtol=1e-12)
    monkeypatch.setenv("PIB_FORCE_FULL", "1")
    full = _lib.variogram_host(pts, lags)
    monkeypatch.delenv("PIB_FORCE_FULL")
    np.testing.assert_array_equal(ring["count"], full["count"])
    for key in ("sum_sq", "sum_prod", "sum_val"):
        np.testing.assert_allclose(ring[key], full[key], rtol=1e-12)


def test_edge_cases(pib):
    from pyinterpolate_b200 import _lib
    lags = np.array([1.0, 2.0, 3.0])
    one = _lib.variogram_host(np.array([[0.0, 0.0, 1.0]]), lags)
    assert one["count"].sum() == 0
    # duplicates (distance 0) are in no lag; a distance exactly on an edge belongs to the lower lag
    pts = np.array([[0.0, 0.0, 1.0], [0.0, 0.0, 2.0], [1.0, 0.0, 4.0], [3.0, 0.0, 8.0]])
    res = _lib.variogram_host(pts, lags)
    np.testing.assert_array_equal(res["count"], [2, 1, 2])
    np.testing.assert_allclose(res["sum_sq"], [9 + 4, 16, 49 + 36])
    with pytest.raises(ValueError):
        pib.calculate_semivariance(pts)
    with pytest.raises(ValueError):
        pib.calculate_semivariance(pts, 1, 3, direction=10)
    with pytest.raises(ValueError):
        pib.calculate_semivariance(pts, 1, 3, direction=10, tolerance=0.5, dir_neighbors_selection_method="x")
    cloud = pib.ExperimentalVariogram(pts, 1, 3, as_cloud=True).point_cloud_semivariances
    assert [len(v) for v in cloud.values()] == [4, 2]          # np.arange(1, 3, 1) = two lags, ordered pairs
    with pytest.raises(_lib.PibError):
        _lib.variogram_host(pts, np.array([2.0, 1.0]))


def test_shards_add_up(pib):
    from pyinterpolate_b200 import _lib
    pts = dem_like_field(20000, seed=2)
    lags = np.arange(500.0, 40500.0, 500.0)
    full = _lib.variogram_host(pts, lags)
    parts = [_lib.variogram_host(pts, lags, shard=(r, 3)) for r in range(3)]
    np.testing.assert_array_equal(sum(p["count"] for p in parts), full["count"])
    np.testing.assert_allclose(sum(p["sum_sq"] for p in parts), full["sum_sq"], rtol=1e-12)
    assert sum(p["pairs_evaluated"] for p in parts) == full["pairs_evaluated"]
    assert full["pairs_evaluated"] <= 20000 * 19999 // 2
    assert full["count"].sum() <= full["pairs_evaluated"]


def test_100k_counts_vs_oracle(pib, oracle):
    """5e9 pairs: the largest size the CPU oracle finishes in a few seconds per core."""
    from pyinterpolate_b200 import _lib
    pts = dem_like_field(100_000, seed=14)
    lags = np.arange(500.0, 40500.0, 500.0)
    got = _lib.variogram_host(pts, lags)
    want = oracle.variogram_sums(pts, lags)
    np.testing.assert_array_equal(got["count"], want["count"])
    np.testing.assert_allclose(got["sum_sq"], want["sum_sq"], rtol=RTOL)
    np.testing.assert_allclose(got["sum_prod"], want["sum_prod"], rtol=RTOL)
    np.testing.assert_allclose(got["sum_val"], want["sum_val"], rtol=RTOL)


def test_1m_points_properties(pib):
    """BASELINE config 2 size (1M points, 80 lags): properties that need no oracle —
    permutation invariance (counts exact), z -> a z + b scaling, device/host entry points agree."""
    import torch
    from pyinterpolate_b200 import _lib
    pts = dem_like_field(1_000_000, seed=2)
    lags = np.arange(500.0, 40500.0, 500.0)
    assert len(lags) == 80
    base = _lib.variogram_host(pts, lags)
    n = len(pts)
    assert 0 < base["count"].sum() < n * (n - 1) // 2
    assert base["count"].sum() <= base["pairs_evaluated"] < n * (n - 1) // 2      # tile culling skipped work
    perm = np.random.default_rng(0).permutation(n)
    shuffled = _lib.variogram_host(pts[perm], lags)
    np.testing.assert_array_equal(shuffled["count"], base["count"])
    np.testing.assert_allclose(shuffled["sum_sq"], base["sum_sq"], rtol=1e-11)
    scaled = pts.copy(); scaled[:, 2] = 3.0 * scaled[:, 2] - 100.0
    sc = _lib.variogram_host(scaled, lags)
    np.testing.assert_array_equal(sc["count"], base["count"])
    np.testing.assert_allclose(sc["sum_sq"], 9.0 * base["sum_sq"], rtol=1e-11)
    dev = _lib.variogram_device(torch.from_numpy(pts).cuda(), lags)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(dev["count"].cpu().numpy(), base["count"])
    np.testing.assert_array_equal(dev["sum_sq"].cpu().numpy(), base["sum_sq"])   # deterministic reduction order


def test_baseline_config_directional_full_size(pib):
    """BASELINE directional config at full size (200k points, 4 ellipse directions, tolerance 0.2, 80 lags):
    permutation invariance with exact counts, shard additivity, and the fused 4-direction call equals four
    single-direction calls bit for bit."""
    from pyinterpolate_b200 import _lib, core
    pts = dem_like_field(200_000, seed=6)
    lags = np.arange(500.0, 40500.0, 500.0)
    angles = [90, 0, 45, 135]
    W = np.stack([core.define_whitening_matrix(a, 0.2) for a in angles]).reshape(-1, 4)
    lower = core.ellipse_lower_limits(lags)
    base = _lib.variogram_host(pts, lags, lower=lower, W=W)
    assert base["count"].shape == (4, 80) and (base["count"] > 0).all()
    n = len(pts)
    assert base["count"].sum(axis=1).max() < n * (n - 1) // 2
    perm = np.random.default_rng(1).permutation(n)
    shuffled = _lib.variogram_host(pts[perm], lags, lower=lower, W=W)
    np.testing.assert_array_equal(shuffled["count"], base["count"])
    np.testing.assert_allclose(shuffled["sum_sq"], base["sum_sq"], rtol=1e-11)
    for t in (0, 3):
        one = _lib.variogram_host(pts, lags, lower=lower, W=W[t:t + 1])
        np.testing.assert_array_equal(one["count"][0], base["count"][t])
        np.testing.assert_array_equal(one["sum_sq"][0], base["sum_sq"][t])
    parts = [_lib.variogram_host(pts, lags, lower=lower, W=W, shard=(r, 4)) for r in range(4)]
    np.testing.assert_array_equal(sum(p["count"] for p in parts), base["count"])
    np.testing.assert_allclose(sum(p["sum_sq"] for p in parts), base["sum_sq"], rtol=1e-12)
