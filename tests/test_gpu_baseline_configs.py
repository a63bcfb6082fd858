"""BASELINE.json configs 2-5 at their STATED sizes, with SURVEY.md §8d's exact definitions (seeds, extents, lags,
models, neighbour counts), pinned to the CPU oracle — not only to size-independent properties.

Variogram configs: the oracle evaluates rows [0, R) of the pair triangle (`oracle.variogram_sums(rows=)`, ~3 s of CPU
at 1M points); on the GPU  full(all N points) - full(points R..N)  is exactly that pair set (every pair whose smaller
index is below R), so the pair COUNTS must match bit for bit and the sums to 1e-9.
Kriging configs: the whole batch runs on the GPU, a random sample of the unknown locations is checked against
`oracle.krige` (neighbour indices bit-exact, zhat and sigma^2 to 1e-9).

Reference paths: semivariogram/experimental/functions/general.py:235-303, functions/directional.py:395-465,
kriging/point/ordinary.py:134-221, simple.py:129-223, kriging/point/indicator.py:200-325.
"""
import os

import numpy as np
import pytest

from pyinterpolate_b200.synthetic import dem_like_field, uniform_locations

pytestmark = pytest.mark.gpu

RTOL = 1e-9                      # BASELINE.json: semivariances / predictions / variances within 1e-9 relative
LAGS_80 = np.arange(500.0, 40500.0, 500.0)      # step 500, max_range 40500 -> 80 lags (§8d C2)


@pytest.fixture(scope="module")
def pib():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import pyinterpolate_b200 as p
    return p


def _rows_by_subtraction(_lib, pts, lags, rows, **kw):
    """Sums over the pairs (i, j), i < j, i < rows — as the difference of two full GPU passes."""
    full = _lib.variogram_host(pts, lags, **kw)
    tail = _lib.variogram_host(pts[rows:], lags, **kw)
    return {k: full[k] - tail[k] for k in ("sum_sq", "sum_prod", "sum_val", "count")}, full


def _check_sums(got, want, z, count_floor=0):
    np.testing.assert_array_equal(got["count"], want["count"])
    big = want["count"] > count_floor
    # the GPU side is a difference of two sums ~N/R times larger: its own rounding is amplified by that factor,
    # still far inside the 1e-9 bar
    np.testing.assert_allclose(got["sum_sq"][big], want["sum_sq"][big], rtol=RTOL)
    np.testing.assert_allclose(got["sum_val"][big], want["sum_val"][big], rtol=RTOL)
    scale = want["count"] * (np.mean(z) ** 2 + np.var(z))     # a covariance is a difference of two terms of this size
    np.testing.assert_allclose(got["sum_prod"][big], want["sum_prod"][big], rtol=0, atol=RTOL * scale[big].max())


@pytest.mark.parametrize("extent", [100_000.0, 1_000_000.0], ids=["dense_100km", "sparse_1000km"])
def test_config2_omni_1m_vs_oracle(pib, oracle, extent):
    """C2: N = 1e6, seed 2, E = 100 km (dense, ~40 % of the pairs in range) and the sparse E = 1000 km variant."""
    from pyinterpolate_b200 import _lib
    pts = dem_like_field(1_000_000, extent=extent, seed=2)
    rows = 2500
    got, full = _rows_by_subtraction(_lib, pts, LAGS_80, rows)
    want = oracle.variogram_sums(pts, LAGS_80, rows=(0, rows))
    assert want["count"].sum() > 0
    _check_sums(got, want, pts[:, 2])
    # the public class on the same field: semivariance / covariance / counts of the full pass are self-consistent
    n = len(pts)
    assert full["count"].sum() <= full["pairs_evaluated"] <= n * (n - 1) // 2


def test_config3_directional_200k_vs_oracle(pib, oracle):
    """C3: N = 200 000, seed 3, directions (90, 0, 45, 135) = DirectionalVariogram.directions, tolerance 0.2, method 'e'."""
    from pyinterpolate_b200 import _lib, core
    pts = dem_like_field(200_000, extent=100_000.0, seed=3)
    angles = [90, 0, 45, 135]
    W = np.stack([core.define_whitening_matrix(a, 0.2) for a in angles]).reshape(-1, 4)
    lower = core.ellipse_lower_limits(LAGS_80)
    rows = 2500
    got, full = _rows_by_subtraction(_lib, pts, LAGS_80, rows, lower=lower, W=W)
    want = oracle.variogram_sums(pts, LAGS_80, directions=angles, tolerance=0.2, rows=(0, rows))
    assert got["count"].shape == (4, 80) and (want["count"] > 0).all()
    _check_sums(got, want, pts[:, 2])
    # and through the public class (one fused call for the four directions + ISO)
    dv = pib.DirectionalVariogram(pts, step_size=500, max_range=40500, tolerance=0.2, dir_neighbors_selection_method="e")
    for t, name in enumerate(["NS", "WE", "NE-SW", "NW-SE"]):
        np.testing.assert_array_equal(dv.directional_variograms[name].points_per_lag, 2.0 * full["count"][t])


def _check_kriging_sample(pib, oracle, known, unk, out, idx, sample, model, nugget, sill, rang, k, mode, mean=0.0):
    ref, ref_idx, ref_st = oracle.krige(known, unk[sample], model, nugget, sill, rang, k, mode, process_mean=mean)
    assert (ref_st == 0).all()
    if idx is not None:
        np.testing.assert_array_equal(idx[sample], ref_idx)
    np.testing.assert_allclose(out[sample, 0], ref[:, 0], rtol=RTOL)
    np.testing.assert_allclose(out[sample, 1], ref[:, 1], rtol=RTOL, atol=1e-12)
    np.testing.assert_array_equal(out[sample, 2:], unk[sample])


def test_config4_ok_sk_10m_vs_oracle(pib, oracle):
    """C4: known N = 1e6 (seed 4), unknown M = 1e7 (seed 5), no_neighbors = 64, spherical {nugget 1, sill var(z),
    rang 20 000}; ordinary AND simple kriging (process_mean = mean(z)) from ONE call that shares the neighbour search."""
    known = dem_like_field(1_000_000, extent=100_000.0, seed=4)
    unk = uniform_locations(10_000_000, extent=100_000.0, seed=5)
    var, mean = float(np.var(known[:, 2])), float(np.mean(known[:, 2]))
    mdl = {"variogram_model_type": "spherical", "nugget": 1.0, "sill": var, "rang": 20000.0}
    ok, sk = pib.ordinary_and_simple_kriging(mdl, known, unk, process_mean=mean, no_neighbors=64)
    assert ok.shape == sk.shape == (10_000_000, 4)
    assert np.isfinite(ok[:, :2]).all() and np.isfinite(sk[:, :2]).all()
    sample = np.random.default_rng(45).choice(len(unk), 256, replace=False)
    _check_kriging_sample(pib, oracle, known, unk, ok, None, sample, "spherical", 1.0, var, 20000.0, 64, "ok")
    _check_kriging_sample(pib, oracle, known, unk, sk, None, sample, "spherical", 1.0, var, 20000.0, 64, "sk", mean)
    # the two single-mode public calls give the same rows as the shared call, bit for bit
    sub = np.sort(sample)
    ok1, idx1 = pib.ordinary_kriging(mdl, known, unk[sub], no_neighbors=64, return_neighbors=True)
    sk1 = pib.simple_kriging(mdl, known, unk[sub], process_mean=mean, no_neighbors=64)
    np.testing.assert_array_equal(ok1, ok[sub])
    np.testing.assert_array_equal(sk1, sk[sub])
    _, ref_idx, _ = oracle.krige(known, unk[sub], "spherical", 1.0, var, 20000.0, 64, "ok")
    np.testing.assert_array_equal(idx1, ref_idx)


def test_config5_indicator_500k_vs_oracle(pib, oracle):
    """C5: N = 500 000 (seed 6), thresholds = 8 quantiles (semivariogram/indicator/indicator.py:47-71), 8 spherical
    models {nugget 0.01, sill p (1 - p), rang 20 000}, M = 1e6 unknown (seed 7), no_neighbors = 32, kriging_type 'ok'."""
    from pyinterpolate_b200 import indicator
    known = dem_like_field(500_000, extent=100_000.0, seed=6)
    unk = uniform_locations(1_000_000, extent=100_000.0, seed=7)
    thresholds = indicator.select_variogram_thresholds(known[:, 2], 8)
    probs = np.linspace(0, 1, 9)[1:]
    models = [{"variogram_model_type": "spherical", "nugget": 0.01, "sill": max(float(p * (1 - p)), 1e-3), "rang": 20000.0}
              for p in probs]
    zhat = pib.kriging.indicator_kriging(models, thresholds, known, unk, no_neighbors=32, reference_variance_column=False)
    sig = pib.kriging.indicator_kriging(models, thresholds, known, unk, no_neighbors=32, reference_variance_column=True)
    assert zhat.shape == sig.shape == (1_000_000, 8)
    sample = np.random.default_rng(67).choice(len(unk), 200, replace=False)
    for t, (thr, m) in enumerate(zip(thresholds, models)):
        coded = known.copy()
        coded[:, 2] = (known[:, 2] <= thr).astype(np.float64)       # kriging/point/indicator.py:272-275
        ref, _, st = oracle.krige(coded, unk[sample], "spherical", m["nugget"], m["sill"], m["rang"], 32, "ok")
        assert (st == 0).all()
        np.testing.assert_allclose(zhat[sample, t], np.clip(ref[:, 0], 0, 1), rtol=RTOL, atol=1e-12)
        want_sig = np.where(np.isnan(ref[:, 1]), np.nan, np.clip(ref[:, 1], 0, 1))     # the column the reference stores (:312-313)
        np.testing.assert_allclose(sig[sample, t], want_sig, rtol=RTOL, atol=1e-12)


def test_smooth_field_fine_lags_vs_oracle(pib, oracle, monkeypatch):
    """The z_j-form / group-sum closed forms cancel when gamma(h) << var(z).  A noise-free smooth surface (the
    reference's own DEM sample, tests/test_kriging/point_kriging_ds/dem2180.csv, shipped as a fixture; plus a synthetic
    noise-free trend) with fine lags is where it bites: the result must still meet 1e-9, through every ring variant."""
    from pyinterpolate_b200 import _lib
    path = os.path.join(os.path.dirname(__file__), "golden", "dem2180.csv")
    dem = np.loadtxt(path, delimiter=",", skiprows=1)
    assert dem.shape == (6895, 3)
    rng = np.random.default_rng(8)
    xy = rng.uniform(0.0, 20_000.0, size=(60_000, 2))
    smooth = np.column_stack([xy, 200.0 + 50.0 * np.sin(xy[:, 0] / 2.0e4) * np.cos(xy[:, 1] / 1.5e4)])      # no noise term
    cases = [("dem2180", dem, np.arange(50.0, 2000.0, 50.0)),
             ("smooth_60k", smooth, np.arange(20.0, 1600.0, 20.0)),
             ("smooth_60k_offset", smooth + np.array([0.0, 0.0, 1.0e6]), np.arange(20.0, 1600.0, 20.0))]
    for name, pts, lags in cases:
        want = oracle.variogram_sums(pts, lags)
        for force in (None, "ring16", "ring32", "v5"):
            for k in ("PIB_FORCE_RING", "PIB_RING", "PIB_PAIR_V"):
                monkeypatch.delenv(k, raising=False)
            if force in ("ring16", "ring32"):
                monkeypatch.setenv("PIB_FORCE_RING", "1")
                monkeypatch.setenv("PIB_RING", force[4:])
            if force == "v5":
                monkeypatch.setenv("PIB_FORCE_RING", "1")
                monkeypatch.setenv("PIB_PAIR_V", "5")
            got = _lib.variogram_host(pts, lags)
            np.testing.assert_array_equal(got["count"], want["count"], err_msg=f"{name} {force}")
            ok = want["count"] > 0
            np.testing.assert_allclose(got["sum_sq"][ok], want["sum_sq"][ok], rtol=RTOL, err_msg=f"{name} {force}")
            np.testing.assert_allclose(got["sum_val"][ok], want["sum_val"][ok], rtol=RTOL, err_msg=f"{name} {force}")
