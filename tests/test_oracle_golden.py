"""Pins the CPU oracle (oracle/pyinterp_oracle.c) to the unmodified reference's outputs.

Tolerances: pair counts and neighbour indices bit-exact; floating-point values 1e-12 relative
(the oracle sums in a different order than numpy's pairwise mean, nothing else differs).
Covariances subtract two ~4e4-sized terms, so their tolerance is relative to mean(z)^2.
"""
import numpy as np
import pytest

from pyinterpolate_b200.synthetic import dem_like_field, uniform_locations

RTOL = 1e-12


def _check_table(got, want, scale=None):
    assert got.shape == want.shape
    np.testing.assert_array_equal(got[:, 0], want[:, 0])                    # lags
    np.testing.assert_array_equal(got[:, 2], want[:, 2])                    # pair counts (nan == nan)
    if scale is None:
        np.testing.assert_allclose(got[:, 1], want[:, 1], rtol=RTOL, atol=0, equal_nan=True)
    else:
        np.testing.assert_allclose(got[:, 1], want[:, 1], rtol=0, atol=RTOL * scale, equal_nan=True)


def _cov_scale(points):
    z = np.asarray(points, dtype=float)[:, -1]
    return float(np.mean(z) ** 2 + np.var(z))


def test_line13_known_answer(oracle, golden):
    pts = golden["line13_points"]
    sem = oracle.calculate_semivariance(pts, 1, 4)
    # the reference's own golden values (tests/test_semivariogram/test_experimental_semivariance.py:32-42)
    np.testing.assert_allclose(sem, [[1, 4.625, 24], [2, 5.2272727272727275, 22], [3, 6.0, 20]], rtol=1e-15)
    _check_table(sem, golden["line13_sem"])
    _check_table(oracle.calculate_covariance(pts, 1, 4), golden["line13_cov"], _cov_scale(pts))


def test_armstrong_omni_and_ellipse(oracle, golden):
    pts = golden["armstrong_points"]
    _check_table(oracle.calculate_semivariance(pts, 1, 6), golden["armstrong_omni_sem"])
    _check_table(oracle.calculate_covariance(pts, 1, 6), golden["armstrong_omni_cov"], _cov_scale(pts))
    _check_table(oracle.calculate_semivariance(pts, 1.5, 6, direction=15, tolerance=0.25),
                 golden["armstrong_ell_sem"])
    _check_table(oracle.calculate_covariance(pts, 1.5, 6, direction=15, tolerance=0.25),
                 golden["armstrong_ell_cov"], _cov_scale(pts))
    # SURVEY §8c known answers
    np.testing.assert_array_equal(golden["armstrong_omni_sem"][:, 2], [224, 388, 640, 648, 772])
    np.testing.assert_array_equal(golden["armstrong_ell_sem"][:, 2], [112, 488, 820])


def test_armstrong_custom_bins(oracle, golden):
    pts = golden["armstrong_points"]
    bins = golden["armstrong_custom_bins"]
    _check_table(oracle.calculate_semivariance(pts, custom_bins=bins), golden["armstrong_custom_sem"])
    _check_table(oracle.calculate_covariance(pts, custom_bins=bins), golden["armstrong_custom_cov"], _cov_scale(pts))
    _check_table(oracle.calculate_semivariance(pts, custom_bins=golden["armstrong_custom_ell_bins"],
                                               direction=45, tolerance=0.5),
                 golden["armstrong_custom_ell_sem"])
    # the lag the reference raised on is empty for the oracle too
    assert bool(golden["armstrong_empty_dir_lag_raises"])
    t = oracle.calculate_semivariance(pts, custom_bins=bins, direction=45, tolerance=0.5)
    assert np.isnan(t[1:, 2]).any() or (t[:, 2] == 0).any()


def test_sparse_empty_lags(oracle, golden):
    pts = dem_like_field(40, extent=1000.0, seed=11)
    want = golden["sparse_sem"]
    assert np.isnan(want[:, 1]).any()                       # the case really has empty lags
    _check_table(oracle.calculate_semivariance(pts, 2.0, 60.0), want)
    _check_table(oracle.calculate_covariance(pts, 2.0, 60.0), golden["sparse_cov"], _cov_scale(pts))


@pytest.mark.parametrize("name,n,seed,extent,step,maxr", [
    ("omni2k", 2000, 21, 100_000.0, 500, 40000),
    ("omni10k", 10000, 0, 100_000.0, 500, 40000),
    ("omni600", 600, 23, 10.0, 0.1, 3.05),
])
def test_synthetic_omni(oracle, golden, name, n, seed, extent, step, maxr):
    pts = dem_like_field(n, extent=extent, seed=seed)
    _check_table(oracle.calculate_semivariance(pts, step, maxr), golden[f"{name}_sem"])
    _check_table(oracle.calculate_covariance(pts, step, maxr), golden[f"{name}_cov"], _cov_scale(pts))


@pytest.mark.parametrize("ang", [90, 0, 45, 135, 15])
def test_synthetic_ellipse(oracle, golden, ang):
    pts = dem_like_field(1500, seed=22)
    _check_table(oracle.calculate_semivariance(pts, 2500, 40000, direction=ang, tolerance=0.2),
                 golden[f"ell1500_dir{ang}_sem"])
    _check_table(oracle.calculate_covariance(pts, 2500, 40000, direction=ang, tolerance=0.2),
                 golden[f"ell1500_dir{ang}_cov"], _cov_scale(pts))


def test_ellipse_float_step(oracle, golden):
    pts = dem_like_field(600, extent=10.0, seed=23)
    _check_table(oracle.calculate_semivariance(pts, 0.1, 3.05, direction=30, tolerance=0.35), golden["ell600_sem"])
    _check_table(oracle.calculate_covariance(pts, 0.1, 3.05, direction=30, tolerance=0.35),
                 golden["ell600_cov"], _cov_scale(pts))


def test_fused_directions_equal_single(oracle):
    pts = dem_like_field(700, seed=5)
    lags = oracle.get_lags(2500, 40000)
    fused = oracle.variogram_sums(pts, lags, [90, 0, 45, 135], 0.2)
    for t, ang in enumerate([90, 0, 45, 135]):
        single = oracle.variogram_sums(pts, lags, [ang], 0.2)
        np.testing.assert_array_equal(fused["count"][t], single["count"][0])
        np.testing.assert_allclose(fused["sum_sq"][t], single["sum_sq"][0], rtol=1e-13)


def test_whitening_matrix(oracle, golden):
    for (ang, tol), W in zip(golden["whiten_cfg"], golden["whiten_W"]):
        np.testing.assert_array_equal(oracle.whitening_matrix(ang, tol), W)


def test_grid5_ellipse_selection(oracle, golden):
    """Golden selections of the reference's test_select_points_within_ellipse.py, through the
    oracle's pair pass: centre point (0,0) vs the 5x5 grid, one lag (0, 2]."""
    grid = golden["grid5_points"]
    centre = int(np.where((grid == 0).all(axis=1))[0][0])
    for (ang, tol), mask in zip(golden["grid5_cfg"], golden["grid5_masks"]):
        # values: indicator of "is the centre" so that sum_val counts pairs touching the centre
        z = np.zeros(len(grid)); z[centre] = 1.0
        pts = np.column_stack([grid, z])
        sums = oracle.variogram_sums(pts, np.array([2.0]), [ang], tol)
        assert int(sums["sum_val"][0, 0]) == int(mask.sum()), (ang, tol)


def test_gamma_models(oracle, golden):
    h = golden["gamma_h"]
    for mi, name in enumerate(golden["gamma_models"]):
        for ti, (nug, sill, rang) in enumerate(golden["gamma_params"]):
            got = oracle.gamma(str(name), h, nug, sill, rang)
            np.testing.assert_allclose(got, golden["gamma_values"][mi, ti], rtol=1e-13, atol=1e-15)


def test_armstrong_kriging_known_answer(oracle, golden):
    arm = golden["armstrong_points"].astype(float)
    unk = golden["armstrong_krige_unknown"]
    sill = float(golden["armstrong_krige_sill"])
    # SURVEY §8c known answers for u = (3.3, 4.7)
    np.testing.assert_allclose(golden["armstrong_ok"][0, :2], [20.429753072465108, 2.410750487874925], rtol=1e-13)
    np.testing.assert_allclose(golden["armstrong_sk"][0, :2], [20.328741461044903, 2.4200627412499633], rtol=1e-13)
    # rows 0..3 are tie-free enough?  Armstrong is a regular grid: distance ties are everywhere, so only
    # compare locations whose k-th and (k+1)-th neighbour distances differ (the set is then well defined
    # and the solution does not depend on the order within the set beyond rounding).
    ok, _, st = oracle.krige(arm, unk, "spherical", 0.0, sill, 4.0, 8, "ok")
    sk, _, _ = oracle.krige(arm, unk, "spherical", 0.0, sill, 4.0, 8, "sk", process_mean=float(arm[:, -1].mean()))
    for u in range(len(unk)):
        d = np.sort(np.hypot(arm[:, 0] - unk[u, 0], arm[:, 1] - unk[u, 1]))
        if d[7] == d[8]:
            continue
        np.testing.assert_allclose(ok[u, :2], golden["armstrong_ok"][u, :2], rtol=1e-10)
        np.testing.assert_allclose(sk[u, :2], golden["armstrong_sk"][u, :2], rtol=1e-10)
    np.testing.assert_array_equal(ok[:, 2:], unk)


@pytest.mark.parametrize("name", ["linear", "spherical", "exponential", "circular", "cubic"])
def test_kriging_10k(oracle, golden, name):
    known = dem_like_field(10000, seed=0)
    unk = uniform_locations(200, seed=1)
    var = float(golden["krige10k_var"])
    nug = 0.0 if name == "linear" else 1.0
    ok, idx, st = oracle.krige(known, unk, name, nug, var, 8000.0, 32, "ok")
    sk, idx2, _ = oracle.krige(known, unk, name, nug, var, 8000.0, 32, "sk", process_mean=float(known[:, -1].mean()))
    np.testing.assert_array_equal(idx, golden["krige10k_nbr_idx"])          # neighbour indices bit-exact
    np.testing.assert_array_equal(idx2, idx)
    assert (st == 0).all()
    np.testing.assert_allclose(ok[:, :2], golden[f"krige10k_{name}_ok"][:, :2], rtol=1e-9)
    np.testing.assert_allclose(sk[:, :2], golden[f"krige10k_{name}_sk"][:, :2], rtol=1e-9)
    np.testing.assert_array_equal(ok[:, 2:], golden[f"krige10k_{name}_ok"][:, 2:])


def test_kriging_200k_k64(oracle, golden):
    known = dem_like_field(200_000, extent=44_721.36, seed=4)
    unk = uniform_locations(40, extent=44_721.36, seed=5)
    var = float(golden["krige200k_var"])
    ok, idx, _ = oracle.krige(known, unk, "spherical", 1.0, var, 20000.0, 64, "ok")
    sk, _, _ = oracle.krige(known, unk, "spherical", 1.0, var, 20000.0, 64, "sk", process_mean=float(known[:, -1].mean()))
    np.testing.assert_array_equal(idx, golden["krige200k_nbr_idx"])
    np.testing.assert_allclose(ok[:, :2], golden["krige200k_ok"][:, :2], rtol=1e-9)
    np.testing.assert_allclose(sk[:, :2], golden["krige200k_sk"][:, :2], rtol=1e-9)


def test_kriging_edge_cases(oracle, golden):
    few = dem_like_field(5, extent=100.0, seed=31)
    unk = np.array([[50.0, 50.0], [10.0, 80.0]])
    ok, idx, _ = oracle.krige(few, unk, "spherical", 0.5, 4.0, 60.0, 8, "ok")
    sk, _, _ = oracle.krige(few, unk, "spherical", 0.5, 4.0, 60.0, 8, "sk", process_mean=200.0)
    np.testing.assert_allclose(ok[:, :2], golden["few_ok"][:, :2], rtol=1e-10)
    np.testing.assert_allclose(sk[:, :2], golden["few_sk"][:, :2], rtol=1e-10)
    assert (idx[:, 5:] == -1).all()
    known = dem_like_field(10000, seed=0)
    unk = uniform_locations(200, seed=1)[:50]
    ok, _, _ = oracle.krige(known, unk, "spherical", 1.0, float(golden["krige10k_var"]), 500.0, 16, "ok")
    np.testing.assert_allclose(ok[:, :2], golden["krige10k_shortrange_ok"][:, :2], rtol=1e-9)
    # duplicated coordinates: the reference raises RuntimeError (singular); the oracle flags status 1
    arm = golden["armstrong_points"].astype(float)
    dup = np.vstack([arm, arm[:3]])
    assert bool(golden["dup_raises"])
    _, _, st = oracle.krige(dup, np.array([[0.4, 0.3]]), "spherical", 0.0, float(golden["armstrong_krige_sill"]), 4.0, 8, "ok")
    assert st[0] == 1


def test_lstsq_fallback_golden(oracle):
    """allow_approximate_solutions (point_kriging_solve.py:134-147): the oracle reproduces the reference on every
    row where the reference itself went through np.linalg.lstsq and on every non-singular row.  (On the other
    duplicate-coordinate rows the reference's LAPACK misses the singularity and returns garbage; the oracle
    treats them as singular — tests/golden/make_golden_lsa.py records which is which.)"""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden_lsa.npz"))
    known, unk, var = g["known"], g["unknown"], float(g["var"])
    for name, nug in (("spherical", 0.0), ("exponential", 0.7), ("linear", 0.0)):
        for mode, k, pm in (("ok", 16, 0.0), ("sk", 12, float(known[:, 2].mean()))):
            out, _, st = oracle.krige_with_fallback(known, unk, name, nug, var, 9000.0, k, mode, pm)
            ref, ls, sing = g[f"{mode}_{name}"], g[f"{mode}_{name}_lstsq"], g[f"singular{k}"]
            np.testing.assert_array_equal(st == 4, sing)
            keep = ls | ~sing
            assert ls.sum() >= 10
            np.testing.assert_array_equal(np.isnan(out[keep, 1]), np.isnan(ref[keep, 1]))
            np.testing.assert_allclose(np.nan_to_num(out[keep, :2]), np.nan_to_num(ref[keep, :2]), rtol=1e-10)
    out, _, st = oracle.krige_with_fallback(known, unk[:10], "spherical", 0.0, 0.0, 9000.0, 6, "sk", 3.5)
    assert (st == 5).all()
    np.testing.assert_array_equal(out, g["sk_zero"])
