"""GPU parity for the rank-1 callers of the kriging path: interpolate_raster, UniversalKriging,
interpolate_points_dask — against fixtures from the unmodified reference (tests/golden/make_golden_pipelines.py)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = 1e-9


@pytest.fixture(scope="module")
def pib():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pyinterpolate_b200 as pib
    return pib


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_golden_pipelines.npz"))


def test_interpolate_raster(pib, gold):
    known = gold["known"]
    var = float(np.var(known[:, 2]))
    mdl = {"variogram_model_type": "spherical", "nugget": 0.3, "sill": var, "rang": 6000.0}
    r = pib.interpolate_raster(known, dim=37, number_of_neighbors=8, semivariogram_model=mdl)
    assert r["result"].shape == gold["raster_result"].shape == (37, 38)
    np.testing.assert_allclose(r["result"], gold["raster_result"], rtol=RTOL)
    np.testing.assert_allclose(r["error"], gold["raster_error"], rtol=RTOL, atol=1e-9 * var)
    got = np.array([r["params"][k] for k in ("pixel size", "min x", "max x", "min y", "max y")])
    np.testing.assert_array_equal(got, gold["raster_params"])
    with pytest.raises(NotImplementedError):
        pib.interpolate_raster(known, dim=10)


def test_interpolate_points_dask(pib, gold):
    known, unk = gold["known"], gold["unknown"]
    mdl = pib.TheoreticalVariogram({"variogram_model_type": "spherical", "nugget": 0.3, "sill": float(np.var(known[:, 2])),
                                    "rang": 6000.0})
    got = pib.interpolate_points_dask(mdl, known, unk, no_neighbors=12, number_of_workers=4)
    np.testing.assert_allclose(got[:, :2], gold["dask1"][:, :2], rtol=RTOL)
    np.testing.assert_array_equal(got[:, 2:], gold["dask1"][:, 2:])


def test_universal_kriging(pib, gold):
    known, unk = gold["known"], gold["unknown"]
    uk = pib.UniversalKriging(known)
    uk.fit_trend()
    np.testing.assert_allclose(np.append(uk.trend_model.coefficients, uk.trend_model.intercept), gold["uk_coefficients"],
                               rtol=1e-10)
    uk.detrend()
    sill = float(np.var(uk.bias_values))
    np.testing.assert_allclose(sill, float(gold["uk_bias_sill"]), rtol=1e-10)
    bias_model = {"variogram_model_type": "exponential", "nugget": 0.5, "sill": float(gold["uk_bias_sill"]), "rang": 5000.0}
    with pytest.raises(NotImplementedError):
        uk.fit_bias(step_size=500, max_range=8000)                       # no model fitting in this package
    assert uk.bias_experimental_model is not None and uk.bias_experimental_model.covariances is None
    uk.fit_bias(step_size=500, max_range=8000, bias_model=bias_model)
    pred = uk.predict(unk, no_neighbors=10)
    assert pred.shape == gold["uk_pred"].shape
    np.testing.assert_allclose(pred[:, 0], gold["uk_pred"][:, 0], rtol=RTOL)
    np.testing.assert_array_equal(pred[:, 1:], gold["uk_pred"][:, 1:])


def test_validate_kriging_leave_one_out(pib, gold):
    """evaluate/cross_validation.py:11-138 as one batched launch with the point's own row excluded."""
    pts = gold["cv_points"]
    mdl = {"variogram_model_type": "spherical", "nugget": 0.4, "sill": float(np.var(pts[:, 2])), "rang": 2500.0}
    for how, kw in (("ok", {}), ("sk", {"sk_mean": float(pts[:, 2].mean())})):
        mpe, mve, arr = pib.validate_kriging(pts, mdl, how=how, no_neighbors=8, **kw)
        want = gold[f"cv_{how}_arr"]
        np.testing.assert_array_equal(arr[:, :2], want[:, :2])
        np.testing.assert_allclose(arr[:, 2], want[:, 2], rtol=1e-8, atol=1e-9 * np.abs(pts[:, 2]).max())
        np.testing.assert_allclose(arr[:, 3], want[:, 3], rtol=RTOL)
        np.testing.assert_allclose([mpe, mve], gold[f"cv_{how}_stats"], rtol=1e-7, atol=1e-12)
    _, _, arr = pib.validate_kriging(pts[:40], mdl, how="ok", no_neighbors=64)          # k > n - 1: all other rows
    np.testing.assert_allclose(arr[:, 2:], gold["cv_small_arr"][:, 2:], rtol=1e-7, atol=1e-9 * np.abs(pts[:, 2]).max())
    with pytest.raises(KeyError):
        pib.validate_kriging(pts, mdl, how="uk")


def test_leave_one_out_equals_explicit_deletion(pib):
    """The exclusion in the neighbour search is the same as kriging each point from the array without it —
    checked on clustered data (incremental search path) and with a directional model."""
    rng = np.random.default_rng(95)
    centres = rng.uniform(0, 30_000, size=(6, 2))
    xy = np.vstack([c + rng.normal(0, 150.0, size=(60, 2)) for c in centres])
    pts = np.column_stack([xy, 20 + np.sin(xy[:, 0] / 3000.0) + rng.normal(0, 0.2, len(xy))])
    for mdl in ({"variogram_model_type": "exponential", "nugget": 0.01, "sill": float(np.var(pts[:, 2])), "rang": 4000.0},
                {"variogram_model_type": "exponential", "nugget": 0.01, "sill": float(np.var(pts[:, 2])), "rang": 4000.0,
                 "direction": 40.0}):
        _, _, arr = pib.validate_kriging(pts, mdl, how="ok", no_neighbors=6)
        for i in (0, 59, 60, 200, 359):
            one = pib.ordinary_kriging(mdl, np.delete(pts, i, 0), pts[i:i + 1, :2], no_neighbors=6)[0]
            np.testing.assert_allclose(pts[i, 2] - arr[i, 2], one[0], rtol=1e-12, equal_nan=True)
            np.testing.assert_allclose(arr[i, 3], one[1], rtol=1e-12, equal_nan=True)
