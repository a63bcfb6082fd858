"""GPU parity for the indicator path against the unmodified reference's outputs
(tests/golden/make_golden_indicator.py)."""
import os
import types

import numpy as np
import pytest

from pyinterpolate_b200.synthetic import dem_like_field

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pib():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pyinterpolate_b200 as pib
    return pib


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_golden_indicator.npz"))


def test_experimental_indicator_variogram_golden(pib, gold):
    pts = dem_like_field(1200, seed=41)
    eiv = pib.ExperimentalIndicatorVariogram(pts, number_of_thresholds=4, step_size=2500, max_range=40000)
    np.testing.assert_array_equal(np.array(eiv.ds.thresholds), gold["thresholds"])
    assert list(eiv.experimental_models.keys()) == [str(k) for k in gold["keys"]]
    for i, m in enumerate(eiv.experimental_models.values()):
        np.testing.assert_array_equal(m.lags, gold["lags"])
        np.testing.assert_array_equal(m.points_per_lag, gold["ppl"][i])          # pair counts bit-exact
        # indicator values are 0/1: the sums are exact integers, so these agree to rounding of the final division
        np.testing.assert_allclose(m.semivariances, gold["sem"][i], rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(m.covariances, gold["cov"][i], rtol=0, atol=1e-12)
        assert m.variance == gold["var"][i]


def test_indicator_kriging_golden(pib, gold):
    pts = dem_like_field(1200, seed=41)
    models = {}
    for k, p in zip(gold["keys"], (0.25, 0.5, 0.75, 1.0)):
        models[str(k)] = pib.TheoreticalVariogram({"variogram_model_type": "spherical", "nugget": 0.01,
                                                   "sill": max(p * (1 - p), 0.05), "rang": 20000.0})
    holder = types.SimpleNamespace(theoretical_indicator_variograms=models)
    ik = pib.IndicatorKriging(holder, pts, gold["ik_unknown"], kriging_type="ok", no_neighbors=16)
    np.testing.assert_allclose(ik.indicator_predictions, gold["ik_predictions"], rtol=1e-9, atol=1e-12, equal_nan=True)
    np.testing.assert_allclose(ik.expected_values, gold["ik_expected"], rtol=1e-9)
    np.testing.assert_allclose(ik.variances, gold["ik_variances"], rtol=1e-7)
    maps = ik.get_indicator_maps()
    assert len(maps) == 4 and next(iter(maps.values())).shape == (25, 3)
    with pytest.raises(TypeError):
        pib.IndicatorKriging(holder, pts, gold["ik_unknown"], kriging_type="sk", process_mean=0.5)
