"""GPU parity for custom_weights (weighted semivariance) against the unmodified reference."""
import os

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


def _check(got, want):
    np.testing.assert_array_equal(got[:, 0], want[:, 0])
    np.testing.assert_array_equal(got[:, 2], want[:, 2])                         # pair counts bit-exact
    # the directional formula subtracts a mean value from a mean square: tolerance relative to the larger term
    scale = np.maximum(np.abs(want[:, 1]), 1.0)
    np.testing.assert_allclose(got[:, 1], want[:, 1], rtol=0, atol=1e-9 * scale.max())


def test_weighted_semivariance_golden(pib, golden):
    gold = np.load(os.path.join(ROOT, "tests", "golden", "reference_golden_weighted.npz"))
    arm = golden["armstrong_points"]
    _check(pib.calculate_semivariance(arm, step_size=1, max_range=6, custom_weights=gold["arm_w"]), gold["arm_omni"])
    _check(pib.calculate_semivariance(arm, step_size=1.5, max_range=6, direction=15, tolerance=0.25,
                                      custom_weights=gold["arm_w"]), gold["arm_dir"])
    pts = dem_like_field(500, seed=72)
    w = gold["f500_w"]
    _check(pib.calculate_semivariance(pts, step_size=4000, max_range=40000, custom_weights=w), gold["f500_omni"])
    _check(pib.calculate_semivariance(pts, step_size=4000, max_range=40000, direction=120, tolerance=0.4,
                                      custom_weights=w), gold["f500_dir"])
    ev = pib.ExperimentalVariogram(pts, step_size=4000, max_range=40000, custom_weights=w)
    np.testing.assert_array_equal(ev.points_per_lag, gold["f500_ev_ppl"])
    np.testing.assert_allclose(ev.semivariances, gold["f500_ev_sem"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(ev.covariances, gold["f500_ev_cov"], rtol=0, atol=1e-9 * float(np.mean(pts[:, 2]) ** 2))
    with pytest.raises(ValueError):
        pib.calculate_semivariance(arm, step_size=1, max_range=6, custom_weights=np.zeros(len(arm)))
    with pytest.raises(IndexError):
        pib.calculate_semivariance(arm, step_size=1, max_range=6, custom_weights=np.ones(3))
