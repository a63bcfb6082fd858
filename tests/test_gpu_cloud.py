"""GPU parity for the variogram point cloud against the unmodified reference: per lag the
same number of values, in the same order, bit-exact (they are single fp64 operations: (z_i - z_j)^2)."""
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


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_golden_cloud.npz"))


def _check(cloud, gold, name):
    lags, sizes, flat = gold[f"{name}_lags"], gold[f"{name}_sizes"], gold[f"{name}_values"]
    np.testing.assert_array_equal(np.array(list(cloud.keys()), dtype=float), lags)
    np.testing.assert_array_equal([len(v) for v in cloud.values()], sizes)
    got = np.concatenate([np.asarray(v, dtype=float) for v in cloud.values()]) if sizes.sum() else np.zeros(0)
    np.testing.assert_array_equal(got, flat)                     # same values, same order


def test_cloud_golden(pib, golden, gold):
    arm = golden["armstrong_points"]
    _check(pib.point_cloud_semivariance(arm, step_size=1, max_range=6), gold, "arm")
    _check(pib.point_cloud_semivariance(arm, step_size=1.5, max_range=6, direction=15, tolerance=0.25,
                                        dir_neighbors_selection_method="e"), gold, "arm_e")
    pts = dem_like_field(350, seed=51)
    _check(pib.point_cloud_semivariance(pts, step_size=4000, max_range=40000), gold, "f350")
    _check(pib.point_cloud_semivariance(pts, step_size=4000, max_range=40000, direction=45, tolerance=0.3,
                                        dir_neighbors_selection_method="e"), gold, "f350_e")
    sparse = dem_like_field(30, extent=1000.0, seed=52)
    cloud = pib.point_cloud_semivariance(sparse, step_size=5.0, max_range=100.0)
    _check(cloud, gold, "sparse")
    assert any(isinstance(v, list) and v == [] for v in cloud.values())        # empty lags are [] like the reference


def test_triangle_cloud_golden(pib, golden, gold):
    """The reference's DEFAULT directional method (triangle masks, functions/directional.py:213-273): same values in the
    same order, the q == p pairs of the first lag included; empty lags (None in the reference) come back as []."""
    arm = golden["armstrong_points"]
    _check(pib.point_cloud_semivariance(arm, step_size=1.5, max_range=6, direction=15, tolerance=0.25), gold, "arm_t")
    _check(pib.point_cloud_semivariance(arm, step_size=1, max_range=6, direction=90, tolerance=0.5,
                                        dir_neighbors_selection_method="triangle"), gold, "arm_t90")
    pts = dem_like_field(350, seed=51)
    _check(pib.point_cloud_semivariance(pts, step_size=4000, max_range=40000, direction=45, tolerance=0.3), gold, "f350_t")
    sparse = dem_like_field(30, extent=1000.0, seed=52)
    cloud = pib.point_cloud_semivariance(sparse, step_size=5.0, max_range=100.0, direction=135, tolerance=0.1)
    _check(cloud, gold, "sparse_t")
    assert sum(1 for v in cloud.values() if isinstance(v, list) and v == []) == 18
    # the cloud of the default method agrees with the aggregated triangle variogram (counts, mean / 2)
    pts = dem_like_field(2000, seed=54)
    vc = pib.VariogramCloud(pts, step_size=4000, max_range=40000, direction=30, tolerance=0.4)
    ev = pib.ExperimentalVariogram(pts, step_size=4000, max_range=40000, direction=30, tolerance=0.4)
    np.testing.assert_array_equal(vc.points_per_lag, ev.points_per_lag)
    np.testing.assert_allclose(vc.average_semivariance()[:, 1], ev.semivariances, rtol=1e-12)


def test_cloud_is_consistent_with_the_aggregated_variogram(pib):
    """The reference's own consistency test (test_experimental_variogram_point_cloud.py:42-64, :125-150):
    cloud lengths equal the pair counts and mean(cloud) / 2 equals the semivariance."""
    pts = dem_like_field(3000, seed=53)
    for kw in ({}, dict(direction=90, tolerance=0.2, dir_neighbors_selection_method="e")):
        vc = pib.VariogramCloud(pts, step_size=4000, max_range=40000, **kw)
        ev = pib.ExperimentalVariogram(pts, step_size=4000, max_range=40000, **kw)
        np.testing.assert_array_equal(vc.points_per_lag, ev.points_per_lag)
        avg = vc.average_semivariance()
        np.testing.assert_allclose(avg[:, 1], ev.semivariances, rtol=1e-12)
        d = vc.describe()
        assert d[vc.lags[0]]["count"] == ev.points_per_lag[0]
