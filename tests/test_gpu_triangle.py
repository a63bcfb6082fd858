"""GPU parity for the triangle selection method (the reference's DEFAULT for directional variograms,
SURVEY §8f-4) against the unmodified reference: pair counts bit-exact — including the self pairs the
reference counts in the first lag and points that lie exactly on triangle edges of a regular grid."""
import os

import numpy as np
import pytest

from pyinterpolate_b200.synthetic import dem_like_field

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
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_golden_triangle.npz"))


def _check(got, want, cov_scale=None):
    np.testing.assert_array_equal(got[:, 0], want[:, 0])
    np.testing.assert_array_equal(got[:, 2], want[:, 2])
    if cov_scale is None:
        np.testing.assert_allclose(got[:, 1], want[:, 1], rtol=RTOL, atol=1e-12)
    else:
        np.testing.assert_allclose(got[:, 1], want[:, 1], rtol=0, atol=RTOL * cov_scale)


def test_armstrong_triangle_golden(pib, golden, gold):
    arm = golden["armstrong_points"]
    scale = float(np.mean(arm[:, 2]) ** 2 + np.var(arm[:, 2]))
    _check(pib.calculate_semivariance(arm, step_size=1.5, max_range=6, direction=15, tolerance=0.25), gold["arm_sem"])
    _check(pib.calculate_covariance(arm, step_size=1.5, max_range=6, direction=15, tolerance=0.25), gold["arm_cov"], scale)
    for ang in (0, 45, 90, 135):                 # grid-aligned: many points exactly on triangle edges
        _check(pib.calculate_semivariance(arm, step_size=1, max_range=6, direction=ang, tolerance=0.5,
                                          dir_neighbors_selection_method="t"), gold[f"arm_grid_{ang}"])


@pytest.mark.parametrize("ang,tol", [(30, 0.3), (90, 0.2), (0, 0.2), (45, 0.2), (135, 0.2), (200, 1.0), (-60, 0.05)])
def test_synthetic_triangle_golden(pib, gold, ang, tol):
    pts = dem_like_field(800, seed=91)
    scale = float(np.mean(pts[:, 2]) ** 2 + np.var(pts[:, 2]))
    _check(pib.calculate_semivariance(pts, step_size=4000, max_range=40000, direction=ang, tolerance=tol),
           gold[f"f800_{ang}_{tol}_sem"])
    _check(pib.calculate_covariance(pts, step_size=4000, max_range=40000, direction=ang, tolerance=tol),
           gold[f"f800_{ang}_{tol}_cov"], scale)


def test_default_method_classes_golden(pib, gold):
    small = dem_like_field(300, extent=10.0, seed=92)
    _check(pib.calculate_semivariance(small, step_size=0.4, max_range=4.1, direction=77, tolerance=0.4), gold["f300_sem"])
    pts = dem_like_field(800, seed=91)
    scale = float(np.mean(pts[:, 2]) ** 2 + np.var(pts[:, 2]))
    ev = pib.ExperimentalVariogram(pts, step_size=4000, max_range=40000, direction=30, tolerance=0.3)
    np.testing.assert_array_equal(ev.points_per_lag, gold["ev_ppl"])
    np.testing.assert_allclose(ev.semivariances, gold["ev_sem"], rtol=RTOL)
    np.testing.assert_allclose(ev.covariances, gold["ev_cov"], rtol=0, atol=RTOL * scale)
    dv = pib.DirectionalVariogram(pts, step_size=4000, max_range=40000, tolerance=0.2)       # default method 't'
    for name in ("ISO", "NS", "WE", "NE-SW", "NW-SE"):
        v, want = dv.get(name), gold[f"dv_{name}"]
        np.testing.assert_array_equal(v.points_per_lag, want[:, 3])
        np.testing.assert_allclose(v.semivariances, want[:, 1], rtol=RTOL)
        np.testing.assert_allclose(v.covariances, want[:, 2], rtol=0, atol=RTOL * scale)
