"""Triangle-method (dir_neighbors_selection_method='t', the reference's default) fixtures from the UNMODIFIED
reference.   python tests/golden/make_golden_triangle.py"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
warnings.filterwarnings("ignore")
from ref_import import import_reference  # noqa: E402

import_reference()
from pyinterpolate import DirectionalVariogram, ExperimentalVariogram, calculate_covariance, calculate_semivariance  # noqa: E402

from pyinterpolate_b200.synthetic import dem_like_field  # noqa: E402


def main():
    g = {}
    arm = np.load("/root/reference/tests/test_semivariogram/armstrong_data.npy")
    g["arm_sem"] = calculate_semivariance(arm, step_size=1.5, max_range=6, direction=15, tolerance=0.25)
    g["arm_cov"] = calculate_covariance(arm, step_size=1.5, max_range=6, direction=15, tolerance=0.25)
    # grid-aligned directions on the regular grid: points sit exactly on triangle edges
    for ang in (0, 45, 90, 135):
        g[f"arm_grid_{ang}"] = calculate_semivariance(arm, step_size=1, max_range=6, direction=ang, tolerance=0.5)
    pts = dem_like_field(800, seed=91)
    for ang, tol in ((30, 0.3), (90, 0.2), (0, 0.2), (45, 0.2), (135, 0.2), (200, 1.0), (-60, 0.05)):
        g[f"f800_{ang}_{tol}_sem"] = calculate_semivariance(pts, step_size=4000, max_range=40000, direction=ang, tolerance=tol)
        g[f"f800_{ang}_{tol}_cov"] = calculate_covariance(pts, step_size=4000, max_range=40000, direction=ang, tolerance=tol)
    small = dem_like_field(300, extent=10.0, seed=92)
    g["f300_sem"] = calculate_semivariance(small, step_size=0.4, max_range=4.1, direction=77, tolerance=0.4)
    ev = ExperimentalVariogram(pts, step_size=4000, max_range=40000, direction=30, tolerance=0.3)     # default method
    g["ev_sem"], g["ev_cov"], g["ev_ppl"] = ev.semivariances, ev.covariances, ev.points_per_lag
    dv = DirectionalVariogram(pts, step_size=4000, max_range=40000, tolerance=0.2)                     # default method
    for name in ("ISO", "NS", "WE", "NE-SW", "NW-SE"):
        v = dv.get(name)
        g[f"dv_{name}"] = np.column_stack([v.lags, v.semivariances, v.covariances, v.points_per_lag])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden_triangle.npz"), **g)
    print("wrote", len(g), "arrays"); print(g["arm_sem"]); print(g["f800_30_0.3_sem"][:3]); print(g["arm_grid_45"])


if __name__ == "__main__":
    main()
