"""Weighted-semivariance fixtures (custom_weights) from the UNMODIFIED reference.
python tests/golden/make_golden_weighted.py"""
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
from pyinterpolate import ExperimentalVariogram, calculate_semivariance  # noqa: E402

from pyinterpolate_b200.synthetic import dem_like_field  # noqa: E402


def main():
    g = {}
    arm = np.load("/root/reference/tests/test_semivariogram/armstrong_data.npy")
    w_arm = np.random.default_rng(71).uniform(0.5, 3.0, size=len(arm))
    g["arm_w"] = w_arm
    g["arm_omni"] = calculate_semivariance(arm, step_size=1, max_range=6, custom_weights=w_arm)
    g["arm_dir"] = calculate_semivariance(arm, step_size=1.5, max_range=6, direction=15, tolerance=0.25,
                                          custom_weights=w_arm)
    pts = dem_like_field(500, seed=72)
    w = np.random.default_rng(73).uniform(10.0, 500.0, size=len(pts))
    g["f500_w"] = w
    g["f500_omni"] = calculate_semivariance(pts, step_size=4000, max_range=40000, custom_weights=w)
    g["f500_dir"] = calculate_semivariance(pts, step_size=4000, max_range=40000, direction=120, tolerance=0.4,
                                           custom_weights=w)
    ev = ExperimentalVariogram(pts, step_size=4000, max_range=40000, custom_weights=w)
    g["f500_ev_sem"], g["f500_ev_cov"], g["f500_ev_ppl"] = ev.semivariances, ev.covariances, ev.points_per_lag
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden_weighted.npz"), **g)
    print("wrote", len(g), "arrays"); print(g["arm_omni"]); print(g["f500_dir"][:3])


if __name__ == "__main__":
    main()
