"""Variogram point-cloud fixtures from the UNMODIFIED reference.   python tests/golden/make_golden_cloud.py"""
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
from pyinterpolate.semivariogram.experimental.experimental_semivariogram import point_cloud_semivariance  # noqa: E402

from pyinterpolate_b200.synthetic import dem_like_field  # noqa: E402


def pack(cloud):
    lags = np.array(list(cloud.keys()), dtype=float)
    vals = [np.zeros(0) if v is None else np.asarray(v, dtype=float) for v in cloud.values()]    # (triangle cloud: None = empty lag)
    sizes = np.array([len(v) for v in vals], dtype=np.int64)
    flat = np.concatenate(vals) if sizes.sum() else np.zeros(0)
    return lags, sizes, flat


def main():
    g = {}
    arm = np.load("/root/reference/tests/test_semivariogram/armstrong_data.npy")
    g["arm_lags"], g["arm_sizes"], g["arm_values"] = pack(point_cloud_semivariance(arm, step_size=1, max_range=6))
    g["arm_e_lags"], g["arm_e_sizes"], g["arm_e_values"] = pack(point_cloud_semivariance(
        arm, step_size=1.5, max_range=6, direction=15, tolerance=0.25, dir_neighbors_selection_method="e"))
    pts = dem_like_field(350, seed=51)
    g["f350_lags"], g["f350_sizes"], g["f350_values"] = pack(point_cloud_semivariance(pts, step_size=4000, max_range=40000))
    g["f350_e_lags"], g["f350_e_sizes"], g["f350_e_values"] = pack(point_cloud_semivariance(
        pts, step_size=4000, max_range=40000, direction=45, tolerance=0.3, dir_neighbors_selection_method="e"))
    sparse = dem_like_field(30, extent=1000.0, seed=52)          # has empty lags
    g["sparse_lags"], g["sparse_sizes"], g["sparse_values"] = pack(point_cloud_semivariance(sparse, step_size=5.0, max_range=100.0))
    # the reference's DEFAULT directional method: triangle masks (from_triangle_cloud, directional.py:213-273)
    g["arm_t_lags"], g["arm_t_sizes"], g["arm_t_values"] = pack(point_cloud_semivariance(
        arm, step_size=1.5, max_range=6, direction=15, tolerance=0.25))
    g["arm_t90_lags"], g["arm_t90_sizes"], g["arm_t90_values"] = pack(point_cloud_semivariance(
        arm, step_size=1, max_range=6, direction=90, tolerance=0.5, dir_neighbors_selection_method="triangle"))
    g["f350_t_lags"], g["f350_t_sizes"], g["f350_t_values"] = pack(point_cloud_semivariance(
        pts, step_size=4000, max_range=40000, direction=45, tolerance=0.3))
    g["sparse_t_lags"], g["sparse_t_sizes"], g["sparse_t_values"] = pack(point_cloud_semivariance(
        sparse, step_size=5.0, max_range=100.0, direction=135, tolerance=0.1))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden_cloud.npz"), **g)
    print("wrote", len(g), "arrays", {k: int(v.sum()) for k, v in g.items() if k.endswith("sizes")})


if __name__ == "__main__":
    main()
