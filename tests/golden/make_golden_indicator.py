"""Indicator variogram / indicator kriging fixtures from the UNMODIFIED reference
(see make_golden.py for how the reference is imported).   python tests/golden/make_golden_indicator.py"""
import os
import sys
import types
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
warnings.filterwarnings("ignore")
from ref_import import import_reference  # noqa: E402

import_reference()
from pyinterpolate import TheoreticalVariogram  # noqa: E402
from pyinterpolate.kriging.point.indicator import IndicatorKriging  # noqa: E402
from pyinterpolate.semivariogram.indicator.indicator import ExperimentalIndicatorVariogram  # noqa: E402

from pyinterpolate_b200.synthetic import dem_like_field, uniform_locations  # noqa: E402


def main():
    g = {}
    pts = dem_like_field(1200, seed=41)
    eiv = ExperimentalIndicatorVariogram(pts, number_of_thresholds=4, step_size=2500, max_range=40000)
    g["thresholds"] = np.array(eiv.ds.thresholds)
    g["keys"] = np.array(list(eiv.experimental_models.keys()))
    g["lags"] = next(iter(eiv.experimental_models.values())).lags
    g["sem"] = np.array([m.semivariances for m in eiv.experimental_models.values()])
    g["cov"] = np.array([m.covariances for m in eiv.experimental_models.values()])
    g["ppl"] = np.array([m.points_per_lag for m in eiv.experimental_models.values()])
    g["var"] = np.array([m.variance for m in eiv.experimental_models.values()])
    # indicator kriging with fixed spherical models per threshold
    models = {}
    for k, p in zip(eiv.experimental_models.keys(), (0.25, 0.5, 0.75, 1.0)):
        m = TheoreticalVariogram()
        m.from_dict({"variogram_model_type": "spherical", "nugget": 0.01, "sill": max(p * (1 - p), 0.05), "rang": 20000.0})
        models[k] = m
    holder = types.SimpleNamespace(theoretical_indicator_variograms=models)
    unk = uniform_locations(25, seed=42)
    ik = IndicatorKriging(holder, pts, unk, kriging_type="ok", no_neighbors=16)
    g["ik_unknown"] = unk
    g["ik_predictions"] = ik.indicator_predictions
    g["ik_expected"] = ik.expected_values
    g["ik_variances"] = ik.variances
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden_indicator.npz"), **g)
    print("wrote", len(g), "arrays")


if __name__ == "__main__":
    main()
