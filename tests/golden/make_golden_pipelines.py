"""Rank-1 callers of the kriging path (interpolate_raster, UniversalKriging, interpolate_points_dask) — fixtures
from the UNMODIFIED reference.   python tests/golden/make_golden_pipelines.py"""
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
from pyinterpolate import TheoreticalVariogram  # noqa: E402
from pyinterpolate.core.pipelines.interpolate import interpolate_points_dask  # noqa: E402
from pyinterpolate.kriging.point.universal import UniversalKriging  # noqa: E402
from pyinterpolate.viz.raster import interpolate_raster  # noqa: E402

from pyinterpolate_b200.synthetic import dem_like_field, uniform_locations  # noqa: E402


def main():
    g = {}
    known = dem_like_field(3000, extent=20_000.0, seed=91)
    known[:, 2] += 0.004 * known[:, 0] - 0.002 * known[:, 1]            # a planar trend for universal kriging
    var = float(np.var(known[:, 2]))
    g["known"] = known
    m = TheoreticalVariogram()
    m.from_dict({"variogram_model_type": "spherical", "nugget": 0.3, "sill": var, "rang": 6000.0})
    r = interpolate_raster(known, dim=37, number_of_neighbors=8, semivariogram_model=m)
    g["raster_result"], g["raster_error"] = r["result"], r["error"]
    g["raster_params"] = np.array([r["params"][k] for k in ("pixel size", "min x", "max x", "min y", "max y")])
    unk = uniform_locations(150, extent=20_000.0, seed=92)
    g["unknown"] = unk
    g["dask1"] = interpolate_points_dask(m, known, unk, no_neighbors=12, number_of_workers=1, progress_bar=False)
    uk = UniversalKriging(known)
    uk.fit_trend()
    uk.detrend()
    mb = TheoreticalVariogram()
    mb.from_dict({"variogram_model_type": "exponential", "nugget": 0.5, "sill": float(np.var(uk.bias_values)), "rang": 5000.0})
    uk.bias_model = mb
    g["uk_bias_sill"] = np.array(float(np.var(uk.bias_values)))
    g["uk_coefficients"] = np.append(uk.trend_model.coefficients, uk.trend_model.intercept)
    g["uk_pred"] = uk.predict(unk, no_neighbors=10, show_progress_bar=False)
    # leave-one-out cross-validation (evaluate/cross_validation.py:11-138)
    from pyinterpolate.evaluate.cross_validation import validate_kriging
    cv_pts = dem_like_field(600, extent=8_000.0, seed=94)
    g["cv_points"] = cv_pts
    mcv = TheoreticalVariogram()
    mcv.from_dict({"variogram_model_type": "spherical", "nugget": 0.4, "sill": float(np.var(cv_pts[:, 2])), "rang": 2500.0})
    for how, kw in (("ok", {}), ("sk", {"sk_mean": float(cv_pts[:, 2].mean())})):
        mpe, mve, arr = validate_kriging(cv_pts, mcv, how=how, no_neighbors=8, **kw)
        g[f"cv_{how}_stats"] = np.array([mpe, mve])
        g[f"cv_{how}_arr"] = arr
    mpe, mve, arr = validate_kriging(cv_pts[:40], mcv, how="ok", no_neighbors=64)       # fewer points than neighbours
    g["cv_small_arr"] = arr
    # VariogramCloud.remove_outliers (classes/variogram_cloud.py:334-397 -> transform/statistical.py:165-247)
    from pyinterpolate.transform.statistical import remove_outliers
    rng = np.random.default_rng(93)
    cloud = {500.0: rng.gamma(2.0, 3.0, 400), 1000.0: np.append(rng.normal(10, 1, 300), [40.0, -25.0]),
             1500.0: rng.exponential(5.0, 50)}
    for lag, v in cloud.items():
        g[f"cloud_in_{int(lag)}"] = v
    for name, kw in (("z", dict(method="zscore")), ("z2", dict(method="zscore", z_lower_limit=-1.5, z_upper_limit=2.0)),
                     ("iqr", dict(method="iqr")), ("iqr2", dict(method="iqr", iqr_lower_limit=0.5, iqr_upper_limit=1.0))):
        out = remove_outliers(cloud, **kw)
        for lag, v in out.items():
            g[f"cloud_{name}_{int(lag)}"] = v
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden_pipelines.npz"), **g)
    print(g["raster_result"].shape, g["uk_pred"][:2], g["uk_coefficients"])


if __name__ == "__main__":
    main()
