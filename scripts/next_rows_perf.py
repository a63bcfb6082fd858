"""Timings of the §8f rows (not part of bench.py): python scripts/next_rows_perf.py"""
import sys, time, warnings
import numpy as np
sys.path.insert(0, ".")
import pyinterpolate_b200 as pib
from pyinterpolate_b200.synthetic import dem_like_field, uniform_locations

warnings.filterwarnings("ignore")


def timed(label, fn, reps=2):
    for _ in range(reps):
        t0 = time.time(); out = fn(); dt = time.time() - t0
    print(f"{label}: {dt * 1e3:.1f} ms")
    return out


known = dem_like_field(200_000, seed=5)
unk = uniform_locations(200_000, seed=6)
var = float(np.var(known[:, 2]))
iso = {"variogram_model_type": "spherical", "nugget": 0.5, "sill": var, "rang": 3000.0}
dirm = dict(iso, direction=30.0)
timed("ordinary kriging k=16, 200k x 200k, isotropic", lambda: pib.ordinary_kriging(iso, known, unk, no_neighbors=16))
timed("ordinary kriging k=16, 200k x 200k, directional model (30 deg)", lambda: pib.ordinary_kriging(dirm, known, unk, no_neighbors=16))
timed("ordinary kriging k=16, 200k x 200k, use_all_neighbors_in_range (range 800)",
      lambda: pib.ordinary_kriging(iso, known, unk, no_neighbors=16, neighbors_range=800.0, use_all_neighbors_in_range=True))
thr = np.quantile(known[:, 2], [0.2, 0.4, 0.6, 0.8])
models = [{"variogram_model_type": "spherical", "nugget": 0.02, "sill": 0.2, "rang": 3000.0} for _ in thr]
timed("indicator kriging, 4 thresholds, k=16, 200k x 200k", lambda: pib.kriging.indicator_kriging(models, thr, known, unk, no_neighbors=16))
pts = dem_like_field(30_000, seed=7)
timed("variogram cloud, 30k points, 20 lags to 10 km", lambda: pib.ExperimentalVariogram(pts, 500, 10_000, as_cloud=True))
timed("directional variogram 't' (default), 4 directions + ISO, 100k points", lambda: pib.DirectionalVariogram(known[:100_000], step_size=500, max_range=40_500, tolerance=0.2))
timed("directional variogram 'e', 4 directions + ISO, 100k points", lambda: pib.DirectionalVariogram(known[:100_000], step_size=500, max_range=40_500, tolerance=0.2, dir_neighbors_selection_method="e"))
big = dem_like_field(1_000_000, seed=8)
mb = {"variogram_model_type": "spherical", "nugget": 0.5, "sill": float(np.var(big[:, 2])), "rang": 2000.0}
timed("validate_kriging (leave-one-out), 1M points, k=16", lambda: pib.validate_kriging(big, mb, no_neighbors=16))
timed("interpolate_raster, 1M known points, 1000 x 1000 pixels, k=16", lambda: pib.interpolate_raster(big, dim=1000, number_of_neighbors=16, semivariogram_model=mb))
