"""BASELINE config 1 (the reference's own CPU-runnable case) through the public calls: 10k points, step 500,
max_range 40000, ordinary kriging k=32 of 10k locations.   python scripts/config1_latency.py"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import pyinterpolate_b200 as pib
from pyinterpolate_b200.synthetic import dem_like_field, uniform_locations

pts = dem_like_field(10_000, seed=0)
unk = uniform_locations(10_000, seed=1)
mdl = {"variogram_model_type": "linear", "nugget": 0.0, "sill": float(np.var(pts[:, 2])), "rang": 20_000.0}
for rep in range(3):
    t0 = time.time(); ev = pib.ExperimentalVariogram(pts, step_size=500, max_range=40_000); t1 = time.time()
    ok = pib.ordinary_kriging(mdl, pts, unk, no_neighbors=32); t2 = time.time()
print(f"ExperimentalVariogram (semivariance + covariance), 10k points, 79 lags: {(t1 - t0) * 1e3:.2f} ms (reference: 14.5 s)")
print(f"ordinary_kriging, 10k known, 10k locations, k=32: {(t2 - t1) * 1e3:.2f} ms (reference: ~7.5 s)")
