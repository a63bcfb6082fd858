#!/usr/bin/env python
"""Public kriging calls with host arrays under a setting of the chunk pipeline (PIB_KRIGE_LANES, PIB_KRIGE_CHUNK): time of the
k = 32 call on 10M / 1.25M / 500k locations and of BASELINE config 5 (indicator kriging).
usage: PIB_KRIGE_LANES=n python scripts/exp_public_calls.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyinterpolate_b200 as pib  # noqa: E402
from pyinterpolate_b200 import indicator  # noqa: E402
from pyinterpolate_b200.synthetic import dem_like_field, uniform_locations  # noqa: E402

known = dem_like_field(1_000_000, seed=2)
unk = uniform_locations(10_000_000, seed=5)
var = float(np.var(known[:, 2]))
mdl = {"variogram_model_type": "linear", "nugget": 0.0, "sill": var, "rang": 8000.0}
print("PIB_KRIGE_LANES =", os.environ.get("PIB_KRIGE_LANES", "(default)"), " PIB_KRIGE_CHUNK =", os.environ.get("PIB_KRIGE_CHUNK", "(default)"))
for m in (10_000_000, 1_250_000, 500_000):
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        pib.ordinary_kriging(mdl, known, unk[:m], no_neighbors=32)
        best = min(best, time.perf_counter() - t0)
    print(f"  ordinary_kriging k=32, {m} locations: {1e3 * best:.1f} ms = {m / best:.3e} points/s", flush=True)
k5 = dem_like_field(500_000, extent=100_000.0, seed=6)
u5 = uniform_locations(1_000_000, extent=100_000.0, seed=7)
thr = indicator.select_variogram_thresholds(k5[:, 2], 8)
models = [{"variogram_model_type": "spherical", "nugget": 0.01, "sill": max(float(p * (1 - p)), 1e-3), "rang": 20000.0} for p in np.linspace(0, 1, 9)[1:]]
best = 1e9
for rep in range(3):
    t0 = time.perf_counter()
    pib.kriging.indicator_kriging(models, thr, k5, u5, no_neighbors=32, reference_variance_column=False)
    best = min(best, time.perf_counter() - t0)
print(f"  config 5 (indicator kriging, 8 thresholds, 1M locations): {1e3 * best:.1f} ms")
