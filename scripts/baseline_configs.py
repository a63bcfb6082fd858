"""BASELINE.json configs 4 and 5 at FULL size through the public calls (host arrays), with SURVEY section 8d's definitions
(the same ones tests/test_gpu_baseline_configs.py checks against the oracle), timed, plus a size-independent check: a random
subset of the locations kriged on its own must give bit-identical rows (locations are independent).
    python scripts/baseline_configs.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyinterpolate_b200 as pib  # noqa: E402
from pyinterpolate_b200 import indicator  # noqa: E402
from pyinterpolate_b200.synthetic import dem_like_field, uniform_locations  # noqa: E402


def timed(label, fn, reps=2):
    best, out = None, None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    print(f"{label}: {best:.3f} s", flush=True)
    return out, best


# ---- config 4: known N = 1e6 (seed 4), unknown M = 1e7 (seed 5), no_neighbors = 64, spherical {nugget 1, sill var(z), rang 20 000},
# ordinary AND simple kriging (process_mean = mean(z))
known = dem_like_field(1_000_000, extent=100_000.0, seed=4)
unk = uniform_locations(10_000_000, extent=100_000.0, seed=5)
var, mean = float(np.var(known[:, 2])), float(np.mean(known[:, 2]))
mdl = {"variogram_model_type": "spherical", "nugget": 1.0, "sill": var, "rang": 20000.0}
(ok, sk), dt = timed("config 4: ordinary + simple kriging from one call, 1M known, 10M locations, k=64",
                     lambda: pib.ordinary_and_simple_kriging(mdl, known, unk, process_mean=mean, no_neighbors=64))
print(f"   {len(unk) / dt:.3e} locations/s through the public call; NaN variances: {int(np.isnan(ok[:, 1]).sum())} / {int(np.isnan(sk[:, 1]).sum())}")
sub = np.sort(np.random.default_rng(2).choice(len(unk), 2000, replace=False))
ok_sub = pib.ordinary_kriging(mdl, known, unk[sub], no_neighbors=64)
sk_sub = pib.simple_kriging(mdl, known, unk[sub], process_mean=mean, no_neighbors=64)
assert np.array_equal(ok[sub], ok_sub, equal_nan=True) and np.array_equal(sk[sub], sk_sub, equal_nan=True), "subset differs"
_, dt1 = timed("config 4, the two single calls: ordinary_kriging", lambda: pib.ordinary_kriging(mdl, known, unk, no_neighbors=64), reps=1)
_, dt2 = timed("config 4, the two single calls: simple_kriging", lambda: pib.simple_kriging(mdl, known, unk, process_mean=mean, no_neighbors=64), reps=1)
print(f"   shared call {dt:.3f} s against {dt1 + dt2:.3f} s for the two calls")
del ok, sk, unk

# ---- config 5: N = 500 000 (seed 6), thresholds = 8 quantiles (semivariogram/indicator/indicator.py:47-71), 8 spherical models
# {nugget 0.01, sill p (1 - p), rang 20 000}, M = 1e6 unknown (seed 7), no_neighbors = 32, ordinary kriging of the indicators
known = dem_like_field(500_000, extent=100_000.0, seed=6)
unk = uniform_locations(1_000_000, extent=100_000.0, seed=7)
thresholds = indicator.select_variogram_thresholds(known[:, 2], 8)
probs = np.linspace(0, 1, 9)[1:]
models = [{"variogram_model_type": "spherical", "nugget": 0.01, "sill": max(float(p * (1 - p)), 1e-3), "rang": 20000.0} for p in probs]
ik, dt = timed("config 5: indicator kriging, 8 quantile thresholds, 500k known x 1M locations, k=32",
               lambda: pib.kriging.indicator_kriging(models, thresholds, known, unk, no_neighbors=32, reference_variance_column=False))
print(f"   {len(unk) / dt:.3e} locations/s = {len(unk) * 8 / dt:.3e} threshold solves/s; range of predictions [{ik.min():.3f}, {ik.max():.3f}]")
sub = np.sort(np.random.default_rng(5).choice(len(unk), 2000, replace=False))
ik_sub = pib.kriging.indicator_kriging(models, thresholds, known, unk[sub], no_neighbors=32, reference_variance_column=False)
assert np.array_equal(ik[sub], ik_sub), "subset differs"
assert np.all(np.diff(ik.mean(axis=0)) > 0), "mean indicator probability must grow with the threshold"
print("all checks passed")
