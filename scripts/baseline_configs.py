"""BASELINE.json configs 3 and 4 at FULL size through the public calls (host arrays), with a size-independent check:
a random subset of the locations kriged on its own must give bit-identical rows (locations are independent).
    python scripts/baseline_configs.py"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import pyinterpolate_b200 as pib
from pyinterpolate_b200.synthetic import dem_like_field, uniform_locations


def timed(label, fn):
    t0 = time.time(); out = fn(); dt = time.time() - t0
    print(f"{label}: {dt:.2f} s")
    return out, dt


# config 3: OK + SK of 10M unknown locations from 1M known points, k = 64
known = dem_like_field(1_000_000, seed=0)
unk = uniform_locations(10_000_000, seed=1)
var = float(np.var(known[:, 2]))
mdl = {"variogram_model_type": "spherical", "nugget": 1.0, "sill": var, "rang": 5000.0}
sub = np.random.default_rng(2).choice(len(unk), 2000, replace=False)
ok, dt = timed("config 3: ordinary kriging, 1M known, 10M locations, k=64", lambda: pib.ordinary_kriging(mdl, known, unk, no_neighbors=64))
print(f"   {len(unk) / dt:.3e} points/s through the public call; NaN variances: {int(np.isnan(ok[:, 1]).sum())}")
ok_sub = pib.ordinary_kriging(mdl, known, unk[sub], no_neighbors=64)
assert np.array_equal(ok[sub], ok_sub, equal_nan=True), "subset differs"
sk, dt = timed("config 3: simple kriging, 1M known, 10M locations, k=64", lambda: pib.simple_kriging(mdl, known, unk, process_mean=float(known[:, 2].mean()), no_neighbors=64))
print(f"   {len(unk) / dt:.3e} points/s")
sk_sub = pib.simple_kriging(mdl, known, unk[sub], process_mean=float(known[:, 2].mean()), no_neighbors=64)
assert np.array_equal(sk[sub], sk_sub, equal_nan=True), "subset differs"
del ok, sk, unk

# config 4: indicator kriging with 8 thresholds on 500k points
known = dem_like_field(500_000, seed=3)
unk = uniform_locations(500_000, seed=4)
thr = np.quantile(known[:, 2], np.linspace(0.1, 0.9, 8))
models = [{"variogram_model_type": "spherical", "nugget": 0.02, "sill": 0.2, "rang": 4000.0 + 250.0 * t} for t in range(8)]
ik, dt = timed("config 4: indicator kriging, 8 thresholds, 500k known x 500k locations, k=16",
               lambda: pib.kriging.indicator_kriging(models, thr, known, unk, no_neighbors=16, reference_variance_column=False))
print(f"   {len(unk) * 8 / dt:.3e} threshold-solves/s; range of predictions [{ik.min():.3f}, {ik.max():.3f}]")
sub = np.random.default_rng(5).choice(len(unk), 2000, replace=False)
ik_sub = pib.kriging.indicator_kriging(models, thr, known, unk[sub], no_neighbors=16, reference_variance_column=False)
assert np.array_equal(ik[sub], ik_sub), "subset differs"
assert np.all(np.diff(ik.mean(axis=0)) > 0), "mean indicator probability must grow with the threshold"
print("all checks passed")
