"""Ad-hoc timings on the GPU box (not the bench): python scripts/quick_perf.py [n_points]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from pyinterpolate_b200 import _lib
from pyinterpolate_b200.synthetic import dem_like_field, uniform_locations

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
import os
print("fp64 peak (DFMA probe):", _lib.fp64_peak_tflops(), "TFLOP/s")
only = os.environ.get("QP_ONLY", "")
pts = dem_like_field(n, seed=2)
lags = np.arange(500.0, 40500.0, 500.0)
d = torch.from_numpy(pts).cuda()
for rep in range(0 if only == "krige" else 3):
    torch.cuda.synchronize(); t0 = time.time()
    r = _lib.variogram_device(d, lags, want_pairs=True)
    torch.cuda.synchronize(); dt = time.time() - t0
    ev = r["pairs_evaluated"]
    kms, _ = _lib.last_kernel_ms(_lib.KERNEL_PAIR)
    print(f"variogram n={n}: {dt*1e3:.1f} ms (pair kernel {kms:.1f} ms) evaluated={ev:.3e} pairs -> {ev/dt:.3e} pairs/s evaluated, "
          f"{n*(n-1)/2/dt:.3e} equivalent pairs/s, {13*ev/dt/1e12:.2f} TFLOP/s algorithmic")
if only != "krige":
    from pyinterpolate_b200 import core
    W = np.stack([core.define_whitening_matrix(a, 0.2) for a in (90, 0, 45, 135)]).reshape(-1, 4)
    d5 = torch.from_numpy(dem_like_field(n // 5, seed=3)).cuda()
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        r = _lib.variogram_device(d5, lags, lower=core.ellipse_lower_limits(lags), W=W, want_pairs=True)
        torch.cuda.synchronize(); dt = time.time() - t0
    kms, nl = _lib.last_kernel_ms(_lib.KERNEL_PAIR)
    print(f"ellipse 4 dirs n={n//5} (config 3): {dt*1e3:.1f} ms (pair kernels {kms:.1f} ms, {nl} launches) evaluated={r['pairs_evaluated']:.3e}")
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        r = _lib.variogram_device(d, lags, lower=core.ellipse_lower_limits(lags), W=W[:1], want_pairs=True)
        torch.cuda.synchronize(); dt = time.time() - t0
    print(f"ellipse 1 dir n={n}: {dt*1e3:.1f} ms evaluated={r['pairs_evaluated']:.3e}")

if only == "variogram":
    sys.exit(0)
known = d
M = int(os.environ.get("QP_M", "1000000"))
for k, m in ((32, M), (64, M)):
    unk = torch.from_numpy(uniform_locations(m, seed=5)).cuda()
    params = np.array([[1.0, float(np.var(pts[:, 2])), 20000.0]])
    for mode in (0, 1):
        for rep in range(2):
            torch.cuda.synchronize(); t0 = time.time()
            out, _, st = _lib.krige_device(known, unk, 6, params, mode, k)
            torch.cuda.synchronize(); dt = time.time() - t0
        kn, _ = _lib.last_kernel_ms(_lib.KERNEL_KNN); so, _ = _lib.last_kernel_ms(_lib.KERNEL_SOLVE)
        print(f"krige mode={mode} k={k} n={n} m={m}: {dt*1e3:.1f} ms (knn {kn:.1f} solve {so:.1f}) -> {m/dt:.3e} pts/s")
