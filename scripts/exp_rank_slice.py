#!/usr/bin/env python
"""One rank's slice of BASELINE config 4 at 8 ranks (1.25M of the 10M locations, 1M known points, k = 64, ordinary + simple
kriging) through the public call on ONE GPU: wall time against the kernel time, per chunk size of the host pipeline.
usage: python scripts/exp_rank_slice.py [m] [k]   (k = 32: ordinary kriging with the linear model of config 1 instead)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import pyinterpolate_b200 as pib  # noqa: E402
from pyinterpolate_b200 import _lib  # noqa: E402
from pyinterpolate_b200.synthetic import dem_like_field, uniform_locations  # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 64
known = dem_like_field(1_000_000, seed=4)
unk = uniform_locations(10_000_000, seed=5)[:m]
mdl = {"variogram_model_type": "spherical", "nugget": 1.0, "sill": float(np.var(known[:, 2])), "rang": 20000.0}
mean = float(known[:, 2].mean())
for chunk in ("", "156250", "78125", "39063"):
    if chunk:
        os.environ["PIB_KRIGE_CHUNK"] = chunk
    else:
        os.environ.pop("PIB_KRIGE_CHUNK", None)
    for rep in range(3):
        t0 = time.perf_counter()
        if k == 64:
            ok, sk = pib.ordinary_and_simple_kriging(mdl, known, unk, process_mean=mean, no_neighbors=64)
        else:
            ok = pib.ordinary_kriging(mdl, known, unk, no_neighbors=k)
        t1 = time.perf_counter()
        knn, nk = _lib.last_kernel_ms(_lib.KERNEL_KNN)
        sol, ns = _lib.last_kernel_ms(_lib.KERNEL_SOLVE)
        print(f"chunk {chunk or 'default'} rep {rep}: public call {1e3 * (t1 - t0):.1f} ms = {m / (t1 - t0):.3e} locations/s; kernel sums: knn {knn:.1f} ms "
              f"({nk} launches), solve {sol:.1f} ms ({ns})", flush=True)
d_known = torch.from_numpy(known).cuda()
d_unk = torch.from_numpy(unk).cuda()
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    _lib.krige_device(d_known, d_unk, _lib.MODEL_IDS["spherical"], np.array([[1.0, mdl["sill"], 20000.0]]), 2 if k == 64 else 0, k, process_mean=mean)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"device-resident call: {1e3 * (t1 - t0):.1f} ms (knn {_lib.last_kernel_ms(_lib.KERNEL_KNN)[0]:.1f} + solve {_lib.last_kernel_ms(_lib.KERNEL_SOLVE)[0]:.1f})")
t0 = time.perf_counter(); a = np.empty((2, 1, m, 4)); a[:] = 0; t1 = time.perf_counter()
print(f"first touch of the {a.nbytes >> 20} MB output: {1e3 * (t1 - t0):.1f} ms")
t0 = time.perf_counter(); b = known.copy(); t1 = time.perf_counter()
print(f"host memcpy 24 MB: {1e3 * (t1 - t0):.1f} ms")
