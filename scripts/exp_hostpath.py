#!/usr/bin/env python
"""Where the public kriging call spends its time (host arrays in / out): python scripts/exp_hostpath.py"""
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

pts = dem_like_field(1_000_000, seed=2)
unk = uniform_locations(4_000_000, seed=5)
var = float(np.var(pts[:, 2]))
mdl = {"variogram_model_type": "linear", "nugget": 0.0, "sill": var, "rang": 8000.0}
params = np.array([[0.0, var, 8000.0]])
for chunk in ("", "524288", "2097152"):
    if chunk:
        os.environ["PIB_KRIGE_CHUNK"] = chunk
    else:
        os.environ.pop("PIB_KRIGE_CHUNK", None)
    for rep in range(3):
        t0 = time.perf_counter()
        out, _, st = _lib.krige_host(pts, unk, _lib.MODEL_IDS["linear"], params, _lib.KRIGE_ORDINARY, 32)
        t1 = time.perf_counter()
        res = pib.ordinary_kriging(mdl, pts, unk, no_neighbors=32)
        t2 = time.perf_counter()
        print(f"chunk {chunk or 'default'} rep {rep}: _lib.krige_host {1e3 * (t1 - t0):.1f} ms ({len(unk) / (t1 - t0):.3e} pts/s), "
              f"public call {1e3 * (t2 - t1):.1f} ms ({len(unk) / (t2 - t1):.3e} pts/s)", flush=True)
t0 = time.perf_counter(); a = np.empty((1, len(unk), 4)); a[:] = 0; t1 = time.perf_counter()
print(f"first touch of the 128 MB output: {1e3 * (t1 - t0):.1f} ms")
t0 = time.perf_counter(); b = unk.copy(); t1 = time.perf_counter()
print(f"host memcpy 64 MB: {1e3 * (t1 - t0):.1f} ms")
