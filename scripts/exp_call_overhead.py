#!/usr/bin/env python
"""Fixed cost of a host-array kriging call (1M known points): locations m = 1000 ... 500k, k = 32, library call and public call,
against the device-resident time of the same batch.  usage: python scripts/exp_call_overhead.py"""
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

known = dem_like_field(1_000_000, seed=4)
unk_all = uniform_locations(1_000_000, seed=5)
var = float(np.var(known[:, 2]))
mdl = {"variogram_model_type": "linear", "nugget": 0.0, "sill": var, "rang": 8000.0}
params = np.array([[0.0, var, 8000.0]])
d_known = torch.from_numpy(known).cuda()
for m in (1000, 100_000, 500_000):
    unk = unk_all[:m]
    d_unk = torch.from_numpy(unk).cuda()
    for rep in range(3):
        t0 = time.perf_counter()
        _lib.krige_host(known, unk, _lib.MODEL_IDS["linear"], params, _lib.KRIGE_ORDINARY, 32)
        t1 = time.perf_counter()
        pib.ordinary_kriging(mdl, known, unk, no_neighbors=32)
        t2 = time.perf_counter()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        _lib.krige_device(d_known, d_unk, _lib.MODEL_IDS["linear"], params, _lib.KRIGE_ORDINARY, 32)
        torch.cuda.synchronize()
        t4 = time.perf_counter()
        x = torch.from_numpy(known).cuda(); torch.cuda.synchronize()
        t5 = time.perf_counter()
    print(f"m={m}: _lib.krige_host {1e3 * (t1 - t0):.2f} ms, public call {1e3 * (t2 - t1):.2f} ms, device-resident call {1e3 * (t4 - t3):.2f} ms, "
          f"torch upload of the 24 MB known array {1e3 * (t5 - t4):.2f} ms", flush=True)
