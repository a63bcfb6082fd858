#!/usr/bin/env python
"""One device-resident kriging call (1M known, `m` locations) for profiling: python scripts/run_krige_once.py k mode [m]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from pyinterpolate_b200 import _lib  # noqa: E402
from pyinterpolate_b200.synthetic import dem_like_field, uniform_locations  # noqa: E402

k, mode = int(sys.argv[1]), int(sys.argv[2])
m = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000
pts = dem_like_field(1_000_000, seed=4)
d_pts = torch.from_numpy(pts).cuda()
unk = torch.from_numpy(uniform_locations(m, seed=5)).cuda()
var, mean = float(np.var(pts[:, 2])), float(np.mean(pts[:, 2]))
for _ in range(2):
    out, _, st = _lib.krige_device(d_pts, unk, _lib.MODEL_IDS["spherical"], np.array([[1.0, var, 20000.0]]), mode, k, process_mean=mean)
    torch.cuda.synchronize()
    print("knn ms", _lib.last_kernel_ms(_lib.KERNEL_KNN)[0], "solve ms", _lib.last_kernel_ms(_lib.KERNEL_SOLVE)[0], _lib.last_kernel_name(_lib.KERNEL_SOLVE))
