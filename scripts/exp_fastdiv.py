#!/usr/bin/env python
"""A/B of the range quotients in the Gauss-Jordan kernels: div_by (Markstein correction, default) against plain fp64 divisions
(PIB_FAST_DIV=0) — outputs must be bit-identical; solve times per k and model.  usage: python scripts/exp_fastdiv.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from pyinterpolate_b200 import _lib  # noqa: E402
from pyinterpolate_b200.synthetic import dem_like_field, uniform_locations  # noqa: E402

pts = dem_like_field(1_000_000, seed=4)
d_pts = torch.from_numpy(pts).cuda()
unk = torch.from_numpy(uniform_locations(1_000_000, seed=5)).cuda()
var, mean = float(np.var(pts[:, 2])), float(np.mean(pts[:, 2]))
for k, model, params in ((32, "linear", [0.0, var, 8000.0]), (32, "spherical", [1.0, var, 20000.0]), (64, "spherical", [1.0, var, 20000.0]),
                         (32, "gaussian", [0.1, var, 3000.0]), (16, "exponential", [0.5, var, 9123.456]), (16, "cubic", [0.0, var, 9000.0]),
                         (16, "circular", [0.0, var, 5000.7]), (16, "power", [0.0, var, 7777.7])):
    res = {}
    for fd in ("0", "1"):
        os.environ["PIB_FAST_DIV"] = fd
        for mode in (2,):
            def call():
                return _lib.krige_device(d_pts, unk, _lib.MODEL_IDS[model], np.array([params]), mode, k, process_mean=mean)
            out, _, st = call()
            torch.cuda.synchronize()
            ms = []
            for _ in range(3):
                call(); torch.cuda.synchronize()
                ms.append(_lib.last_kernel_ms(_lib.KERNEL_SOLVE)[0])
            res[fd] = out.cpu().numpy()
            print(f"k={k} {model} PIB_FAST_DIV={fd} {_lib.last_kernel_name(_lib.KERNEL_SOLVE)} OK+SK: solve {min(ms):.2f} ms", flush=True)
    print(f"   bit-identical outputs: {np.array_equal(res['0'], res['1'], equal_nan=True)}", flush=True)
