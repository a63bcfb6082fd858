#!/usr/bin/env python
"""A/B of the two Gauss-Jordan kernels for k <= 32 on a B200: solve_gj2_kernel (two rows per thread, several systems per
warp; default) against solve_gj_kernel (PIB_GJ2=0) — device-resident timing (1M known, 1M locations) per k and mode, and the
agreement of the outputs (identical pivot sequence: only the final reduction order differs).
usage: python scripts/exp_gj2.py [m]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from pyinterpolate_b200 import _lib  # noqa: E402
from pyinterpolate_b200.synthetic import dem_like_field, uniform_locations  # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
pts = dem_like_field(1_000_000, seed=4)
d_pts = torch.from_numpy(pts).cuda()
unk = torch.from_numpy(uniform_locations(m, seed=5)).cuda()
var, mean = float(np.var(pts[:, 2])), float(np.mean(pts[:, 2]))
for k, model, params in ((32, "linear", [0.0, var, 8000.0]), (32, "spherical", [1.0, var, 20000.0]), (24, "gaussian", [0.1, var, 3000.0]),
                         (16, "exponential", [0.5, var, 9000.0]), (8, "cubic", [0.0, var, 9000.0]), (5, "circular", [0.0, var, 5000.0])):
    res = {}
    for ver in ("0", "1"):
        os.environ["PIB_GJ2"] = ver
        for mode in (0, 1, 2):
            def call():
                return _lib.krige_device(d_pts, unk, _lib.MODEL_IDS[model], np.array([params]), mode, k, process_mean=mean)
            out, _, st = call()
            torch.cuda.synchronize()
            ms = []
            for _ in range(3):
                call(); torch.cuda.synchronize()
                ms.append(_lib.last_kernel_ms(_lib.KERNEL_SOLVE)[0])
            res[(ver, mode)] = (out.cpu().numpy(), st.cpu().numpy())
            print(f"k={k} {model} {_lib.last_kernel_name(_lib.KERNEL_SOLVE)} mode {mode}: solve {min(ms):.2f} ms, knn {_lib.last_kernel_ms(_lib.KERNEL_KNN)[0]:.2f} ms, "
                  f"status counts {np.bincount(st.cpu().numpy().ravel(), minlength=6)[:6].tolist()}", flush=True)
    for mode in (0, 1, 2):
        (a, sa), (b, sb) = res[("0", mode)], res[("1", mode)]
        a, b = a.reshape(-1, 4), b.reshape(-1, 4)
        ok = np.isfinite(a[:, 1]) & np.isfinite(b[:, 1])
        print(f"   mode {mode}: max rel diff zhat {np.max(np.abs(a[ok, 0] - b[ok, 0]) / np.abs(a[ok, 0])):.2e}, "
              f"sigma2 {np.max(np.abs(a[ok, 1] - b[ok, 1]) / np.maximum(np.abs(a[ok, 1]), 1e-300)):.2e}, nan mismatch "
              f"{int((np.isnan(a[:, 1]) != np.isnan(b[:, 1])).sum())}, status equal {np.array_equal(sa, sb)}, xy equal {np.array_equal(a[:, 2:], b[:, 2:])}", flush=True)
