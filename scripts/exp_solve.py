#!/usr/bin/env python
"""A/B of the kriging solvers on a B200: device-resident timing (1M known, 1M locations) per mode and k,
old bordered LU (PIB_SOLVE_V=1) vs Gauss-Jordan (default), plus agreement between the two."""
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
peak = _lib.fp64_peak_tflops()
for k, model, params in ((32, "linear", [0.0, var, 8000.0]), (64, "spherical", [1.0, var, 20000.0]), (16, "exponential", [0.5, var, 9000.0])):
    res = {}
    for ver in ("1", "2"):
        os.environ["PIB_SOLVE_V"] = ver
        for mode in (0, 1, 2):
            if ver == "1" and mode == 2:
                continue
            def call():
                return _lib.krige_device(d_pts, unk, _lib.MODEL_IDS[model], np.array([params]), mode, k, process_mean=mean)
            out, _, st = call()
            torch.cuda.synchronize()
            ms = []
            for _ in range(3):
                call(); torch.cuda.synchronize()
                ms.append(_lib.last_kernel_ms(_lib.KERNEL_SOLVE)[0])
            res[(ver, mode)] = out.cpu().numpy()
            dim = k + 1 if mode == 0 else k
            flop = (2 * dim ** 3 / 3 + 2 * dim ** 2) * 1e6 * (2 if mode == 2 else 1)
            print(f"k={k} {model} solver v{ver} mode {mode}: solve {min(ms):.2f} ms  ({flop / (min(ms) * 1e-3) / 1e12 / peak:.3f} of fp64 peak), "
                  f"status counts {np.bincount(st.cpu().numpy().ravel(), minlength=6)[:6].tolist()}", flush=True)
    for mode in (0, 1):
        a, b = res[("1", mode)][0], res[("2", mode)][0]
        both = res[("2", 2)][mode if False else 0] if False else None
        ok = np.isfinite(a[:, 1]) & np.isfinite(b[:, 1])
        print(f"   mode {mode}: max rel diff zhat {np.max(np.abs(a[ok, 0] - b[ok, 0]) / np.abs(a[ok, 0])):.2e}, "
              f"sigma2 {np.max(np.abs(a[ok, 1] - b[ok, 1]) / np.maximum(np.abs(a[ok, 1]), 1e-300)):.2e}, nan mismatch {int((np.isnan(a[:, 1]) != np.isnan(b[:, 1])).sum())}")
    bo = res[("2", 2)]
    print(f"   BOTH == single modes bit for bit: ok {np.array_equal(bo[0], res[('2', 0)][0], equal_nan=True)}, sk {np.array_equal(bo[1], res[('2', 1)][0], equal_nan=True)}", flush=True)
