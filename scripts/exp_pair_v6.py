#!/usr/bin/env python
"""A/B of the pair-kernel variants on a B200: parity against the CPU oracle at 100k points, then the BASELINE
config 2 timing (1M points, 80 lags).  usage: python scripts/exp_pair_v6.py [settings ...]   (setting = "V=5" or "V6=18")"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402
from pyinterpolate_b200 import _lib  # noqa: E402
from pyinterpolate_b200.synthetic import dem_like_field  # noqa: E402

settings = sys.argv[1:] or ["V=5", "V6=1,8,512,16", "V6=1,16,512,16", "V6=2,8,256,16", "V6=2,8,384,12", "V6=2,8,384,10", "V6=1,8,384,16"]
lags = np.arange(500.0, 40500.0, 500.0)
small = dem_like_field(100_000, extent=31623.0, seed=11)      # same density as config 2
ref = None
if os.environ.get("EXP_ORACLE", "1") == "1":
    import oracle
    oracle.set_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    ref = oracle.variogram_sums(small, lags)
    print(f"oracle 100k: {time.perf_counter() - t0:.1f} s", flush=True)
big = torch.from_numpy(dem_like_field(1_000_000, extent=100000.0, seed=2)).cuda()
peak = _lib.fp64_peak_tflops()
print("lib", _lib.LIB_PATH, "fp64 peak", peak, flush=True)
for s in settings:
    for k in ("PIB_PAIR_V", "PIB_V6"):
        os.environ.pop(k, None)
    key, val = s.split("=")
    os.environ["PIB_PAIR_V" if key == "V" else "PIB_V6"] = val
    try:
        got = _lib.variogram_host(small, lags)
        msg = ""
        if ref is not None:
            ok_c = np.array_equal(got["count"], ref["count"])
            rel = np.max(np.abs(got["sum_sq"] - ref["sum_sq"]) / np.abs(ref["sum_sq"]))
            scale = np.abs(ref["count"] * (np.mean(small[:, 2]) ** 2 + np.var(small[:, 2])))
            relp = np.max(np.abs(got["sum_prod"] - ref["sum_prod"]) / scale)
            relv = np.max(np.abs(got["sum_val"] - ref["sum_val"]) / np.abs(ref["sum_val"]))
            msg = f"counts_exact={ok_c} sum_sq_rel={rel:.2e} sum_prod_rel={relp:.2e} sum_val_rel={relv:.2e}"
        ms = []
        for rep in range(4):
            r = _lib.variogram_device(big, lags, want_pairs=True)
            torch.cuda.synchronize()
            ms.append(_lib.last_kernel_ms(_lib.KERNEL_PAIR)[0])
        best = min(ms[1:])
        tf = 13 * r["pairs_evaluated"] / (best * 1e-3) / 1e12
        print(f"{s:8s} {msg}  1M: kernel {best:.2f} ms (all {['%.1f' % m for m in ms]}), evaluated {r['pairs_evaluated']:.4e}, "
              f"{tf:.2f} TFLOP/s = {tf / peak:.3f} of peak; count checksum {int(r['count'].sum().item())}", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"{s:8s} FAILED: {e}", flush=True)
