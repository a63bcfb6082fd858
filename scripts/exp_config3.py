#!/usr/bin/env python
"""BASELINE config 3 timing (200k points, 4 ellipse directions, tolerance 0.2, 80 lags) for the kernel variants, with the
counts checked against each other.  usage: python scripts/exp_config3.py [settings...]  (KEY=VALUE;KEY=VALUE... per setting)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from pyinterpolate_b200 import _lib, core  # noqa: E402
from pyinterpolate_b200.synthetic import dem_like_field  # noqa: E402

settings = sys.argv[1:] or ["PIB_PAIR_V=5", "", "PIB_FORCE_RING=1;PIB_RING=32", "PIB_FORCE_RING=1;PIB_RING=32;PIB_GROUP_PATH=1", "PIB_FORCE_FULL=1"]
lags = np.arange(500.0, 40500.0, 500.0)
p3 = torch.from_numpy(dem_like_field(200_000, extent=100_000.0, seed=3)).cuda()
W = np.stack([core.define_whitening_matrix(a, 0.2) for a in (90, 0, 45, 135)]).reshape(-1, 4)
lower = core.ellipse_lower_limits(lags)
peak = _lib.fp64_peak_tflops()
ref = None
for s in settings:
    for k in ("PIB_PAIR_V", "PIB_FORCE_RING", "PIB_RING", "PIB_GROUP_PATH", "PIB_FORCE_FULL", "PIB_V6", "PIB_LAGS_PER_GROUP"):
        os.environ.pop(k, None)
    for kv in filter(None, s.split(";")):
        k, v = kv.split("=", 1)
        os.environ[k] = v
    ms = []
    for _ in range(4):
        r = _lib.variogram_device(p3, lags, lower=lower, W=W, want_pairs=True)
        torch.cuda.synchronize()
        ms.append(_lib.last_kernel_ms(_lib.KERNEL_PAIR)[0])
    cnt = r["count"].cpu().numpy()
    if ref is None:
        ref = cnt
    tf = 19 * r["pairs_evaluated"] / (min(ms[1:]) * 1e-3) / 1e12
    print(f"{s or 'default':48s} pair kernels {min(ms[1:]):7.2f} ms for 4 directions ({_lib.last_kernel_name(_lib.KERNEL_PAIR)}), evaluated "
          f"{r['pairs_evaluated']:.3e}, {tf:.2f} TFLOP/s = {tf / peak:.3f} of peak (19 FLOP/pair), counts equal: {np.array_equal(cnt, ref)}", flush=True)
