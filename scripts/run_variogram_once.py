#!/usr/bin/env python
"""One device-resident variogram pass of BASELINE config 2 (1M points, 80 lags) for profiling / timing:
python scripts/run_variogram_once.py [repeats]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from pyinterpolate_b200 import _lib  # noqa: E402
from pyinterpolate_b200.synthetic import dem_like_field  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
pts = dem_like_field(1_000_000, seed=2)
d = torch.from_numpy(pts).cuda()
lags = np.arange(500.0, 40500.0, 500.0)
for _ in range(reps):
    r = _lib.variogram_device(d, lags, want_pairs=True)
    torch.cuda.synchronize()
    print("pair kernel ms", _lib.last_kernel_ms(_lib.KERNEL_PAIR)[0], _lib.last_kernel_name(_lib.KERNEL_PAIR),
          "evaluated", r["pairs_evaluated"], "count checksum", int(r["count"].sum()))
