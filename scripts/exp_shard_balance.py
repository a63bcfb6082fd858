#!/usr/bin/env python
"""Work-item granularity of the pair kernel for a sharded run, measured on ONE GPU: the kernel time of each of the 8 ranks'
slices of BASELINE config 2 for several PIB_PAIR_CHUNK values (the slowest rank sets the step time of the 8-GPU run).
usage: python scripts/exp_shard_balance.py [world]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from pyinterpolate_b200 import _lib  # noqa: E402
from pyinterpolate_b200.synthetic import dem_like_field  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
lags = np.arange(500.0, 40500.0, 500.0)
big = torch.from_numpy(dem_like_field(1_000_000, extent=100000.0, seed=2)).cuda()
_lib.variogram_device(big, lags)
torch.cuda.synchronize()
whole = []
for _ in range(3):
    _lib.variogram_device(big, lags); torch.cuda.synchronize()
    whole.append(_lib.last_kernel_ms(_lib.KERNEL_PAIR)[0])
print(f"whole job: {min(whole):.2f} ms -> ideal per rank {min(whole) / world:.2f} ms", flush=True)
for chunk in ("64", "32", "16", "8"):
    os.environ["PIB_PAIR_CHUNK"] = chunk
    ms = []
    for r in range(world):
        best = 1e9
        for _ in range(2):
            _lib.variogram_device(big, lags, shard=(r, world)); torch.cuda.synchronize()
            best = min(best, _lib.last_kernel_ms(_lib.KERNEL_PAIR)[0])
        ms.append(best)
    print(f"chunk {chunk:>2s}: per-rank kernel ms {['%.2f' % m for m in ms]}  max {max(ms):.2f}  mean {np.mean(ms):.2f}", flush=True)
