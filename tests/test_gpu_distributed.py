"""Two ranks over NCCL (one process per GPU): the public sharded calls must reproduce the single-GPU results.

  * experimental_variogram_sharded — every rank evaluates its slice of the tile-pair list, ONE all-reduce of the per-lag
    sums: pair counts bit-exact, sums to 1e-12
  * ordinary_kriging_sharded / ordinary_and_simple_kriging — locations split over the ranks, no data-path collective:
    rows bit-identical to the single-GPU call (section 8e row 2)
Needs two visible GPUs (`gpurun --gpus 2`); skipped otherwise.  The gloo twin of the collective logic runs in the CPU suite
(tests/test_host_cpu.py)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import pyinterpolate_b200 as pib
from pyinterpolate_b200 import distributed as pd
from pyinterpolate_b200.synthetic import dem_like_field, uniform_locations
pts = dem_like_field(60_000, extent=40_000.0, seed=21)
# variogram: host arrays in, numpy out — the collective buffers are staged on the GPU for NCCL
lags, sem, cov, npairs = pd.experimental_variogram_sharded(pts, step_size=500, max_range=12000)
d_lags, d_sem, d_cov, d_np = pd.experimental_variogram_sharded(pts, step_size=500, max_range=12000, directions=[0, 45],
                                                                tolerance=0.3, dir_neighbors_selection_method="e")
unk = uniform_locations(30_001, extent=40_000.0, seed=22)
mdl = {{"variogram_model_type": "spherical", "nugget": 1.0, "sill": float(np.var(pts[:, 2])), "rang": 9000.0}}
ok = pd.ordinary_kriging_sharded(mdl, pts, unk, no_neighbors=16)
if rank == 0:
    ev = pib.ExperimentalVariogram(pts, step_size=500, max_range=12000)
    assert np.array_equal(npairs, ev.points_per_lag), "sharded pair counts differ"
    np.testing.assert_allclose(sem, ev.semivariances, rtol=1e-12)
    np.testing.assert_allclose(cov, ev.covariances, rtol=1e-9, atol=1e-9 * float(np.mean(pts[:, 2]) ** 2))
    for t, a in enumerate([0, 45]):
        one = pib.ExperimentalVariogram(pts, step_size=500, max_range=12000, direction=a, tolerance=0.3,
                                        dir_neighbors_selection_method="e")
        assert np.array_equal(d_np[t], one.points_per_lag)
        np.testing.assert_allclose(d_sem[t], one.semivariances, rtol=1e-12)
    single = pib.ordinary_kriging(mdl, pts, unk, no_neighbors=16)
    assert np.array_equal(ok, single, equal_nan=True), "sharded kriging rows differ"
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_rank_nccl_sharded_calls_match_single_gpu(tmp_path):
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-4000:]
    assert "rank 0 ok" in out.stdout and "rank 1 ok" in out.stdout
