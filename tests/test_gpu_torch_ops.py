"""torch.ops.pyinterpolate_b200.* on CUDA tensors against the public calls (same library entry points underneath)."""
import numpy as np
import pytest

from pyinterpolate_b200.synthetic import dem_like_field, uniform_locations

pytestmark = pytest.mark.gpu


def test_torch_ops_match_the_public_calls():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pyinterpolate_b200 as pib
    from pyinterpolate_b200 import _lib, core, torch_ops
    torch_ops.register()
    pts = dem_like_field(20_000, seed=61)
    lags = np.arange(1000.0, 21000.0, 1000.0)
    xyz = torch.from_numpy(pts).cuda()
    sums, cnt = torch.ops.pyinterpolate_b200.variogram_sums(xyz, torch.from_numpy(lags), torch_ops.empty_f64(), torch_ops.empty_f64(), 0, 1)
    ev = pib.ExperimentalVariogram(pts, step_size=1000.0, max_range=21000.0)
    np.testing.assert_array_equal(cnt.cpu().numpy() * 2, ev.points_per_lag)
    np.testing.assert_allclose((sums[0] / cnt / 2.0).cpu().numpy(), ev.semivariances, rtol=1e-12)
    # four ellipse directions in one call
    W = np.stack([core.define_whitening_matrix(a, 0.3) for a in (90, 0, 45, 135)]).reshape(-1)
    sums4, cnt4 = torch.ops.pyinterpolate_b200.variogram_sums(xyz, torch.from_numpy(lags), torch.from_numpy(core.ellipse_lower_limits(lags)),
                                                              torch.from_numpy(W), 0, 1)
    dv = pib.ExperimentalVariogram(pts, step_size=1000.0, max_range=21000.0, direction=45, tolerance=0.3, dir_neighbors_selection_method="e")
    np.testing.assert_array_equal(cnt4[2].cpu().numpy() * 2, dv.points_per_lag)
    # kriging: ordinary + simple from one call
    unk = uniform_locations(3000, seed=62)
    var, mean = float(np.var(pts[:, 2])), float(pts[:, 2].mean())
    out, status = torch.ops.pyinterpolate_b200.krige(xyz, torch.from_numpy(unk).cuda(), torch.tensor([[1.0, var, 15000.0]], dtype=torch.float64),
                                                     torch_ops.model_id("spherical"), _lib.KRIGE_BOTH, mean, 16)
    assert int(status.sum()) == 0
    mdl = {"variogram_model_type": "spherical", "nugget": 1.0, "sill": var, "rang": 15000.0}
    ok = pib.ordinary_kriging(mdl, pts, unk, no_neighbors=16)
    sk = pib.simple_kriging(mdl, pts, unk, process_mean=mean, no_neighbors=16)
    np.testing.assert_array_equal(out[0, 0].cpu().numpy(), ok)
    np.testing.assert_array_equal(out[1, 0].cpu().numpy(), sk)
