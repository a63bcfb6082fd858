"""CPU-side tests (-m "not gpu"): the C-ABI library loads and exports every declared symbol, the
host-side mirror of the reference interface validates like the reference, and the multi-rank plumbing
works over gloo with world_size 2.  No CUDA compute call is made here."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_library_exports_every_declared_symbol():
    import __graft_entry__ as entry
    from pyinterpolate_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        entry.build()
    header = open(os.path.join(ROOT, "include", "pyinterp_b200.h")).read()
    declared = set(re.findall(r"PIB_API\s+(?:const\s+)?\w+\s*\*?\s*(pib_\w+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.pib_version() >= 100
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in nm.splitlines() if " T " in line}
    assert declared <= exported
    # nothing but the C ABI leaks out of the library
    assert all(s.startswith("pib_") for s in exported if not s.startswith("_")), exported


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pyinterpolate_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower() or f == "synthetic.py", f


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from pyinterpolate_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.PibError, match="no CPU fallback"):
        _lib.variogram_host(np.zeros((4, 3)), np.array([1.0]))
    monkeypatch.undo()


def test_lags_and_validation_match_reference_rules():
    from pyinterpolate_b200 import core
    np.testing.assert_array_equal(core.get_lags(500, 40000, None), np.arange(500, 40000, 500))
    assert len(core.get_lags(500, 40000, None)) == 79 and len(core.get_lags(500, 40500, None)) == 80
    np.testing.assert_array_equal(core.get_lags(None, None, [0, 1, 2.5]), [1, 2.5])
    np.testing.assert_array_equal(core.get_lags(None, None, [1, 2.5]), [1, 2.5])
    with pytest.raises(ValueError):
        core.validate_bins(None, None, None)
    core.validate_direction_and_tolerance(None, None)
    core.validate_direction_and_tolerance(-360, 1)
    for bad in ((361, 0.5), (10, 0), (10, 1.5), (None, 0.5), (10, None)):
        with pytest.raises(ValueError):
            core.validate_direction_and_tolerance(*bad)
    with pytest.raises(IndexError):
        core.validate_semivariance_weights(np.zeros((3, 3)), [1, 2])
    with pytest.raises(ValueError):
        core.validate_semivariance_weights(np.zeros((2, 3)), [1, 0])
    assert core.interpolation_points([1.0, 2.0]).shape == (1, 2)
    assert core.interpolation_points([[1.0, 2.0], [3.0, 4.0]]).shape == (2, 2)
    assert core.variogram_points([[0, 0, 1], [1, 1, 2]]).shape == (2, 3)


def test_whitening_matrix_and_lower_limits_match_golden(golden):
    from pyinterpolate_b200 import core
    for (ang, tol), W in zip(golden["whiten_cfg"], golden["whiten_W"]):
        np.testing.assert_array_equal(core.define_whitening_matrix(ang, tol), W)
    for lags in (np.arange(500.0, 40500.0, 500.0), np.arange(0.1, 3.05, 0.1), np.arange(1.5, 6, 1.5)):
        lower = core.ellipse_lower_limits(lags)
        np.testing.assert_array_equal(lower, np.concatenate([[0.0], lags[:-1]]))   # regular bins tile exactly


def test_finalize_matches_oracle_conventions(oracle):
    from pyinterpolate_b200.variogram import _finalize
    sums = dict(sum_sq=np.array([0.0, 8.0, 0.0]), sum_prod=np.array([0.0, 6.0, 0.0]),
                sum_val=np.array([0.0, 10.0, 0.0]), count=np.array([0, 2, 0]))
    sem, cov, n = _finalize(sums, directional=False)
    o_sem, o_cov, o_n = oracle.finalize(sums)
    np.testing.assert_array_equal(sem, o_sem)
    np.testing.assert_array_equal(cov, o_cov)
    np.testing.assert_array_equal(n, o_n)
    assert sem[0] == 0 and np.isnan(sem[2]) and n[1] == 4


def test_model_container():
    import pyinterpolate_b200 as pib
    m = pib.TheoreticalVariogram({"variogram_model_type": "spherical", "nugget": 1, "sill": 2, "rang": 3})
    assert m.to_dict() == dict(nugget=1, sill=2, rang=3, variogram_model_type="spherical", direction=None)
    assert m.name == "spherical"
    with pytest.raises(KeyError):
        pib.TheoreticalVariogram({"variogram_model_type": "matern", "nugget": 1, "sill": 2, "rang": 3})
    from pyinterpolate_b200.models import model_parameters

    class Duck:
        nugget, sill, rang, model_type, direction = 0.5, 2.0, 10.0, "cubic", None
    assert model_parameters(Duck()) == (1, 0.5, 2.0, 10.0, None)


def test_shard_bounds_cover_everything():
    from pyinterpolate_b200.distributed import shard_bounds
    for m in (0, 1, 7, 10_000_001):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(m, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == m
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1


_WORKER = r'''
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "oracle"))
import numpy as np, torch.distributed as dist
import oracle
from pyinterpolate_b200 import distributed as pd
from pyinterpolate_b200.synthetic import dem_like_field
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank, world = pd.rank_world()
pts = dem_like_field(1500, seed=3)
lags = np.arange(2500.0, 40000.0, 2500.0)
# each rank contributes the pair rows it owns (the GPU path shards tile pairs; the reduction is the same)
b, e = pd.shard_bounds(len(pts), rank, world)
part = oracle.variogram_sums(pts, lags, rows=(b, e))
part["pairs_evaluated"] = 10 + rank
total = pd.allreduce_variogram_sums(part)
full = oracle.variogram_sums(pts, lags)
assert np.array_equal(total["count"], full["count"])
assert np.allclose(total["sum_sq"], full["sum_sq"], rtol=1e-12)
assert total["pairs_evaluated"] == 21
rows = np.arange(2 * 11, dtype=np.float64).reshape(11, 2)
b, e = pd.shard_bounds(11, rank, world)
back = pd.gather_rows(rows[b:e] * 1.0, 11)
assert np.array_equal(back, rows)
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_rank_gloo_allreduce_and_gather(tmp_path, oracle):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"rank {r} ok" in o


def test_triangle_lag_table_matches_reference_vertices():
    """The nine doubles per lag handed to the triangle kernel are the reference's own vertex offsets
    (distance/angular.py:283-339 generate_triangles): apex = p + (h cos a, h sin a), base = p +- (h tol)(cos, sin)(a +- 90)."""
    from pyinterpolate_b200.variogram import _triangle_lag_table
    lags = np.arange(500.0, 3000.0, 500.0)
    table, c, s = _triangle_lag_table(lags, 30.0, 0.2)
    assert table.shape == (5, 9)
    a = np.radians(30.0)
    np.testing.assert_array_equal(table[:, 0], lags)
    np.testing.assert_array_equal(table[:, 1], lags * np.cos(float(a)))
    np.testing.assert_array_equal(table[:, 3], -lags * np.cos(float(a)))
    np.testing.assert_array_equal(table[:, 5], (lags * 0.2) * np.cos(a + np.pi / 2))
    np.testing.assert_array_equal(table[:, 8], (lags * 0.2) * np.sin(a - np.pi / 2))
    assert c == float(np.cos(a)) and s == float(np.sin(a))
    # integer lags (np.arange of ints) give the same doubles
    t_int, _, _ = _triangle_lag_table(np.arange(1, 4, 1), 15, 0.25)
    t_flt, _, _ = _triangle_lag_table(np.arange(1.0, 4.0, 1.0), 15, 0.25)
    np.testing.assert_array_equal(t_int, t_flt)


def test_indicator_coding_and_thresholds():
    from pyinterpolate_b200.indicator import IndicatorVariogramData, code_indicators, select_variogram_thresholds
    ds = np.column_stack([np.arange(10.0), np.arange(10.0), np.arange(10.0)])
    thr = select_variogram_thresholds(ds[:, 2], 4)
    assert thr == [np.quantile(ds[:, 2], q) for q in (0.25, 0.5, 0.75, 1.0)]
    ids = code_indicators(ds, thr)
    assert ids.shape == (10, 6) and ids[:, 2:].sum(axis=0).tolist() == [3, 5, 7, 10]
    assert IndicatorVariogramData(ds, 4).ids.shape == (10, 6)


def test_variogram_cloud_remove_outliers_matches_reference():
    """Host-side outlier filters of VariogramCloud (classes/variogram_cloud.py:334-397) against fixtures from the
    reference's remove_outliers (tests/golden/make_golden_pipelines.py); no GPU involved."""
    from pyinterpolate_b200.variogram import VariogramCloud
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_golden_pipelines.npz"))
    vc = VariogramCloud.__new__(VariogramCloud)
    vc.semivariances = {lag: g[f"cloud_in_{int(lag)}"] for lag in (500.0, 1000.0, 1500.0)}
    vc.lags = np.array([500.0, 1000.0, 1500.0])
    cases = (("z", dict(method="zscore")), ("z2", dict(method="zscore", z_lower_limit=-1.5, z_upper_limit=2.0)),
             ("iqr", dict(method="iqr")), ("iqr2", dict(method="iqr", iqr_lower_limit=0.5, iqr_upper_limit=1.0)))
    for name, kw in cases:
        out = vc.remove_outliers(**kw)
        for lag in vc.lags:
            np.testing.assert_array_equal(out.semivariances[lag], g[f"cloud_{name}_{int(lag)}"])
        assert out is not vc and len(vc.semivariances[1000.0]) == 302
    with pytest.raises(KeyError):
        vc.remove_outliers(method="mad")
    with pytest.raises(ValueError):
        vc.remove_outliers(z_lower_limit=1)
    assert vc.remove_outliers(method="iqr", inplace=True) is None
    np.testing.assert_array_equal(vc.semivariances[1000.0], g["cloud_iqr_1000"])


def test_indicator_expected_values_match_per_location_splines():
    """IndicatorKriging._get_expected_values evaluates ONE spline basis for all locations (a matrix product); the
    reference fits a spline per location (kriging/point/indicator.py:327-380).  Same numbers, row by row."""
    from scipy.interpolate import UnivariateSpline
    from pyinterpolate_b200.indicator import IndicatorKriging
    rng = np.random.default_rng(3)
    for n_thr in (3, 4, 8):
        thr = np.sort(rng.uniform(150.0, 250.0, n_thr))
        preds = np.sort(rng.uniform(0.0, 1.0, (200, n_thr)), axis=1)          # indicator probabilities grow with the threshold
        preds[:5] = rng.uniform(0.0, 1.0, (5, n_thr))                            # and a few non-monotone rows
        obj = IndicatorKriging.__new__(IndicatorKriging)
        obj.thresholds, obj.indicator_predictions = thr, preds
        got_e, got_v = obj._get_expected_values()
        grid = np.linspace(thr.min(), thr.max(), 100)
        for r in range(len(preds)):
            spl = UnivariateSpline(thr, preds[r], k=3 if n_thr >= 4 else n_thr - 1, s=0)
            vals = spl(grid)
            ccdf = 1 - vals.cumsum() / vals.sum()
            idx = len(ccdf) - np.searchsorted(np.flip(ccdf), np.mean(ccdf))
            want_e = grid[min(idx, len(grid) - 1)]
            want_v = np.mean(np.abs(vals - want_e) ** 2)
            # an exact tie of the ccdf with its mean can move the index by one between the two evaluation orders
            assert abs(got_e[r] - want_e) <= (grid[1] - grid[0]) * 1.0000001
            if got_e[r] == want_e:
                np.testing.assert_allclose(got_v[r], want_v, rtol=1e-9)


def test_directional_variogram_passes_custom_weights(monkeypatch):
    """DirectionalVariogram hands custom_weights to all five variograms (classes/directional_variogram.py:126-143)."""
    import pyinterpolate_b200.variogram as vg
    seen = []

    class Fake:
        def __init__(self, *a, **k):
            seen.append(k.get("custom_weights"))
    monkeypatch.setattr(vg, "ExperimentalVariogram", Fake)
    w = np.ones(10)
    vg.DirectionalVariogram(np.zeros((10, 3)), 1.0, 4.0, custom_weights=w, dir_neighbors_selection_method="e")
    assert len(seen) == 5 and all(s is w for s in seen)


def test_torch_ops_are_registered_and_have_no_cpu_kernel():
    """torch.ops.pyinterpolate_b200.* (pyinterpolate_b200/torch_ops.py): schemas, output shapes of the fake kernels, and no CPU
    implementation behind them."""
    import torch
    from torch._subclasses.fake_tensor import FakeTensorMode
    from pyinterpolate_b200 import _lib, torch_ops
    torch_ops.register()
    torch_ops.register()
    with pytest.raises(NotImplementedError):
        torch.ops.pyinterpolate_b200.variogram_sums(torch.zeros(4, 3, dtype=torch.float64), torch.tensor([1.0, 2.0], dtype=torch.float64),
                                                    torch_ops.empty_f64(), torch_ops.empty_f64(), 0, 1)
    with FakeTensorMode():
        x = torch.empty(10, 3, dtype=torch.float64, device="cuda")
        s, c = torch.ops.pyinterpolate_b200.variogram_sums(x, torch.empty(5, dtype=torch.float64), torch.empty(0, dtype=torch.float64),
                                                           torch.empty(8, dtype=torch.float64), 0, 1)
        assert tuple(s.shape) == (3, 2, 5) and tuple(c.shape) == (2, 5) and c.dtype == torch.int64
        o, st = torch.ops.pyinterpolate_b200.krige(x, torch.empty(7, 2, dtype=torch.float64, device="cuda"),
                                                   torch.empty(1, 3, dtype=torch.float64), 0, _lib.KRIGE_BOTH, 0.0, 4)
        assert tuple(o.shape) == (2, 1, 7, 4) and tuple(st.shape) == (2, 1, 7) and st.dtype == torch.uint8


def test_kriging_inputs_with_non_finite_coordinates_are_refused():
    """The Python layer rejects NaN / inf coordinates before anything reaches the library (a NaN location would select no
    neighbours); a non-finite VALUE is not a coordinate error."""
    from pyinterpolate_b200.kriging import _prepare
    mdl = {"variogram_model_type": "linear", "nugget": 0.0, "sill": 1.0, "rang": 10.0}
    known = np.random.default_rng(3).uniform(0, 100, (50, 3))
    unk = np.random.default_rng(4).uniform(0, 100, (7, 2))
    _prepare(mdl, known, unk, False, False)
    for bad in (np.nan, np.inf, -np.inf):
        k2 = known.copy(); k2[17, 1] = bad
        with pytest.raises(ValueError):
            _prepare(mdl, k2, unk, False, False)
        u2 = unk.copy(); u2[3, 0] = bad
        with pytest.raises(ValueError):
            _prepare(mdl, known, u2, False, False)
    k3 = known.copy(); k3[5, 2] = np.nan                    # a missing value is the caller's business, like in the reference
    _prepare(mdl, k3, unk, False, False)
    k4 = known.copy(); k4[:, 2] = 1e308                      # values whose sum overflows: exact scan of the coordinates decides
    _prepare(mdl, k4, unk, False, False)


def test_timing_record_switch_needs_no_device():
    """pib_timing_accumulate only manages the host-side records: callable (and harmless) without a GPU — bench.py brackets its
    timed region with it."""
    from pyinterpolate_b200 import _lib
    _lib.timing_accumulate(True)
    _lib.timing_accumulate(False)
