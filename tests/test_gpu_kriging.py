"""GPU parity tests for ordinary / simple point kriging (run with -m gpu on the B200 box).

Bar: neighbour indices bit-exact (tie-free inputs; on gridded data only locations whose k-th and
(k+1)-th distances differ are compared); predictions and kriging variances within 1e-9 relative.
"""
import numpy as np
import pytest

from pyinterpolate_b200.synthetic import dem_like_field, uniform_locations

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.fixture(scope="module")
def pib():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pyinterpolate_b200 as pib
    return pib


def _model(pib, name, nugget, sill, rang):
    return pib.TheoreticalVariogram({"variogram_model_type": name, "nugget": nugget, "sill": sill, "rang": rang})


def test_gamma_models_golden(pib, golden):
    h = golden["gamma_h"]
    for mi, name in enumerate(golden["gamma_models"]):
        for ti, (nug, sill, rang) in enumerate(golden["gamma_params"]):
            got = _model(pib, str(name), nug, sill, rang).predict(h)
            np.testing.assert_allclose(got, golden["gamma_values"][mi, ti], rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("name", ["linear", "spherical", "exponential", "circular", "cubic"])
def test_kriging_10k_golden(pib, golden, name):
    """BASELINE config 1 shape: 10k known points, k=32."""
    known = dem_like_field(10000, seed=0)
    unk = uniform_locations(200, seed=1)
    var = float(golden["krige10k_var"])
    mdl = _model(pib, name, 0.0 if name == "linear" else 1.0, var, 8000.0)
    ok, idx = pib.ordinary_kriging(mdl, known, unk, no_neighbors=32, return_neighbors=True)
    sk, idx2 = pib.simple_kriging(mdl, known, unk, process_mean=float(known[:, -1].mean()), no_neighbors=32,
                                  return_neighbors=True)
    np.testing.assert_array_equal(idx, golden["krige10k_nbr_idx"])
    np.testing.assert_array_equal(idx2, idx)
    np.testing.assert_allclose(ok[:, :2], golden[f"krige10k_{name}_ok"][:, :2], rtol=RTOL)
    np.testing.assert_allclose(sk[:, :2], golden[f"krige10k_{name}_sk"][:, :2], rtol=RTOL)
    np.testing.assert_array_equal(ok[:, 2:], golden[f"krige10k_{name}_ok"][:, 2:])


def test_kriging_200k_k64_golden(pib, golden):
    known = dem_like_field(200_000, extent=44_721.36, seed=4)
    unk = uniform_locations(40, extent=44_721.36, seed=5)
    mdl = _model(pib, "spherical", 1.0, float(golden["krige200k_var"]), 20000.0)
    ok, idx = pib.ordinary_kriging(mdl, known, unk, no_neighbors=64, return_neighbors=True)
    sk = pib.simple_kriging(mdl, known, unk, process_mean=float(known[:, -1].mean()), no_neighbors=64)
    np.testing.assert_array_equal(idx, golden["krige200k_nbr_idx"])
    np.testing.assert_allclose(ok[:, :2], golden["krige200k_ok"][:, :2], rtol=RTOL)
    np.testing.assert_allclose(sk[:, :2], golden["krige200k_sk"][:, :2], rtol=RTOL)


def test_armstrong_golden(pib, golden):
    arm = golden["armstrong_points"].astype(float)
    unk = golden["armstrong_krige_unknown"]
    sill = float(golden["armstrong_krige_sill"])
    mdl = _model(pib, "spherical", 0.0, sill, 4.0)
    ok = pib.ordinary_kriging(mdl, arm, unk, no_neighbors=8)
    sk = pib.simple_kriging(mdl, arm, unk, process_mean=float(arm[:, -1].mean()), no_neighbors=8)
    compared = 0
    for u in range(len(unk)):
        d = np.sort(np.hypot(arm[:, 0] - unk[u, 0], arm[:, 1] - unk[u, 1]))
        if d[7] == d[8]:
            continue                                   # neighbour set is tie-dependent in the reference itself
        np.testing.assert_allclose(ok[u, :2], golden["armstrong_ok"][u, :2], rtol=RTOL)
        np.testing.assert_allclose(sk[u, :2], golden["armstrong_sk"][u, :2], rtol=RTOL)
        compared += 1
    assert compared >= 1
    # the reference's README spelling of the keyword
    one = pib.ordinary_kriging(mdl, arm, unknown_location=unk[0], no_neighbors=8)
    assert one.shape == (1, 4)
    ok12 = pib.ordinary_kriging(_model(pib, "exponential", 0.0, 9.0, 3.0), arm, unk[:4], no_neighbors=12)
    for u in range(4):
        d = np.sort(np.hypot(arm[:, 0] - unk[u, 0], arm[:, 1] - unk[u, 1]))
        if d[11] != d[12]:
            np.testing.assert_allclose(ok12[u, :2], golden["armstrong_exp_ok"][u, :2], rtol=RTOL)


def test_edge_cases_golden(pib, golden):
    few = dem_like_field(5, extent=100.0, seed=31)
    unk = np.array([[50.0, 50.0], [10.0, 80.0]])
    mdl = _model(pib, "spherical", 0.5, 4.0, 60.0)
    ok, idx = pib.ordinary_kriging(mdl, few, unk, no_neighbors=8, return_neighbors=True)
    sk = pib.simple_kriging(mdl, few, unk, process_mean=200.0, no_neighbors=8)
    np.testing.assert_allclose(ok[:, :2], golden["few_ok"][:, :2], rtol=RTOL)      # fewer known points than k
    np.testing.assert_allclose(sk[:, :2], golden["few_sk"][:, :2], rtol=RTOL)
    assert (idx[:, 5:] == -1).all() and (idx[:, :5] >= 0).all()
    known = dem_like_field(10000, seed=0)
    unk = uniform_locations(200, seed=1)[:50]
    ok = pib.ordinary_kriging(_model(pib, "spherical", 1.0, float(golden["krige10k_var"]), 500.0), known, unk,
                              no_neighbors=16)
    np.testing.assert_allclose(ok[:, :2], golden["krige10k_shortrange_ok"][:, :2], rtol=RTOL)
    # duplicated coordinates: RuntimeError like the reference
    arm = golden["armstrong_points"].astype(float)
    dup = np.vstack([arm, arm[:3]])
    with pytest.raises(RuntimeError, match="Singular matrix"):
        pib.ordinary_kriging(_model(pib, "spherical", 0.0, float(golden["armstrong_krige_sill"]), 4.0), dup,
                             np.array([[0.4, 0.3]]), no_neighbors=8)
    with pytest.raises(KeyError):
        pib.ordinary_kriging({"variogram_model_type": "matern", "nugget": 0, "sill": 1, "rang": 1}, arm, unk)


def test_exact_interpolation_at_known_points(pib):
    """Kriging at a known location with nugget 0 returns that value and zero variance (reference probe)."""
    known = dem_like_field(5000, seed=8)
    mdl = _model(pib, "spherical", 0.0, float(np.var(known[:, 2])), 9000.0)
    ok = pib.ordinary_kriging(mdl, known, known[:300, :2], no_neighbors=16)
    np.testing.assert_allclose(ok[:, 0], known[:300, 2], rtol=1e-9)
    assert np.nanmax(np.abs(ok[:, 1])) < 1e-6 * float(np.var(known[:, 2]))


@pytest.mark.parametrize("k,model", [(64, "spherical"), (32, "exponential"), (4, "linear"), (100, "spherical")])
def test_vs_oracle_1m_known(pib, oracle, k, model):
    """BASELINE config 4 shape: 1M known points; sampled unknown locations against the full known set."""
    known = dem_like_field(1_000_000, seed=4)
    unk = uniform_locations(300, seed=5)
    var = float(np.var(known[:, 2]))
    nug = 0.0 if model == "linear" else 1.0
    mdl = _model(pib, model, nug, var, 20000.0)
    ok, idx = pib.ordinary_kriging(mdl, known, unk, no_neighbors=k, return_neighbors=True)
    sk = pib.simple_kriging(mdl, known, unk, process_mean=float(known[:, 2].mean()), no_neighbors=k)
    ref_ok, ref_idx, _ = oracle.krige(known, unk, model, nug, var, 20000.0, k, "ok")
    ref_sk, _, _ = oracle.krige(known, unk, model, nug, var, 20000.0, k, "sk", process_mean=float(known[:, 2].mean()))
    np.testing.assert_array_equal(idx, ref_idx)
    np.testing.assert_allclose(ok[:, :2], ref_ok[:, :2], rtol=RTOL)
    np.testing.assert_allclose(sk[:, :2], ref_sk[:, :2], rtol=RTOL)


@pytest.mark.parametrize("k,model,nug,rang", [(5, "circular", 0.0, 15000.0), (8, "cubic", 0.5, 15000.0), (12, "gaussian", 1.0, 1500.0),
                                              (16, "exponential", 0.0, 15000.0), (16, "spherical", 1.0, 15000.0),
                                              (24, "linear", 0.0, 15000.0), (32, "spherical", 1.0, 15000.0)])
def test_two_row_solver_vs_oracle_and_one_row_solver(pib, oracle, monkeypatch, k, model, nug, rang):
    """solve_gj2_kernel (two rows per thread, 2 - 8 systems per warp; the default up to k = 16, PIB_GJ2=1 up to 32) against the
    oracle and against solve_gj_kernel (PIB_GJ2=0): same pivot sequence, so the outputs agree to the last reduction.  The
    known points hold DUPLICATED coordinates, so singular systems (every lane of theirs masked) share warps with regular ones;
    a location count that is not a multiple of the systems per CTA leaves masked systems at the end of the batch."""
    from pyinterpolate_b200 import _lib
    known = dem_like_field(20_000, seed=31)
    known[5000:5040, :2] = known[100:140, :2]                # 40 duplicated coordinates
    unk = np.concatenate([uniform_locations(1501, seed=32), known[100:140, :2] + 1.0])
    var, mean = float(np.var(known[:, 2])), float(known[:, 2].mean())
    params = np.array([[nug, var, rang]])
    res = {}
    for ver in ("1", "0"):
        monkeypatch.setenv("PIB_GJ2", ver)
        res[ver] = _lib.krige_host(known, unk, _lib.MODEL_IDS[model], params, _lib.KRIGE_BOTH, k, process_mean=mean)
        assert _lib.last_kernel_name(_lib.KERNEL_SOLVE) == ("solve_gj2_kernel" if ver == "1" else "solve_gj_kernel")
    (o1, _, s1), (o0, _, s0) = res["1"], res["0"]
    np.testing.assert_array_equal(s1, s0)
    assert (s1 == 1).any() and (s1 == 0).any()                # singular and regular systems side by side
    np.testing.assert_allclose(o1, o0, rtol=1e-12, equal_nan=True)
    for mi, mode in enumerate(("ok", "sk")):
        ref, _, rst = oracle.krige(known, unk, model, nug, var, rang, k, mode, process_mean=mean)
        fine = (rst == 0) & (s1[mi, 0] == 0)
        assert fine.sum() > 1400
        np.testing.assert_array_equal(s1[mi, 0] == 1, rst == 1)
        np.testing.assert_allclose(o1[mi, 0][fine, :2], ref[fine, :2], rtol=RTOL)


def test_column_entry_point_equals_the_full_rows(pib):
    """pib_krige_column_host (one column of the result rows travels back; what indicator kriging keeps) against the full
    [zhat, sigma^2, x, y] rows of pib_krige_ex_host: bit-identical, over several chunks, two value / parameter sets, both modes."""
    from pyinterpolate_b200 import _lib
    known = dem_like_field(30_000, seed=71)
    unk = uniform_locations(300_001, seed=72)                 # more than one chunk of the host pipeline, ragged last chunk
    var, mean = float(np.var(known[:, 2])), float(known[:, 2].mean())
    params = np.array([[1.0, var, 15000.0], [0.05, 0.25, 9000.0]])
    values = np.stack([known[:, 2], (known[:, 2] <= np.median(known[:, 2])).astype(np.float64)])
    mid = _lib.MODEL_IDS["spherical"]
    for mode in (_lib.KRIGE_ORDINARY, _lib.KRIGE_BOTH):
        full, _, st = _lib.krige_host(known, unk, mid, params, mode, 12, process_mean=mean, values=values)
        for col in (0, 1):
            got, st2 = _lib.krige_column_host(known, unk, mid, params, mode, 12, col, process_mean=mean, values=values)
            np.testing.assert_array_equal(got, full[..., col])
            np.testing.assert_array_equal(st2, st)
    with pytest.raises(_lib.PibError):
        _lib.krige_column_host(known, unk[:10], mid, params, _lib.KRIGE_ORDINARY, 12, 2, values=values)


def test_clustered_and_outside_locations(pib, oracle):
    """Strongly non-uniform density and locations far outside the data's bounding box."""
    rng = np.random.default_rng(3)
    centres = rng.uniform(0, 1000, size=(6, 2))
    xy = np.vstack([c + rng.normal(0, 4.0, size=(800, 2)) for c in centres] + [rng.uniform(0, 1000, size=(200, 2))])
    z = 10 + np.sin(xy[:, 0] / 90.0) + 0.1 * rng.normal(size=len(xy))
    known = np.column_stack([xy, z])
    unk = np.vstack([rng.uniform(-500, 1500, size=(200, 2)), centres, [[5000.0, -4000.0]]])
    mdl = _model(pib, "exponential", 0.01, float(np.var(z)), 120.0)
    ok, idx = pib.ordinary_kriging(mdl, known, unk, no_neighbors=24, return_neighbors=True)
    ref, ref_idx, _ = oracle.krige(known, unk, "exponential", 0.01, float(np.var(z)), 120.0, 24, "ok")
    np.testing.assert_array_equal(idx, ref_idx)
    np.testing.assert_allclose(ok[:, 0], ref[:, 0], rtol=RTOL)
    np.testing.assert_allclose(ok[:, 1], ref[:, 1], rtol=1e-7, atol=1e-9)    # far outside: variance saturates, ill-scaled


def test_indicator_kriging_batched_equals_per_threshold(pib):
    """BASELINE config 5 shape (scaled down): T thresholds share one neighbour search."""
    known = dem_like_field(20000, seed=6)
    unk = uniform_locations(500, seed=7)
    qs = np.quantile(known[:, 2], np.arange(1, 9) / 8.0)
    models = [_model(pib, "spherical", 0.01, max(p * (1 - p), 0.01), 20000.0) for p in np.arange(1, 9) / 8.0]
    got = pib.indicator_kriging(models, qs, known, unk, no_neighbors=32, reference_variance_column=False)
    var_col = pib.indicator_kriging(models, qs, known, unk, no_neighbors=32)
    assert got.shape == (500, 8) and var_col.shape == (500, 8)
    for t in (0, 3, 7):
        ind = known.copy(); ind[:, 2] = (ind[:, 2] <= qs[t]).astype(float)
        single = pib.ordinary_kriging(models[t], ind, unk, no_neighbors=32)
        np.testing.assert_allclose(got[:, t], np.clip(single[:, 0], 0, 1), rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(var_col[:, t], np.clip(np.nan_to_num(single[:, 1], nan=np.nan), 0, 1),
                                   rtol=1e-12, atol=1e-14, equal_nan=True)


def test_large_batch_properties(pib):
    """2M locations from 1M known points, k=64: splitting the batch changes nothing (locations are
    independent -> multi-GPU sharding is exact), device and host entry points agree."""
    import torch
    from pyinterpolate_b200 import _lib
    known = dem_like_field(1_000_000, seed=4)
    unk = uniform_locations(2_000_000, seed=5)
    params = np.array([[1.0, float(np.var(known[:, 2])), 20000.0]])
    out, _, st = _lib.krige_host(known, unk, _lib.MODEL_IDS["spherical"], params, _lib.KRIGE_ORDINARY, 64)
    assert (st == 0).all() and np.isfinite(out[0, :, :2]).all()
    a, _, _ = _lib.krige_host(known, unk[:700_001], _lib.MODEL_IDS["spherical"], params, _lib.KRIGE_ORDINARY, 64)
    np.testing.assert_array_equal(a[0], out[0, :700_001])
    d_out, _, _ = _lib.krige_device(torch.from_numpy(known).cuda(), torch.from_numpy(unk[:100_000]).cuda(),
                                    _lib.MODEL_IDS["spherical"], params, _lib.KRIGE_ORDINARY, 64)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(d_out[0].cpu().numpy(), out[0, :100_000])
    # predictions stay inside the data range up to the kriging overshoot one expects from a smooth field
    assert out[0, :, 0].min() > known[:, 2].min() - 10 and out[0, :, 0].max() < known[:, 2].max() + 10


def test_more_locations_than_one_chunk(pib):
    """The library processes locations in chunks of 2^21; 2.6M locations cross a chunk boundary
    (BASELINE config 4 has 10M).  Results must not depend on where the boundaries fall."""
    from pyinterpolate_b200 import _lib
    known = dem_like_field(200_000, seed=4)
    m = (1 << 21) + 500_001
    unk = uniform_locations(m, seed=15)
    params = np.array([[1.0, float(np.var(known[:, 2])), 20000.0]])
    out, _, st = _lib.krige_host(known, unk, _lib.MODEL_IDS["spherical"], params, _lib.KRIGE_ORDINARY, 8)
    assert (st == 0).all() and np.isfinite(out[0, :, :2]).all()
    np.testing.assert_array_equal(out[0, :, 2:], unk)
    lo = (1 << 21) - 1000
    part, _, _ = _lib.krige_host(known, unk[lo:lo + 3000], _lib.MODEL_IDS["spherical"], params, _lib.KRIGE_ORDINARY, 8)
    np.testing.assert_array_equal(part[0], out[0, lo:lo + 3000])


def test_baseline_config3_full_size(pib):
    """BASELINE config 3 at full size: simple kriging of 10M locations from 1M known points, k=64 (five chunks).
    Size-independent property: a random subset kriged on its own reproduces its rows bit for bit."""
    known = dem_like_field(1_000_000, seed=0)
    unk = uniform_locations(10_000_000, seed=1)
    mdl = {"variogram_model_type": "spherical", "nugget": 1.0, "sill": float(np.var(known[:, 2])), "rang": 5000.0}
    pm = float(known[:, 2].mean())
    sk = pib.simple_kriging(mdl, known, unk, process_mean=pm, no_neighbors=64)
    assert sk.shape == (10_000_000, 4) and np.isfinite(sk[:, 0]).all()
    np.testing.assert_array_equal(sk[:, 2:], unk)
    sub = np.random.default_rng(2).choice(len(unk), 3000, replace=False)
    np.testing.assert_array_equal(pib.simple_kriging(mdl, known, unk[sub], process_mean=pm, no_neighbors=64), sk[sub])


def test_baseline_config4_full_size(pib):
    """BASELINE config 4 at full size: indicator kriging with 8 thresholds on 500k points (one neighbour search,
    eight solves per location)."""
    known = dem_like_field(500_000, seed=3)
    unk = uniform_locations(500_000, seed=4)
    thr = np.quantile(known[:, 2], np.linspace(0.1, 0.9, 8))
    models = [{"variogram_model_type": "spherical", "nugget": 0.02, "sill": 0.2, "rang": 4000.0 + 250.0 * t} for t in range(8)]
    ik = pib.kriging.indicator_kriging(models, thr, known, unk, no_neighbors=16, reference_variance_column=False)
    assert ik.shape == (500_000, 8) and ik.min() >= 0.0 and ik.max() <= 1.0
    assert np.all(np.diff(ik.mean(axis=0)) > 0)                   # P(z <= threshold) grows with the threshold
    sub = np.random.default_rng(5).choice(len(unk), 3000, replace=False)
    np.testing.assert_array_equal(
        pib.kriging.indicator_kriging(models, thr, known, unk[sub], no_neighbors=16, reference_variance_column=False), ik[sub])
