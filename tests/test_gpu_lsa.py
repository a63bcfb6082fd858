"""GPU parity for the singular-system fallbacks of solve_weights (kriging/utils/point_kriging_solve.py:134-147):
``allow_approximate_solutions=True`` (np.linalg.lstsq) and the zero-matrix rule.

The reference only reaches its fallback when LAPACK meets an EXACT zero pivot; for neighbour sets with duplicated
coordinates (two identical rows) that happens for roughly half of the systems with OpenBLAS, the others come back
as unbounded garbage.  Oracle and CUDA path treat every such system as singular, so the fixtures are
compared on the rows where the reference itself used lstsq and on every non-singular row; all rows are compared
against the oracle.  Tolerance: 1e-9 relative on regular rows; 1e-8 on least-squares rows (two different
backward-stable minimum-norm solvers — Jacobi eigen-decomposition here, LAPACK dgelsd in numpy — on systems with
condition numbers of 1e4..1e6)."""
import os
import warnings

import numpy as np
import pytest

from pyinterpolate_b200.synthetic import dem_like_field, uniform_locations

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pib():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pyinterpolate_b200 as pib
    return pib


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_golden_lsa.npz"))


def _close(got, want, rtol):
    np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
    m = ~np.isnan(want)
    np.testing.assert_allclose(got[m], want[m], rtol=rtol, atol=1e-9)


@pytest.mark.parametrize("name,nugget", [("spherical", 0.0), ("exponential", 0.7), ("linear", 0.0)])
def test_lstsq_rows_match_reference(pib, gold, name, nugget):
    known, unk, var = gold["known"], gold["unknown"], float(gold["var"])
    mdl = {"variogram_model_type": name, "nugget": nugget, "sill": var, "rang": 9000.0}
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        ok = pib.ordinary_kriging(mdl, known, unk, no_neighbors=16, allow_approximate_solutions=True)
        sk = pib.simple_kriging(mdl, known, unk, process_mean=float(known[:, 2].mean()), no_neighbors=12,
                                allow_approximate_solutions=True)
    assert sum("approximate solution" in str(x.message) for x in w) == 2          # one warning per call
    for got, key, k in ((ok, f"ok_{name}", 16), (sk, f"sk_{name}", 12)):
        ref, used_lstsq, sing = gold[key], gold[key + "_lstsq"], gold[f"singular{k}"]
        assert used_lstsq.sum() >= 10 and (~sing).sum() >= 10
        _close(got[~sing, :2], ref[~sing, :2], 1e-9)
        _close(got[used_lstsq, :2], ref[used_lstsq, :2], 1e-8)
        np.testing.assert_array_equal(got[:, 2:], ref[:, 2:])


@pytest.mark.parametrize("mode,k", [("ok", 16), ("sk", 12), ("ok", 100)])
def test_all_rows_match_oracle(pib, oracle, gold, mode, k):
    """Every row (also the duplicate systems the reference's LAPACK does not notice), status bytes included;
    k = 100 runs the shared-memory solver and the global-scratch eigenvector slab."""
    from pyinterpolate_b200 import _lib
    known, unk, var = gold["known"], gold["unknown"][:60], float(gold["var"])
    pm = float(known[:, 2].mean())
    want, _, wst = oracle.krige_with_fallback(known, unk, "spherical", 0.3, var, 9000.0, k, mode, pm)
    got, _, gst = _lib.krige_host(known, unk, 6, [[0.3, var, 9000.0]], _lib.KRIGE_ORDINARY if mode == "ok" else _lib.KRIGE_SIMPLE,
                                  k, process_mean=pm, allow_lsa=True)
    lsa = wst == 4
    assert lsa.sum() >= 10
    np.testing.assert_array_equal(gst[0] == 4, lsa)
    _close(got[0][~lsa, :2], want[~lsa, :2], 1e-9)
    _close(got[0][lsa, :2], want[lsa, :2], 1e-7 if k == 100 else 1e-8)
    # without the flag the same rows stay marked singular and the public call raises like the reference
    _, _, st0 = _lib.krige_host(known, unk, 6, [[0.3, var, 9000.0]], _lib.KRIGE_ORDINARY if mode == "ok" else _lib.KRIGE_SIMPLE,
                                k, process_mean=pm)
    np.testing.assert_array_equal(st0[0] == 1, lsa)


def test_duplicates_with_different_values(pib, oracle):
    """Copies of a location that carry different values: the minimum-norm solution splits the weight evenly."""
    base = dem_like_field(2000, seed=81)
    rng = np.random.default_rng(82)
    dup = base[rng.choice(2000, 150, replace=False)].copy()
    dup[:, 2] += rng.normal(0, 5, len(dup))
    known = np.vstack([base, dup])
    unk = uniform_locations(80, seed=83)
    var = float(np.var(known[:, 2]))
    want, _, wst = oracle.krige_with_fallback(known, unk, "exponential", 1.0, var, 7000.0, 24, "ok")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = pib.ordinary_kriging({"variogram_model_type": "exponential", "nugget": 1.0, "sill": var, "rang": 7000.0},
                                   known, unk, no_neighbors=24, allow_approximate_solutions=True)
    assert (wst == 4).sum() >= 10
    _close(got[:, :2], want[:, :2], 1e-8)


def test_zero_matrix_rule(pib, gold):
    known, unk = gold["known"], gold["unknown"][:10]
    mdl = {"variogram_model_type": "spherical", "nugget": 0.0, "sill": 0.0, "rang": 9000.0}
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        sk = pib.simple_kriging(mdl, known, unk, process_mean=3.5, no_neighbors=6)
    assert any("populated only be zeros" in str(x.message) for x in w)
    np.testing.assert_array_equal(sk, gold["sk_zero"])


def test_indicator_kriging_accepts_fallback_flags(pib, gold):
    """indicator.py:269-313 forwards use_all_neighbors_in_range / allow_approximate_solutions to ordinary_kriging."""
    known, unk, var = gold["known"], gold["unknown"][:40], float(gold["var"])
    thr = np.quantile(known[:, 2], [0.3, 0.6])
    models = [{"variogram_model_type": "spherical", "nugget": 0.01, "sill": 0.2, "rang": r} for r in (1500.0, 2500.0)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = pib.kriging.indicator_kriging(models, thr, known, unk, no_neighbors=8, use_all_neighbors_in_range=True,
                                            allow_approximate_solutions=True, reference_variance_column=False)
        for t in range(2):
            ind = known.copy()
            ind[:, 2] = (known[:, 2] <= thr[t]).astype(float)
            one = pib.ordinary_kriging(models[t], ind, unk, no_neighbors=8, use_all_neighbors_in_range=True,
                                       allow_approximate_solutions=True)
            np.testing.assert_allclose(got[:, t], np.clip(one[:, 0], 0, 1), rtol=1e-12, atol=1e-12)
