"""allow_approximate_solutions=True on data with duplicated coordinates (singular kriging systems solved by
np.linalg.lstsq, kriging/utils/point_kriging_solve.py:134-147) — fixtures from the UNMODIFIED reference.
    python tests/golden/make_golden_lsa.py"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
warnings.filterwarnings("ignore")
from ref_import import import_reference  # noqa: E402

import_reference()
from pyinterpolate import TheoreticalVariogram, ordinary_kriging, simple_kriging  # noqa: E402

from pyinterpolate_b200.synthetic import dem_like_field, uniform_locations  # noqa: E402


def duplicated_field(n, n_dup, seed):
    """n rows of a DEM-like field followed by n_dup exact copies (same x, y AND z, so that which copy a tie in
    the neighbour sort keeps does not change the answer)."""
    base = dem_like_field(n, seed=seed)
    rng = np.random.default_rng(seed + 1)
    dup = base[rng.choice(n, n_dup, replace=False)]
    return np.vstack([base, dup])


def main():
    g = {}
    known = duplicated_field(3000, 100, seed=71)
    unk = uniform_locations(120, seed=73)
    var = float(np.var(known[:, 2]))
    g["known"] = known
    g["unknown"] = unk
    g["var"] = np.array(var)
    # Which rows the reference sends through lstsq depends on its LAPACK: OpenBLAS's blocked LU meets an exact zero
    # pivot (LinAlgError) for only some of the duplicate-induced singular systems and returns unbounded garbage for
    # the others.  Record per row whether lstsq ran, by counting calls around one-location calls.
    calls = []
    real_lstsq = np.linalg.lstsq

    def counting_lstsq(*a, **k):
        calls.append(1)
        return real_lstsq(*a, **k)

    np.linalg.lstsq = counting_lstsq
    for name, nugget in (("spherical", 0.0), ("exponential", 0.7), ("linear", 0.0)):
        m = TheoreticalVariogram()
        m.from_dict({"variogram_model_type": name, "nugget": nugget, "sill": var, "rang": 9000.0})
        ok_rows, sk_rows, ok_l, sk_l = [], [], [], []
        for p in unk:
            n0 = len(calls)
            ok_rows.append(ordinary_kriging(m, known, p[None, :], no_neighbors=16, allow_approximate_solutions=True,
                                            progress_bar=False)[0])
            ok_l.append(len(calls) > n0)
            n0 = len(calls)
            sk_rows.append(simple_kriging(m, known, p[None, :], process_mean=float(known[:, 2].mean()),
                                          no_neighbors=12, allow_approximate_solutions=True, progress_bar=False)[0])
            sk_l.append(len(calls) > n0)
        g[f"ok_{name}"], g[f"sk_{name}"] = np.array(ok_rows), np.array(sk_rows)
        g[f"ok_{name}_lstsq"], g[f"sk_{name}_lstsq"] = np.array(ok_l), np.array(sk_l)
        print(name, "lstsq rows ok/sk:", int(np.sum(ok_l)), int(np.sum(sk_l)))
    np.linalg.lstsq = real_lstsq
    # rows whose neighbour set holds a duplicated coordinate (exactly singular system)
    from pyinterpolate.transform.select_points import select_kriging_data
    sing = []
    sing12 = []
    for p in unk:
        nb = select_kriging_data(p, known, 9000.0, 16, False)
        xy = nb[:, :2]
        sing.append(len(np.unique(xy, axis=0)) < len(xy))
        sing12.append(len(np.unique(xy[:12], axis=0)) < 12)
    g["singular16"] = np.array(sing)
    g["singular12"] = np.array(sing12)
    # zero matrix: a model with nugget = sill = 0 (all gammas zero) -> simple kriging weights are zeros
    mz = TheoreticalVariogram()
    mz.from_dict({"variogram_model_type": "spherical", "nugget": 0.0, "sill": 0.0, "rang": 9000.0})
    g["sk_zero"] = simple_kriging(mz, known, unk[:10], process_mean=3.5, no_neighbors=6, progress_bar=False)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden_lsa.npz"), **g)
    print("singular rows (k=16):", int(np.sum(sing)), "of", len(sing))
    print(g["sk_zero"][:3])


if __name__ == "__main__":
    main()
