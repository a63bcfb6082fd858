"""use_all_neighbors_in_range=True with an isotropic model (transform/select_points.py:93-101) — fixtures from the
UNMODIFIED reference.   python tests/golden/make_golden_use_all.py"""
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
from pyinterpolate.transform.select_points import select_kriging_data  # noqa: E402

from pyinterpolate_b200.synthetic import dem_like_field, uniform_locations  # noqa: E402


def main():
    g = {}
    known = dem_like_field(20000, seed=61)
    unk = uniform_locations(150, seed=63)
    var = float(np.var(known[:, 2]))
    g["var"] = np.array(var)
    m = TheoreticalVariogram()
    m.from_dict({"variogram_model_type": "exponential", "nugget": 0.5, "sill": var, "rang": 2000.0})
    # model range 2000 m holds ~25 points: some locations have >= 24 in range (all are used), others fewer (24 nearest)
    g["ok"] = ordinary_kriging(m, known, unk, no_neighbors=24, use_all_neighbors_in_range=True, progress_bar=False)
    g["sk"] = simple_kriging(m, known, unk, process_mean=float(known[:, 2].mean()), no_neighbors=24,
                             use_all_neighbors_in_range=True, progress_bar=False)
    g["ok_range3000"] = ordinary_kriging(m, known, unk, neighbors_range=3000.0, no_neighbors=8,
                                         use_all_neighbors_in_range=True, progress_bar=False)
    g["counts"] = np.array([len(select_kriging_data(p, known, 2000.0, 24, True)) for p in unk])
    g["counts3000"] = np.array([len(select_kriging_data(p, known, 3000.0, 8, True)) for p in unk])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden_use_all.npz"), **g)
    print(np.bincount(g["counts"])[20:45], g["counts3000"].min(), g["counts3000"].max())


if __name__ == "__main__":
    main()
