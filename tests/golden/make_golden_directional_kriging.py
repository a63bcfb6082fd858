"""Directional-model kriging fixtures from the UNMODIFIED reference (angular neighbour selection,
transform/select_points.py:104-257).   python tests/golden/make_golden_directional_kriging.py"""
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
from pyinterpolate.transform.select_points import select_kriging_data_from_direction  # noqa: E402

from pyinterpolate_b200.synthetic import dem_like_field, uniform_locations  # noqa: E402


def model(direction, rang, var):
    m = TheoreticalVariogram()
    m.from_dict({"variogram_model_type": "spherical", "nugget": 1.0, "sill": var, "rang": rang, "direction": direction})
    return m


def main():
    g = {}
    known = dem_like_field(20000, seed=61)
    unk = uniform_locations(120, seed=62)
    var = float(np.var(known[:, 2]))
    g["var"] = np.array(var)
    cases = [("a", 30.0, 8000.0, 8, 5.0, False), ("b", 135.0, 3000.0, 16, 2.0, False), ("c", 0.0, 6000.0, 4, 10.0, False),
             ("d", 270.0, 1500.0, 6, 5.0, False), ("e", 45.0, 2500.0, 4, 3.0, True)]
    g["cases"] = np.array([[c[1], c[2], c[3], c[4], float(c[5])] for c in cases])
    lut = {tuple(r): i for i, r in enumerate(known)}
    for name, direction, rang, k, tick, use_all in cases:
        mdl = model(direction, rang, var)
        g[f"{name}_ok"] = ordinary_kriging(mdl, known, unk, no_neighbors=k, max_tick=tick,
                                           use_all_neighbors_in_range=use_all, progress_bar=False)
        g[f"{name}_sk"] = simple_kriging(mdl, known, unk, process_mean=float(known[:, 2].mean()), no_neighbors=k,
                                         max_tick=tick, use_all_neighbors_in_range=use_all, progress_bar=False)
        counts = []
        idx = np.full((len(unk), 128), -1, dtype=np.int64)
        for u, pt in enumerate(unk):
            rows = select_kriging_data_from_direction(pt, known, rang, direction, k, use_all, tick)
            if np.isnan(rows[0][0]):
                counts.append(0)
                continue
            counts.append(len(rows))
            idx[u, :len(rows)] = [lut[tuple(r[:3])] for r in rows][:128]
        g[f"{name}_counts"] = np.array(counts)
        g[f"{name}_idx"] = idx[:, :max(max(counts), 1)]
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden_directional_kriging.npz"), **g)
    print({k: (v.min(), v.max()) for k, v in g.items() if k.endswith("counts")})


if __name__ == "__main__":
    main()
