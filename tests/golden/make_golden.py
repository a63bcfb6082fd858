"""Generate tests/golden/*.npz by running the UNMODIFIED reference (pyinterpolate 1.0.3,
imported from /root/reference/src through oracle/ref_import.py) in the build container.

    python tests/golden/make_golden.py            # ~2-3 minutes on 8 vCPU

The fixtures hold the reference's OUTPUTS plus the small non-synthetic inputs; synthetic inputs
are regenerated from (n, extent, seed) by pyinterpolate_b200.synthetic.  Versions used are
recorded in each file (numpy / scipy decide the third-party arithmetic the reference delegates to).
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
warnings.filterwarnings("ignore")

from ref_import import import_reference  # noqa: E402

ref = import_reference()
import scipy  # noqa: E402
from pyinterpolate import (ExperimentalVariogram, TheoreticalVariogram, calculate_covariance,  # noqa: E402
                           calculate_semivariance, ordinary_kriging, simple_kriging)
from pyinterpolate.distance.angular import define_whitening_matrix, select_points_within_ellipse  # noqa: E402
from pyinterpolate.semivariogram.theoretical.variogram_models.functions import ALL_MODELS  # noqa: E402
from pyinterpolate.transform.select_points import select_kriging_data  # noqa: E402

from pyinterpolate_b200.synthetic import dem_like_field, uniform_locations  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
VERSIONS = np.array([f"numpy {np.__version__}", f"scipy {scipy.__version__}", "pyinterpolate 1.0.3"])

LINE13 = np.array([[0, 0, 8], [1, 0, 6], [2, 0, 4], [3, 0, 3], [4, 0, 6], [5, 0, 5], [6, 0, 7],
                   [7, 0, 2], [8, 0, 8], [9, 0, 9], [10, 0, 5], [11, 0, 6], [12, 0, 3]])
ARMSTRONG = np.load("/root/reference/tests/test_semivariogram/armstrong_data.npy")


def model(name, nugget, sill, rang):
    m = TheoreticalVariogram()
    m.from_dict({"variogram_model_type": name, "nugget": nugget, "sill": sill, "rang": rang})
    return m


def neighbour_indices(known, unknown, rang, k):
    """Row indices the reference's select_kriging_data picked (rows matched back by x,y,z)."""
    lut = {tuple(r): i for i, r in enumerate(known)}
    out = np.zeros((len(unknown), k), dtype=np.int64)
    for u, pt in enumerate(unknown):
        rows = select_kriging_data(pt, known, rang, k)
        out[u] = [lut[tuple(r[:3])] for r in rows]
    return out


def variogram_case(points, directional=None, **kw):
    if directional is None:
        s = calculate_semivariance(points, **kw)
        c = calculate_covariance(points, **kw)
    else:
        s = calculate_semivariance(points, dir_neighbors_selection_method="e", **directional, **kw)
        c = calculate_covariance(points, dir_neighbors_selection_method="e", **directional, **kw)
    return s, c


def main():
    g = {}
    # ---- variogram, small known-answer sets -------------------------------------------------
    g["line13_points"] = LINE13
    g["line13_sem"], g["line13_cov"] = variogram_case(LINE13, step_size=1, max_range=4)
    g["armstrong_points"] = ARMSTRONG
    g["armstrong_omni_sem"], g["armstrong_omni_cov"] = variogram_case(ARMSTRONG, step_size=1, max_range=6)
    g["armstrong_ell_sem"], g["armstrong_ell_cov"] = variogram_case(
        ARMSTRONG, directional=dict(direction=15, tolerance=0.25), step_size=1.5, max_range=6)
    # custom bins (leading zero is dropped; uneven widths)
    bins = [0, 1.0, 1.5, 2.5, 4.0, 7.0]
    g["armstrong_custom_bins"] = np.array(bins)
    g["armstrong_custom_sem"] = calculate_semivariance(ARMSTRONG, custom_bins=bins)
    g["armstrong_custom_cov"] = calculate_covariance(ARMSTRONG, custom_bins=bins)
    bins_e = [0, 1.5, 2.5, 4.0, 7.0]
    g["armstrong_custom_ell_bins"] = np.array(bins_e)
    g["armstrong_custom_ell_sem"] = calculate_semivariance(
        ARMSTRONG, custom_bins=bins_e, direction=45, tolerance=0.5, dir_neighbors_selection_method="e")
    try:                                   # a directional lag without pairs raises (directional.py:68-74)
        calculate_semivariance(ARMSTRONG, custom_bins=bins, direction=45, tolerance=0.5,
                               dir_neighbors_selection_method="e")
        g["armstrong_empty_dir_lag_raises"] = np.array(False)
    except RuntimeError:
        g["armstrong_empty_dir_lag_raises"] = np.array(True)
    # omni: empty lags (nan rows) — sparse points, fine bins
    sparse = dem_like_field(40, extent=1000.0, seed=11)
    g["sparse_sem"] = calculate_semivariance(sparse, step_size=2.0, max_range=60.0)
    g["sparse_cov"] = calculate_covariance(sparse, step_size=2.0, max_range=60.0)
    # ExperimentalVariogram class attributes (semivariance + covariance + variance)
    ev = ExperimentalVariogram(ARMSTRONG, step_size=1, max_range=6)
    g["armstrong_ev_variance"] = np.array(ev.variance)
    g["armstrong_ev_lags"] = ev.lags
    g["armstrong_ev_sem"] = ev.semivariances
    g["armstrong_ev_cov"] = ev.covariances
    g["armstrong_ev_ppl"] = ev.points_per_lag

    # ---- variogram, synthetic fields --------------------------------------------------------
    f2k = dem_like_field(2000, seed=21)
    g["omni2k_sem"], g["omni2k_cov"] = variogram_case(f2k, step_size=500, max_range=40000)
    f10k = dem_like_field(10000, seed=0)                       # BASELINE config 1
    g["omni10k_sem"], g["omni10k_cov"] = variogram_case(f10k, step_size=500, max_range=40000)
    f15 = dem_like_field(1500, seed=22)
    for ang in (90, 0, 45, 135, 15):
        s, c = variogram_case(f15, directional=dict(direction=ang, tolerance=0.2),
                              step_size=2500, max_range=40000)
        g[f"ell1500_dir{ang}_sem"], g[f"ell1500_dir{ang}_cov"] = s, c
    # float step (lower = cur-(cur-prev) identity exercised), small extent
    f600 = dem_like_field(600, extent=10.0, seed=23)
    g["ell600_sem"], g["ell600_cov"] = variogram_case(
        f600, directional=dict(direction=30, tolerance=0.35), step_size=0.1, max_range=3.05)
    g["omni600_sem"], g["omni600_cov"] = variogram_case(f600, step_size=0.1, max_range=3.05)

    # ---- ellipse selection masks on the 5x5 grid (reference test_select_points_within_ellipse) ----
    grid = np.array([[x, y] for x in range(-2, 3) for y in range(-2, 3)], dtype=float)
    g["grid5_points"] = grid
    sel_cfg = [(90, 0.01), (0, 0.01), (135, 0.01), (45, 0.01), (90, 0.1), (0, 0.1), (45, 0.5), (15, 0.25)]
    g["grid5_cfg"] = np.array(sel_cfg)
    masks, Ws = [], []
    for ang, tol in sel_cfg:
        W = define_whitening_matrix(theta=ang, minor_axis_size=tol)
        Ws.append(W)
        masks.append(select_points_within_ellipse(np.array([0.0, 0.0]), grid, lag=2, step_size=2, w_matrix=W))
    g["grid5_masks"] = np.array(masks)
    g["grid5_W"] = np.array(Ws)
    angs = [0, 15, 30, 45, 90, 135, 180, 225, 270, 315, 359.5]
    tols = [0.01, 0.1, 0.2, 0.25, 0.5, 1.0]
    g["whiten_cfg"] = np.array([(a, t) for a in angs for t in tols])
    g["whiten_W"] = np.array([define_whitening_matrix(theta=a, minor_axis_size=t) for a in angs for t in tols])

    # ---- theoretical models ---------------------------------------------------------------------
    h = np.concatenate([[0.0], np.linspace(0.01, 12.0, 60), [4.0, 8.0]])
    g["gamma_h"] = h
    triples = [(0.0, 12.85, 4.0), (1.0, 5.0, 8.0), (0.3, 2.0, 20.0)]
    g["gamma_params"] = np.array(triples)
    g["gamma_models"] = np.array(ALL_MODELS)
    g["gamma_values"] = np.array([[model(name, *t).predict(h.copy()) for t in triples] for name in ALL_MODELS])

    # ---- kriging --------------------------------------------------------------------------------
    arm = ARMSTRONG.astype(float)
    sill = float(np.var(arm[:, -1]))
    mdl = model("spherical", 0.0, sill, 4.0)
    unk = np.array([[3.3, 4.7], [0.2, 0.1], [7.9, 6.4], [4.5, 4.5], [2.0, 3.0]])
    unk_no_exact = unk[:4]
    g["armstrong_krige_unknown"] = unk
    g["armstrong_krige_sill"] = np.array(sill)
    g["armstrong_ok"] = ordinary_kriging(mdl, arm, unk, no_neighbors=8, progress_bar=False)
    g["armstrong_sk"] = simple_kriging(mdl, arm, unk, process_mean=float(arm[:, -1].mean()),
                                       no_neighbors=8, progress_bar=False)
    # synthetic C1-shaped: 10k known, k=32, 200 unknown, five parity-testable models
    unk200 = uniform_locations(200, seed=1)
    var10k = float(np.var(f10k[:, -1]))
    g["krige10k_var"] = np.array(var10k)
    for name in ("linear", "spherical", "exponential", "circular", "cubic"):
        mdl = model(name, 0.0 if name == "linear" else 1.0, var10k, 8000.0)
        g[f"krige10k_{name}_ok"] = ordinary_kriging(mdl, f10k, unk200, no_neighbors=32, progress_bar=False)
        g[f"krige10k_{name}_sk"] = simple_kriging(mdl, f10k, unk200, process_mean=float(f10k[:, -1].mean()),
                                                  no_neighbors=32, progress_bar=False)
    g["krige10k_nbr_idx"] = neighbour_indices(f10k, unk200, 8000.0, 32)
    # gaussian / power: recorded for information (ill-conditioned, not parity-testable at 1e-9)
    for name in ("gaussian", "power"):
        mdl = model(name, 1.0, var10k, 8000.0)
        g[f"krige10k_{name}_ok"] = ordinary_kriging(mdl, f10k, unk200[:50], no_neighbors=32, progress_bar=False)
    # k=64 at C4 density, subsampled: 200k known on a 44.7 km field (same density as 1M on 100 km)
    f200k = dem_like_field(200_000, extent=44_721.36, seed=4)
    unk40 = uniform_locations(40, extent=44_721.36, seed=5)
    var200k = float(np.var(f200k[:, -1]))
    g["krige200k_var"] = np.array(var200k)
    mdl = model("spherical", 1.0, var200k, 20000.0)
    g["krige200k_ok"] = ordinary_kriging(mdl, f200k, unk40, no_neighbors=64, progress_bar=False)
    g["krige200k_sk"] = simple_kriging(mdl, f200k, unk40, process_mean=float(f200k[:, -1].mean()),
                                       no_neighbors=64, progress_bar=False)
    g["krige200k_nbr_idx"] = neighbour_indices(f200k, unk40, 20000.0, 64)
    # fewer known points than neighbours: uses all N
    few = dem_like_field(5, extent=100.0, seed=31)
    mdl = model("spherical", 0.5, 4.0, 60.0)
    g["few_ok"] = ordinary_kriging(mdl, few, np.array([[50.0, 50.0], [10.0, 80.0]]), no_neighbors=8, progress_bar=False)
    g["few_sk"] = simple_kriging(mdl, few, np.array([[50.0, 50.0], [10.0, 80.0]]), process_mean=200.0,
                                 no_neighbors=8, progress_bar=False)
    # neighbours beyond the model range (plateau values gamma = nugget + sill)
    mdl = model("spherical", 1.0, var10k, 500.0)
    g["krige10k_shortrange_ok"] = ordinary_kriging(mdl, f10k, unk200[:50], no_neighbors=16, progress_bar=False)
    mdl = model("exponential", 0.0, 9.0, 3.0)
    g["armstrong_exp_ok"] = ordinary_kriging(mdl, arm, unk_no_exact, no_neighbors=12, progress_bar=False)
    # duplicate coordinates -> singular -> RuntimeError in the reference
    dup = np.vstack([arm, arm[:3]])
    try:
        ordinary_kriging(model("spherical", 0.0, sill, 4.0), dup, np.array([[0.4, 0.3]]), no_neighbors=8,
                         progress_bar=False)
        g["dup_raises"] = np.array(False)
    except RuntimeError:
        g["dup_raises"] = np.array(True)

    g["versions"] = VERSIONS
    np.savez_compressed(os.path.join(OUT, "reference_golden.npz"), **g)
    print("wrote", len(g), "arrays")


if __name__ == "__main__":
    main()
