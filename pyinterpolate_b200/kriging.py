"""Point kriging — the reference's public calls, backed by the batched CUDA kNN + solve kernels.

Mirrors (file:line relative to /root/reference/src/pyinterpolate/):
  * kriging/point/ordinary.py:134-221     ordinary_kriging        (+ ok_calc :30-131)
  * kriging/point/simple.py:129-223       simple_kriging          (+ sk_calc :26-126)
  * core/pipelines/interpolate.py:12-92   interpolate_points
  * kriging/point/indicator.py:200-325    IndicatorKriging._estimate (batched: one neighbour search,
                                          one solve per threshold)

Models that carry a ``direction`` use the reference's angular neighbour selection
(transform/select_points.py:104-257) inside ``neighbors_range`` (default: the model range).
``use_all_neighbors_in_range=True`` is supported up to 128 rows per location (more raises NotImplementedError).
``allow_approximate_solutions=True`` solves singular systems like ``np.linalg.lstsq`` on the device
(kriging/utils/point_kriging_solve.py:134-147) and warns like the reference.
"""
import warnings

import numpy as np

from . import _lib, core
from .models import model_parameters

# kriging/utils/errors.py:10-15, verbatim message
_SINGULAR_MSG = ("Singular matrix in Kriging system detected, "
                 "check if you have duplicated coordinates "
                 "in the ``known_locations`` variable. If your "
                 "data doesn't have duplicates then set "
                 "``allow_lsa`` parameter to ``True``.")


def _prepare(theoretical_model, known_locations, unknown_locations, use_all_neighbors_in_range,
             allow_approximate_solutions, neighbors_range=None, max_tick=5.):
    mid, nugget, sill, rang, direction = model_parameters(theoretical_model)
    known = np.ascontiguousarray(core.variogram_points(known_locations), dtype=np.float64)
    unknown = np.ascontiguousarray(core.interpolation_points(unknown_locations), dtype=np.float64)
    if known.ndim != 2 or known.shape[1] != 3:
        raise ValueError("known_locations must be rows of [x, y, value]")
    if unknown.ndim != 2 or unknown.shape[1] != 2:
        raise ValueError("unknown_locations must be rows of [x, y]")
    # a finite sum means every term is finite (NaN and +-inf propagate); only an overflowing sum (or a non-finite VALUE, which is
    # not an error here) needs the exact scan of the coordinate columns.  The sum runs over the whole contiguous array: the
    # strided `known[:, :2].sum()` cost 3.5 ms per million points, most of the fixed cost of a call
    if not (np.isfinite(known.sum()) and np.isfinite(unknown.sum())):
        if not (np.isfinite(known[:, :2]).all() and np.isfinite(unknown).all()):
            raise ValueError("coordinates must be finite")
    # neighbors_range defaults to the model range (kriging/utils/point_kriging_solve.py:68-69)
    selection = dict(neighbors_range=rang if neighbors_range is None else float(neighbors_range),
                     direction=direction, max_tick=max_tick, use_all=bool(use_all_neighbors_in_range),
                     allow_lsa=bool(allow_approximate_solutions))
    return mid, np.array([[nugget, sill, rang]]), known, unknown, selection


# core/ptp_warnings/ptp_warnings.py:8-34 — the reference warns with the message STRING (a UserWarning)
_ZEROS_MSG = repr('Matrix in your Kriging system is populated only be zeros. It is probably data error. '
                  'Prediction will return 0 as a predicted value and np.nan as an error.')
_LSA_MSG = repr('Kriging system solution is based on the approximate solution, output may be incorrect!')


def _raise_on_singular(status):
    # kriging/utils/point_kriging_solve.py:134-147
    if not status.any():                  # one pass: every system was regular
        return
    if (status == _lib.STATUS_ZERO_MATRIX).any():
        warnings.warn(_ZEROS_MSG)
    if (status == _lib.STATUS_LSA).any():
        warnings.warn(_LSA_MSG)
    if (status == _lib.STATUS_SINGULAR).any():
        # kriging/utils/errors.py:1-15 via ordinary.py:130-131
        raise RuntimeError(_SINGULAR_MSG)
    if (status == _lib.STATUS_TOO_MANY_NEIGHBORS).any():
        raise NotImplementedError("use_all_neighbors_in_range selected more than 128 neighbours for a location; "
                                  "pyinterpolate_b200 solves systems of at most 128 neighbours")


def ordinary_kriging(theoretical_model, known_locations, unknown_locations=None, neighbors_range=None,
                     no_neighbors=4, max_tick=5., use_all_neighbors_in_range=False,
                     allow_approximate_solutions=False, progress_bar=True, unknown_location=None,
                     return_neighbors=False):
    """``[predicted value, variance error, x, y]`` rows — kriging/point/ordinary.py:134-221.

    For isotropic models ``neighbors_range`` does not change the result when ``use_all_neighbors_in_range`` is
    False (transform/select_points.py:95-101: the k nearest are used either way); for models with a
    ``direction`` it bounds the angular search and ``max_tick`` caps the tolerance widening.
    ``progress_bar`` has nothing to show: one batched launch replaces the loop.
    ``unknown_location=`` (the README's spelling) is accepted as an alias."""
    if unknown_locations is None:
        unknown_locations = unknown_location
    mid, params, known, unknown, sel = _prepare(theoretical_model, known_locations, unknown_locations,
                                                use_all_neighbors_in_range, allow_approximate_solutions,
                                                neighbors_range, max_tick)
    out, idx, status = _lib.krige_host(known, unknown, mid, params, _lib.KRIGE_ORDINARY, int(no_neighbors),
                                       want_neighbors=return_neighbors, **sel)
    _raise_on_singular(status)
    return (out[0], idx) if return_neighbors else out[0]


def simple_kriging(theoretical_model, known_locations, unknown_locations=None, process_mean=None,
                   neighbors_range=None, no_neighbors=4, max_tick=5., use_all_neighbors_in_range=False,
                   allow_approximate_solutions=False, progress_bar=True, unknown_location=None,
                   return_neighbors=False):
    """``[predicted value, variance error, x, y]`` rows — kriging/point/simple.py:129-223."""
    if unknown_locations is None:
        unknown_locations = unknown_location
    if process_mean is None:
        raise TypeError("simple_kriging() missing required argument: 'process_mean'")
    mid, params, known, unknown, sel = _prepare(theoretical_model, known_locations, unknown_locations,
                                                use_all_neighbors_in_range, allow_approximate_solutions,
                                                neighbors_range, max_tick)
    out, idx, status = _lib.krige_host(known, unknown, mid, params, _lib.KRIGE_SIMPLE, int(no_neighbors),
                                       process_mean=float(process_mean), want_neighbors=return_neighbors, **sel)
    _raise_on_singular(status)
    return (out[0], idx) if return_neighbors else out[0]


def ordinary_and_simple_kriging(theoretical_model, known_locations, unknown_locations, process_mean, neighbors_range=None,
                                no_neighbors=4, max_tick=5., use_all_neighbors_in_range=False,
                                allow_approximate_solutions=False):
    """``(ordinary_kriging(...), simple_kriging(...))`` from ONE pass: the two calls of the reference
    (kriging/point/ordinary.py:134-221, simple.py:129-223) select the same neighbours and build the same Gamma
    (ordinary.py:100-120, simple.py:102-113), so the neighbour search, the assembly and the elimination are shared
    (BASELINE config 4 asks for both on 10M locations).  Each result equals the single call bit for bit."""
    mid, params, known, unknown, sel = _prepare(theoretical_model, known_locations, unknown_locations,
                                                use_all_neighbors_in_range, allow_approximate_solutions,
                                                neighbors_range, max_tick)
    out, _, status = _lib.krige_host(known, unknown, mid, params, _lib.KRIGE_BOTH, int(no_neighbors),
                                     process_mean=float(process_mean), **sel)
    _raise_on_singular(status)
    return out[0, 0], out[1, 0]


def validate_kriging(points, theoretical_model, how='ok', neighbors_range=None, no_neighbors=4,
                     use_all_neighbors_in_range=False, sk_mean=None, allow_approximate_solutions=False):
    """Leave-one-out cross-validation — evaluate/cross_validation.py:11-138 — as ONE batched launch: every known
    point is kriged from the others (the neighbour search skips the point's own row, which is what
    ``np.delete(points, idx, 0)`` does in the reference's loop).

    Returns ``(mean prediction error, mean variance error, [x, y, prediction error, kriging variance] rows)``."""
    if how not in ('ok', 'sk'):
        raise KeyError('Allowed kriging types (parameter "how") are: "ok" - ordinary kriging, and "sk" - simple kriging.')
    mid, params, known, _, sel = _prepare(theoretical_model, points, np.zeros((1, 2)), use_all_neighbors_in_range,
                                          allow_approximate_solutions, neighbors_range)
    if how == 'sk' and sk_mean is None:
        raise TypeError("validate_kriging(how='sk') needs sk_mean")
    out, status = _lib.krige_loo_host(known, mid, params, _lib.KRIGE_ORDINARY if how == 'ok' else _lib.KRIGE_SIMPLE,
                                      int(no_neighbors), process_mean=0.0 if sk_mean is None else float(sk_mean), **sel)
    _raise_on_singular(status)
    output_arr = np.column_stack([out[:, 2], out[:, 3], known[:, 2] - out[:, 0], out[:, 1]])
    mean_prediction_error = np.mean(output_arr[:, 2])
    mean_variance_error = np.var(output_arr[:, 2]) / np.mean(output_arr[:, 3])
    return mean_prediction_error, mean_variance_error, output_arr


def interpolate_points(theoretical_model, known_locations, unknown_locations, neighbors_range=None,
                       no_neighbors=4, max_tick=5., use_all_neighbors_in_range=False,
                       allow_approximate_solutions=False, progress_bar=True):
    """core/pipelines/interpolate.py:12-92 — ordinary kriging of many points."""
    return ordinary_kriging(theoretical_model, known_locations, unknown_locations, neighbors_range, no_neighbors,
                            max_tick, use_all_neighbors_in_range, allow_approximate_solutions, progress_bar)


def indicator_kriging(theoretical_models, thresholds, known_locations, unknown_locations, no_neighbors=4,
                      reference_variance_column=True, neighbors_range=None, use_all_neighbors_in_range=False,
                      allow_approximate_solutions=False):
    """Batched core of kriging/point/indicator.py:200-325 for ``kriging_type='ok'``.

    ``theoretical_models`` are the per-threshold models (same model family), ``thresholds`` the cut values;
    indicators are ``(z <= threshold)`` (indicator.py:272-275).  All thresholds share one neighbour search
    (with ``use_all_neighbors_in_range`` and per-threshold ranges that differ, the selection differs per
    threshold, so each threshold gets its own call).
    Returns ``[M, T]`` clipped to [0, 1].  With ``reference_variance_column=True`` the column the reference
    stores is reproduced — it keeps ``_pred_arr[0][1]``, the kriging VARIANCE, not zhat (indicator.py:312-313);
    pass False to get the indicator predictions themselves."""
    known = np.ascontiguousarray(core.variogram_points(known_locations), dtype=np.float64)
    unknown = np.ascontiguousarray(core.interpolation_points(unknown_locations), dtype=np.float64)
    parsed = [model_parameters(m) for m in theoretical_models]
    if len({p[0] for p in parsed}) != 1:
        raise NotImplementedError("indicator models of different families need one call per family")
    if any(p[4] is not None for p in parsed):
        raise NotImplementedError("directional indicator models are not implemented")
    params = np.array([[p[1], p[2], p[3]] for p in parsed])
    z = np.ascontiguousarray(known[:, 2])
    values = (z[None, :] <= np.asarray(thresholds, dtype=np.float64)[:, None]).astype(np.float64)     # [T, N] indicator columns
    sel = dict(use_all=bool(use_all_neighbors_in_range), allow_lsa=bool(allow_approximate_solutions))
    ranges = params[:, 2] if neighbors_range is None else np.full(len(parsed), float(neighbors_range))
    col = 1 if reference_variance_column else 0
    if use_all_neighbors_in_range and len(set(ranges.tolist())) > 1:
        outs, stats = [], []
        for t in range(len(parsed)):
            o, _, st = _lib.krige_host(known, unknown, parsed[0][0], params[t:t + 1], _lib.KRIGE_ORDINARY,
                                       int(no_neighbors), values=values[t:t + 1], neighbors_range=float(ranges[t]), **sel)
            outs.append(o[0])
            stats.append(st[0])
        out, status = np.stack(outs), np.stack(stats)
        cols = out[:, :, col]
    else:
        # only the column the caller keeps travels back (pib_krige_column_host): [T, M] instead of [T, M, 4]
        cols, status = _lib.krige_column_host(known, unknown, parsed[0][0], params, _lib.KRIGE_ORDINARY, int(no_neighbors), col,
                                              values=values, neighbors_range=float(ranges[0]), **sel)
    _raise_on_singular(status)
    # indicator.py:399-403 sets values below 0 to 0 and above 1 to 1 (NaN stays NaN): one clipping pass into the [M, T] result
    res = np.empty((cols.shape[1], cols.shape[0]))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        np.clip(cols.T, 0.0, 1.0, out=res)
    return res
