"""Indicator variograms and indicator kriging on top of the CUDA paths.

Mirrors (file:line relative to /root/reference/src/pyinterpolate/):
  * semivariogram/indicator/indicator.py:14-71     code_indicators, select_variogram_thresholds
  * semivariogram/indicator/indicator.py:74-279    IndicatorVariogramData, ExperimentalIndicatorVariogram
  * kriging/point/indicator.py:17-403              IndicatorKriging

The reference builds one ExperimentalVariogram per threshold (a full pair pass each) and kriges every
(threshold, location) with a separate single-point call.  Here each threshold is one kernel pass over the
0/1 values and all thresholds of IndicatorKriging share ONE neighbour search (`indicator_kriging`).
Theoretical models are fitted by the reference (out of scope); any object with a
``theoretical_indicator_variograms`` dict {threshold: model} is accepted.
"""
import numpy as np

from .kriging import indicator_kriging
from .variogram import ExperimentalVariogram


def select_variogram_thresholds(ds, n_thresh):
    """semivariogram/indicator/indicator.py:47-71 — quantiles q = 1/T .. 1."""
    quantiles = np.linspace(0, 1, n_thresh + 1)
    return [np.quantile(ds, q=q) for q in quantiles[1:]]


def code_indicators(ds, thresholds):
    """``[x, y, ind_1, ..., ind_T]`` with ind_t = (value <= threshold_t) (indicator.py:14-44)."""
    ds = np.asarray(ds)
    ind = (ds[:, 2][:, None] <= np.asarray(thresholds)[None, :]).astype(int)
    return np.column_stack([ds[:, 0], ds[:, 1], ind])


class IndicatorVariogramData:
    def __init__(self, ds, number_of_thresholds):
        if not isinstance(ds, np.ndarray):
            ds = np.array(ds)
        self.input_array = ds
        self.n_thresholds = number_of_thresholds
        self.thresholds = select_variogram_thresholds(ds[:, -1], self.n_thresholds)
        self.ids = code_indicators(ds, self.thresholds)


class ExperimentalIndicatorVariogram:
    """semivariogram/indicator/indicator.py:120-279; ``experimental_models`` maps str(threshold) to an
    ExperimentalVariogram of the indicator-coded values."""

    def __init__(self, ds, number_of_thresholds, step_size, max_range, custom_weights=None, custom_bins=None,
                 direction=None, tolerance=None, dir_neighbors_selection_method='t', fit=True):
        self.ds = IndicatorVariogramData(ds=ds, number_of_thresholds=number_of_thresholds)
        self.step_size = step_size
        self.max_range = max_range
        self.custom_weights = custom_weights
        self.custom_bins = custom_bins
        self.direction = direction
        self.tolerance = tolerance
        self.dir_neighbors_selection_method = dir_neighbors_selection_method
        self.experimental_models = {}
        if fit:
            self.fit()

    def fit(self):
        for idx, indicator in enumerate(self.ds.thresholds):
            self.experimental_models[str(indicator)] = ExperimentalVariogram(
                ds=self.ds.ids[:, [0, 1, 2 + idx]], step_size=self.step_size, max_range=self.max_range,
                custom_weights=self.custom_weights, custom_bins=self.custom_bins, direction=self.direction,
                tolerance=self.tolerance, dir_neighbors_selection_method=self.dir_neighbors_selection_method,
                is_semivariance=True, is_covariance=True)


class IndicatorKriging:
    """kriging/point/indicator.py:17-403.  ``indicator_predictions`` is ``[M, T]`` clipped to [0, 1].

    Reference quirks kept on purpose: with ``kriging_type='ok'`` the stored column is ``_pred_arr[0][1]``,
    the kriging VARIANCE (indicator.py:312-313) — pass ``reference_variance_column=False`` for the indicator
    predictions themselves; ``kriging_type='sk'`` raises TypeError because the reference passes an unknown
    keyword to simple_kriging (indicator.py:294-304)."""

    def __init__(self, indicator_variograms, known_locations, unknown_locations, kriging_type='ok',
                 process_mean=None, neighbors_range=None, no_neighbors=4, max_tick=5.,
                 use_all_neighbors_in_range=False, allow_approximate_solutions=False, get_expected_values=True,
                 reference_variance_column=True):
        models = indicator_variograms.theoretical_indicator_variograms
        self.thresholds = np.array(list(models.keys())).astype(float)
        self.coordinates = np.array(unknown_locations)
        if kriging_type == 'sk':
            raise TypeError("simple_kriging() got an unexpected keyword argument 'unknown_location'")
        if kriging_type != 'ok':
            raise ValueError('Kriging type not supported. Please choose from: '
                             '"ok" - ordinary kriging, "sk" - simple kriging.')
        self.indicator_predictions = indicator_kriging(
            list(models.values()), self.thresholds, known_locations, unknown_locations, no_neighbors=no_neighbors,
            reference_variance_column=reference_variance_column, neighbors_range=neighbors_range,
            use_all_neighbors_in_range=use_all_neighbors_in_range,
            allow_approximate_solutions=allow_approximate_solutions)
        self.expected_values = None
        self.variances = None
        if get_expected_values:
            self.expected_values, self.variances = self._get_expected_values()

    def get_indicator_maps(self):
        return {thr: np.column_stack([self.coordinates[:, 0], self.coordinates[:, 1], self.indicator_predictions[:, i]])
                for i, thr in enumerate(self.thresholds)}

    def get_expected_values_maps(self):
        if self.expected_values is None:
            self.expected_values, self.variances = self._get_expected_values()
        xy = self.coordinates
        return (np.column_stack([xy[:, 0], xy[:, 1], self.expected_values]),
                np.column_stack([xy[:, 0], xy[:, 1], self.variances]))

    def _get_expected_values(self, ccdf_density=100):
        """Expected value and variance per location from the indicator predictions (what indicator.py:327-380 computes
        with one spline fit per location).  The interpolating spline through (threshold, prediction) is LINEAR in the
        predictions, so its values on the ``ccdf_density`` grid are one matrix product for all locations: the basis is the
        spline of each unit vector (T small fits, same scipy call and degree rule as the reference)."""
        from scipy.interpolate import UnivariateSpline
        thr = self.thresholds.astype(float)
        grid = np.linspace(thr.min(), thr.max(), ccdf_density)
        degree = 3 if len(thr) >= 4 else len(thr) - 1
        basis = np.empty((len(thr), ccdf_density))
        for t in range(len(thr)):
            unit = np.zeros(len(thr))
            unit[t] = 1.0
            basis[t] = UnivariateSpline(thr, unit, k=degree, s=0)(grid)
        curves = np.asarray(self.indicator_predictions, dtype=float) @ basis          # [M, ccdf_density]
        with np.errstate(divide="ignore", invalid="ignore"):
            ccdf = 1.0 - np.cumsum(curves, axis=1) / np.sum(curves, axis=1, keepdims=True)
        level = ccdf.mean(axis=1, keepdims=True)
        # the reference looks ``level`` up in the reversed ccdf with np.searchsorted (left side): for a non-increasing
        # ccdf that is the number of entries >= level; rows where the spline makes the ccdf non-monotone fall back to
        # the literal search
        n_ge = np.sum(ccdf >= level, axis=1)
        monotone = np.all(np.diff(ccdf, axis=1) <= 0, axis=1)
        for r in np.nonzero(~monotone)[0]:
            n_ge[r] = ccdf_density - np.searchsorted(ccdf[r, ::-1], level[r, 0])
        expected = grid[np.minimum(n_ge, ccdf_density - 1)]
        variances = np.mean((curves - expected[:, None]) ** 2, axis=1)
        return expected, variances
