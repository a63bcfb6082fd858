"""Callers of the batched ordinary-kriging path, re-pointed at the CUDA kernels.

Mirrors (file:line relative to /root/reference/src/pyinterpolate/):
  * core/pipelines/interpolate.py:95-204    interpolate_points_dask  (the Dask fan-out is pointless here:
                                            one batched launch kriges every location)
  * viz/raster.py:16-60, :62-189            set_dimensions, interpolate_raster
  * kriging/point/universal.py:13-86        MultivariateRegression (host least squares, a handful of columns)
  * kriging/point/universal.py:88-326       UniversalKriging: trend + kriged residuals

Model FITTING (TheoreticalVariogram.autofit) is outside the hot path and is not rebuilt: where the reference
fits a model on the fly the caller passes one (a dict, this package's TheoreticalVariogram or the reference's).
"""
import numpy as np

from .kriging import interpolate_points
from .variogram import ExperimentalVariogram


def interpolate_points_dask(theoretical_model, known_locations, unknown_locations, neighbors_range=None,
                            no_neighbors=4, max_tick=5., use_all_neighbors_in_range=False,
                            allow_approximate_solutions=False, number_of_workers=1, progress_bar=True):
    """core/pipelines/interpolate.py:95-204.  ``number_of_workers`` is accepted and ignored."""
    return interpolate_points(theoretical_model, known_locations, unknown_locations, neighbors_range, no_neighbors,
                              max_tick, use_all_neighbors_in_range, allow_approximate_solutions, progress_bar)


def set_dimensions(xs, ys, dmax):
    """viz/raster.py:16-60: pixel centres along x and (top to bottom) along y; the longer side gets ``dmax`` pixels."""
    xmin, xmax, ymin, ymax = np.min(xs), np.max(xs), np.min(ys), np.max(ys)
    x_abs, y_abs = np.abs(xmax - xmin), np.abs(ymax - ymin)
    step = (x_abs if x_abs > y_abs else y_abs) / dmax
    x_dim_coords = np.arange(xmin + step, xmax + step, step)
    y_dim_coords = np.arange(ymin + step, ymax + step, step)[::-1]
    return x_dim_coords, y_dim_coords, [step, xmin, xmax, ymin, ymax]


def interpolate_raster(data, dim=1000, number_of_neighbors=4, semivariogram_model=None, direction=None,
                       tolerance=None, allow_approx_solutions=True):
    """viz/raster.py:62-189: ordinary kriging of every pixel centre (row-major, y descending).

    ``semivariogram_model`` is required: the reference fits one with ``TheoreticalVariogram.autofit`` when it is
    None (:140-154), and model fitting is not part of this package."""
    data = np.asarray(data, dtype=np.float64)
    if semivariogram_model is None:
        raise NotImplementedError("interpolate_raster needs semivariogram_model: fitting a model on the fly "
                                  "(TheoreticalVariogram.autofit) is outside pyinterpolate_b200")
    x_coords, y_coords, props = set_dimensions(data[:, 0], data[:, 1], dim)
    gx, gy = np.meshgrid(x_coords, y_coords)                     # the reference's double loop: y outer, x inner
    pixels = np.column_stack([gx.ravel(), gy.ravel()])
    k = interpolate_points(theoretical_model=semivariogram_model, known_locations=data, unknown_locations=pixels,
                           no_neighbors=number_of_neighbors, allow_approximate_solutions=allow_approx_solutions)
    shape = (len(y_coords), len(x_coords))
    return {'result': k[:, 0].reshape(shape), 'error': k[:, 1].reshape(shape),
            'params': {'pixel size': props[0], 'min x': props[1], 'max x': props[2],
                       'min y': props[3], 'max y': props[4]}}


class MultivariateRegression:
    """kriging/point/universal.py:13-86: ordinary least squares with an intercept."""

    def __init__(self):
        self.features = None
        self.y = None
        self.coefficients = None
        self.intercept = None

    def fit(self, dataset):
        dataset = np.asarray(dataset, dtype=np.float64)
        self.features = dataset[:, :-1].copy()
        self.y = dataset[:, -1].copy()
        design = np.c_[self.features, np.ones(self.features.shape[0])]
        params = np.linalg.lstsq(design, self.y, rcond=None)[0]
        self.coefficients, self.intercept = params[:-1], params[-1]

    def predict(self, features):
        return np.matmul(self.coefficients, np.asarray(features, dtype=np.float64).T) + self.intercept


class UniversalKriging:
    """kriging/point/universal.py:88-326: regression trend on the coordinates + ordinary kriging of the residuals.

    ``fit_bias`` computes the residuals' experimental variogram on the GPU; the theoretical model of the bias is
    passed in (``bias_model=``) or assigned to ``self.bias_model``, since model fitting is not part of this package."""

    def __init__(self, known_points, fitted_regression_model=None):
        self.known_points = np.asarray(known_points, dtype=np.float64)
        self.trend_model = None
        self.trend_values = None
        if fitted_regression_model is not None:
            self.trend_model = fitted_regression_model
            self.trend_values = self.trend_model.predict(self.known_points[:, :-1])
        self.bias_values = None
        self.bias_experimental_model = None
        self.bias_model = None

    def fit_trend(self):
        trend_model = MultivariateRegression()
        trend_model.fit(self.known_points)
        self.trend_model = trend_model
        self.trend_values = self.trend_model.predict(features=self.known_points[:, :-1])

    def detrend(self, trend_values=None):
        base = self.trend_values if trend_values is None else trend_values
        self.bias_values = self.known_points[:, -1] - base

    def fit_bias(self, step_size, max_range, use_all_models=False, bias_model=None):
        if self.bias_values is None:
            self.detrend()
        bias_arr = np.c_[self.known_points[:, :-1], self.bias_values]
        self.bias_experimental_model = ExperimentalVariogram(ds=bias_arr, step_size=step_size, max_range=max_range,
                                                             is_semivariance=True, is_covariance=False)
        if bias_model is not None:
            self.bias_model = bias_model
        elif self.bias_model is None:
            raise NotImplementedError("UniversalKriging.fit_bias computed the experimental variogram of the bias "
                                      "(self.bias_experimental_model); pass bias_model= or set self.bias_model — "
                                      "TheoreticalVariogram.autofit is outside pyinterpolate_b200")

    def predict(self, points, neighbors_range=None, no_neighbors=4, use_all_neighbors_in_range=False,
                allow_approx_solutions=False, number_of_workers=1, show_progress_bar=True):
        """``[predicted value, x, y]`` rows (universal.py:240-326)."""
        points = np.asarray(points, dtype=np.float64)
        trends = self.trend_model.predict(points)
        biases = interpolate_points(theoretical_model=self.bias_model,
                                    known_locations=np.c_[self.known_points[:, :-1], self.bias_values],
                                    unknown_locations=points, neighbors_range=neighbors_range,
                                    no_neighbors=no_neighbors, use_all_neighbors_in_range=use_all_neighbors_in_range,
                                    allow_approximate_solutions=allow_approx_solutions,
                                    progress_bar=show_progress_bar)[:, 0]
        return np.c_[trends + biases, points[:, 0], points[:, 1]]
