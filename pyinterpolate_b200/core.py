"""Host-side glue that stays in Python, restated from the reference so inputs, lags, whitening
matrices and error behaviour are identical (file:line relative to /root/reference/src/pyinterpolate/):

  * core/data_models/points.py:94-168                 VariogramPoints / InterpolationPoints coercion
  * core/validators/experimental_semivariance.py:51-153  validate_bins / direction+tolerance / weights
  * semivariogram/lags/lags.py:6-31                   get_lags
  * distance/angular.py:221-250, :620-638             define_whitening_matrix
"""
import numpy as np


def variogram_points(points):
    """``[x, y, value]`` rows as an ndarray (core/data_models/points.py:100-130).
    pandas / geopandas frames are accepted by duck typing (``.values``); geometry columns are
    unpacked when the first column holds shapely points."""
    if isinstance(points, np.ndarray):
        return points
    if hasattr(points, "columns") and hasattr(points, "values"):          # DataFrame / GeoDataFrame
        cols = list(points.columns)
        first = points[cols[0]]
        if len(cols) == 2 and hasattr(first, "x") and hasattr(first, "y"):
            return np.column_stack([np.asarray(first.x), np.asarray(first.y), np.asarray(points[cols[1]])])
        return points.values
    return np.array(points)


def interpolation_points(points):
    """``[x, y]`` rows as a 2-D ndarray (core/data_models/points.py:133-168)."""
    if hasattr(points, "x") and hasattr(points, "y") and not isinstance(points, np.ndarray):
        xs, ys = np.asarray(points.x), np.asarray(points.y)             # GeoSeries or a single Point
        return np.column_stack((xs, ys)) if xs.ndim else np.array([[float(xs), float(ys)]])
    if hasattr(points, "columns") and hasattr(points, "values"):
        return points.values
    arr = points if isinstance(points, np.ndarray) else np.array(points)      # (inputs are borrowed: no copy of a large array)
    if arr.ndim == 1:
        arr = arr.reshape(1, -1)
    return arr


def validate_bins(step_size, max_range, custom_bins):
    """core/validators/experimental_semivariance.py:51-73."""
    if step_size is None and max_range is None:
        if custom_bins is None:
            raise ValueError('You must provide at least `step_size` and `max_range` or '
                             '`custom_bins` parameter.')


def validate_direction_and_tolerance(direction, tolerance):
    """core/validators/experimental_semivariance.py:76-115."""
    if direction is not None and tolerance is not None:
        if abs(direction) > 360:
            raise ValueError('Provided direction must be between -360 to 360 degrees.')
        if tolerance <= 0 or tolerance > 1:
            raise ValueError('Provided tolerance should be larger than 0 '
                             '(a straight line) and smaller or equal to 1 (a circle).')
    elif direction is None and tolerance is not None:
        raise ValueError('Tolerance is set but direction is not set')
    elif direction is not None and tolerance is None:
        raise ValueError('Direction is set but tolerance is not set')


def validate_semivariance_weights(points, weights):
    """core/validators/experimental_semivariance.py:118-153."""
    if weights is not None:
        len_p, len_w = len(points), len(weights)
        if len_p != len_w:
            raise IndexError(f'Weights array length must be the same as length of points '
                             f'array but it has {len_w} records and'
                             f' points array has {len_p} records')
        if any([x == 0 for x in weights]):
            raise ValueError('One or more of weights in dataset is set to 0, '
                             'remove point with a zero weight from a dataset, and '
                             'do not pass 0 as a weight.')


def get_lags(step_size, max_range, custom_bins):
    """semivariogram/lags/lags.py:6-31 — np.arange(step, max_range, step); a leading 0 bin is dropped."""
    if custom_bins is not None:
        if custom_bins[0] == 0:
            return np.array(custom_bins)[1:]
        return np.array(custom_bins)
    return np.arange(step_size, max_range, step_size)


def define_whitening_matrix(theta, minor_axis_size):
    """distance/angular.py:221-250 + _rotation_matrix :620-638, with the same library calls, so the
    four doubles handed to the kernel are the reference's own."""
    from scipy.linalg import fractional_matrix_power
    p_lambda = np.array([[1, 0], [0, minor_axis_size]])
    frac_p_lambda = fractional_matrix_power(p_lambda, -0.5)
    t = np.radians(theta)
    rot = np.array([[np.cos(t), np.sin(t)], [-np.sin(t), np.cos(t)]])
    return np.matmul(frac_p_lambda, rot)


def ellipse_lower_limits(lags):
    """lower = lag - step_size with step_size = current - previous
    (semivariogram/experimental/functions/directional.py:428-432, distance/angular.py:672-673)."""
    lags = np.asarray(lags, dtype=np.float64)
    prev = np.concatenate([[0.0], lags[:-1]])
    return lags - (lags - prev)
