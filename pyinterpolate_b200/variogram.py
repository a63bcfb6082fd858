"""Experimental variogram — the reference's public calls, backed by the CUDA pair kernels.

Mirrors (file:line relative to /root/reference/src/pyinterpolate/):
  * semivariogram/experimental/experimental_semivariogram.py:19-229   calculate_semivariance
  * semivariogram/experimental/experimental_covariogram.py:17-172     calculate_covariance
  * semivariogram/experimental/classes/experimental_variogram.py:16-496  ExperimentalVariogram,
                                                                      build_experimental_variogram
  * semivariogram/experimental/classes/directional_variogram.py:8-209 DirectionalVariogram

One kernel pass yields semivariance, covariance and pair counts together, so the class computes
both curves from a single launch where the reference runs its whole pipeline twice.
The point cloud (``as_cloud=True`` / ``point_cloud_semivariance`` / ``VariogramCloud``) comes from a
second kernel that writes (z_i - z_j)^2 per ordered pair in the reference's order.
Directional variograms support both of the reference's selection methods: the ellipse ('e', tiled fast
kernels) and the triangle ('t', the reference's default; exact reproduction of its floating-point masks).
"""
from collections import OrderedDict

import numpy as np

from . import _lib, core


def _finalize(sums, directional):
    """Unordered-pair sums -> the reference's per-lag numbers (ordered pairs):
    semivariance = mean((v0-vh)^2)/2 (functions/semivariance.py:22-23),
    covariance = mean(v0*vh) - mean(vh)^2 (functions/covariance.py:22-26), n = len(v0) = 2*count.
    Empty omnidirectional lag: (0, 0) for lag index 0 else (nan, nan) (functions/general.py:283-287)."""
    cnt = np.asarray(sums["count"], dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        sem = np.asarray(sums["sum_sq"]) / cnt / 2.0
        mean = np.asarray(sums["sum_val"]) / (2.0 * cnt)
        cov = np.asarray(sums["sum_prod"]) / cnt - mean * mean
    npairs = 2.0 * cnt
    empty = cnt == 0
    if empty.any():
        sem = np.where(empty, np.nan, sem)
        cov = np.where(empty, np.nan, cov)
        npairs = np.where(empty, np.nan, npairs)
        if not directional and empty[..., 0].any():
            sem[..., 0] = np.where(empty[..., 0], 0.0, sem[..., 0])
            cov[..., 0] = np.where(empty[..., 0], 0.0, cov[..., 0])
            npairs[..., 0] = np.where(empty[..., 0], 0.0, npairs[..., 0])
    return sem, cov, npairs


def _check_method(direction, method, custom_weights):
    if custom_weights is not None:
        return                              # the weighted directional path always selects with the ellipse
    if direction is not None and method not in ("e", "ellipse", "t", "triangle"):
        raise ValueError("dir_neighbors_selection_method must be 't' / 'triangle' or 'e' / 'ellipse'")


def _triangle_lag_table(lags, direction, tolerance):
    """Per-lag triangle offsets exactly as distance/angular.py:283-339 + :599-617 compute them
    (``distance * np.cos(angle)`` with numpy scalars), handed to the kernel as nine doubles per lag."""
    angle = np.radians(direction)
    rot_90 = np.pi / 2
    rows = []
    for h in lags:
        base_width = h * tolerance
        rows.append([float(h),
                     h * np.cos(float(angle)), h * np.sin(float(angle)),
                     -h * np.cos(float(angle)), -h * np.sin(float(angle)),
                     base_width * np.cos(angle + rot_90), base_width * np.sin(angle + rot_90),
                     base_width * np.cos(angle - rot_90), base_width * np.sin(angle - rot_90)])
    return np.array(rows, dtype=np.float64), float(np.cos(angle)), float(np.sin(angle))


def _triangle_curves(xyz, lags, direction, tolerance):
    """(semivariance, covariance, n_pairs) with the triangle selection: functions/directional.py:137-210.
    The sums run over ORDERED pairs (centre, neighbour) because the reference's masks are evaluated per centre."""
    table, c, s_ = _triangle_lag_table(lags, direction, tolerance)
    sums, cnt = _lib.variogram_triangle_host(xyz, table, c, s_, tolerance)
    n = cnt.astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        sem = sums[0] / n / 2.0                                  # mean((v0 - vh)^2) / 2
        cov = sums[1] / n - (sums[2] / n) ** 2                   # mean(v0 vh) - mean(vh)^2
    return sem, cov, n


def experimental_curves(ds, step_size=None, max_range=None, direction=None, tolerance=None,
                        dir_neighbors_selection_method="t", custom_bins=None, custom_weights=None,
                        directions=None, shard=(0, 1), reduce_fn=None):
    """Shared driver: returns (lags, semivariances, covariances, n_pairs).

    ``directions`` (a sequence of angles) evaluates several ellipse directions in one library call and
    returns ``[D, L]`` arrays.  ``shard`` / ``reduce_fn`` let a multi-GPU caller evaluate one slice of
    the tile-pair list per rank and sum the partial results (see ``distributed.py``)."""
    points = core.variogram_points(ds)
    core.validate_bins(step_size, max_range, custom_bins)
    core.validate_semivariance_weights(points, custom_weights)
    if directions is None:
        core.validate_direction_and_tolerance(direction, tolerance)
    else:
        for angle in directions:
            core.validate_direction_and_tolerance(angle, tolerance)
    lags = core.get_lags(step_size, max_range, custom_bins)
    directional = (direction is not None and tolerance is not None) or directions is not None
    _check_method(direction if directions is None else 0, dir_neighbors_selection_method, custom_weights)
    lags64 = np.ascontiguousarray(lags, dtype=np.float64)
    xyz = np.ascontiguousarray(points, dtype=np.float64)
    if xyz.ndim != 2 or xyz.shape[1] < 3:
        raise ValueError("points must be rows of [x, y, value]")
    xyz = np.ascontiguousarray(xyz[:, [0, 1, xyz.shape[1] - 1]]) if xyz.shape[1] != 3 else xyz
    triangle = directional and dir_neighbors_selection_method in ("t", "triangle")
    if triangle:
        if shard != (0, 1):
            raise NotImplementedError("the triangle method is not sharded across ranks")
        angles = [direction] if directions is None else list(directions)
        curves = [_triangle_curves(xyz, lags, a, tolerance) for a in angles]
        sem, cov, npairs = (np.stack([c[i] for c in curves]) for i in range(3))
        empty = npairs == 0
        sem, cov = np.where(empty, 0.0, sem), np.where(empty, 0.0, cov)
    else:
        if directional:
            angles = [direction] if directions is None else list(directions)
            W = np.stack([core.define_whitening_matrix(a, tolerance) for a in angles]).reshape(len(angles), 4)
            sums = _lib.variogram_host(xyz, lags64, lower=core.ellipse_lower_limits(lags64), W=W, shard=shard)
        else:
            sums = _lib.variogram_host(xyz, lags64, shard=shard)
        if reduce_fn is not None:
            sums = reduce_fn(sums)
        sem, cov, npairs = _finalize(sums, directional)
    if directional:
        if directions is None:
            sem, cov, npairs = sem[0], cov[0], npairs[0]
        bad = np.argwhere(~(np.atleast_2d(npairs) > 0))
        if bad.size:
            # functions/directional.py:68-74 (idx 0 with lags[0] == 0 is the only tolerated empty lag)
            idx = int(bad[0][-1])
            if not (idx == 0 and lags[0] == 0):
                raise RuntimeError(f'There are no neighbors for a lag {lags[idx]},'
                                   f' the process has been stopped.')
    return lags, sem, cov, npairs


def weighted_semivariance(ds, custom_weights, step_size=None, max_range=None, direction=None, tolerance=None,
                          custom_bins=None):
    """``[lag, weighted semivariance, number of point pairs]`` for ``custom_weights``:
    omnidirectional — semivariogram/weights/experimental/weighting.py:4-54 through
    functions/general.py:292-301; directional — functions/directional.py:276-392 (ellipse selection
    whatever ``dir_neighbors_selection_method`` says, like the reference)."""
    points = core.variogram_points(ds)
    core.validate_bins(step_size, max_range, custom_bins)
    core.validate_semivariance_weights(points, custom_weights)
    core.validate_direction_and_tolerance(direction, tolerance)
    lags = core.get_lags(step_size, max_range, custom_bins)
    lags64 = np.ascontiguousarray(lags, dtype=np.float64)
    xyz = np.ascontiguousarray(points, dtype=np.float64)
    w = np.ascontiguousarray(custom_weights, dtype=np.float64)
    directional = direction is not None and tolerance is not None
    if directional:
        W = core.define_whitening_matrix(direction, tolerance).reshape(4)
        sums, cnt = _lib.variogram_weighted_host(xyz, w, lags64, lower=core.ellipse_lower_limits(lags64), W=W)
    else:
        sums, cnt = _lib.variogram_weighted_host(xyz, w, lags64)
    s1, s2, s3, s4 = sums
    out = np.zeros((len(lags), 3))
    out[:, 0] = lags
    for b in range(len(lags)):
        if cnt[b] == 0:
            if directional:
                if b == 0:
                    out[b] = [lags[b], 0, 0]
                    continue
                raise RuntimeError(f'There are no neighbors for a lag {lags[b]}, the process has been stopped.')
            out[b, 1:] = (0, 0) if b == 0 else (np.nan, np.nan)        # functions/general.py:283-287
            continue
        if directional:
            # gamma = sum w_h (dz^2 - avg) / sum w_h, avg = weighted mean of the neighbour values (:377-383)
            out[b, 1] = 0.5 * (s1[b] / s2[b] - s3[b] / s4[b])
        else:
            # sum(w sem - m') / (2 sum w) over ordered pairs, m' = mean of the weighted values (weighting.py:40-50)
            value = (2.0 * s1[b] - s3[b]) / (4.0 * s2[b])
            out[b, 1] = 0 if value < 0 else value
        out[b, 2] = 2 * cnt[b]
    return out


def calculate_semivariance(ds, step_size=None, max_range=None, direction=None, tolerance=None,
                           dir_neighbors_selection_method="t", custom_bins=None, custom_weights=None):
    """``[lag, semivariance, number of point pairs]`` — experimental_semivariogram.py:19-229."""
    if custom_weights is not None:
        return weighted_semivariance(ds, custom_weights, step_size, max_range, direction, tolerance, custom_bins)
    lags, sem, _, npairs = experimental_curves(ds, step_size, max_range, direction, tolerance,
                                               dir_neighbors_selection_method, custom_bins, custom_weights)
    return np.column_stack([lags, sem, npairs])


def calculate_covariance(ds, step_size=None, max_range=None, direction=None, tolerance=None,
                         dir_neighbors_selection_method="t", custom_bins=None):
    """``[lag, covariance, number of point pairs]`` — experimental_covariogram.py:17-172."""
    lags, _, cov, npairs = experimental_curves(ds, step_size, max_range, direction, tolerance,
                                               dir_neighbors_selection_method, custom_bins)
    return np.column_stack([lags, cov, npairs])


def point_cloud_semivariance(ds, step_size=None, max_range=None, direction=None, tolerance=None,
                             dir_neighbors_selection_method="t", custom_bins=None, custom_weights=None):
    """``{lag: [(z_i - z_j)^2 for every ordered pair in the lag]}`` — experimental_semivariogram.py
    (point_cloud_semivariance) -> functions/general.py:13-75 / functions/directional.py:83-134.
    Values come in the reference's order (i ascending, then j ascending); an empty lag maps to ``[]``.
    ``custom_weights`` are ignored for clouds, like in the reference."""
    points = core.variogram_points(ds)
    core.validate_bins(step_size, max_range, custom_bins)
    core.validate_semivariance_weights(points, custom_weights)
    core.validate_direction_and_tolerance(direction, tolerance)
    lags = core.get_lags(step_size, max_range, custom_bins)
    lags64 = np.ascontiguousarray(lags, dtype=np.float64)
    xyz = np.ascontiguousarray(points, dtype=np.float64)
    if direction is not None and tolerance is not None:
        _check_method(direction, dir_neighbors_selection_method, None)
        if dir_neighbors_selection_method in ("t", "triangle"):
            # the reference's default: from_triangle_cloud (experimental_semivariogram.py:282-294, directional.py:213-273)
            table, c, s_ = _triangle_lag_table(lags, direction, tolerance)
            clouds = _lib.variogram_triangle_cloud_host(xyz, table, c, s_, tolerance)
        else:
            W = core.define_whitening_matrix(direction, tolerance).reshape(4)
            clouds = _lib.variogram_cloud_host(xyz, lags64, lower=core.ellipse_lower_limits(lags64), W=W)
    else:
        clouds = _lib.variogram_cloud_host(xyz, lags64)
    out = OrderedDict()
    for lag, values in zip(lags, clouds):
        out[lag] = values if len(values) else []
    return out


class ExperimentalVariogram:
    """classes/experimental_variogram.py:16-402: attributes ``lags``, ``semivariances``,
    ``covariances``, ``points_per_lag``, ``variance``, ``direction``, ``tolerance``, ``step_size``,
    ``max_range``, ``custom_weights``, ``model`` (a plain dict of those)."""

    def __init__(self, ds, step_size=None, max_range=None, direction=None, tolerance=None,
                 dir_neighbors_selection_method="t", custom_bins=None, custom_weights=None,
                 is_semivariance=True, is_covariance=True, as_cloud=False):
        self.ds = core.variogram_points(ds)
        self.lags = None if custom_bins is None else custom_bins
        self.points_per_lag = None
        self.semivariances = None
        self.covariances = None
        self.point_cloud_semivariances = None
        self.variance = np.var(self.ds[:, -1])                       # :155
        self.step_size = step_size
        self.max_range = max_range
        self.custom_weights = custom_weights
        self.direction = direction
        self.tolerance = tolerance
        self.method = dir_neighbors_selection_method
        self.as_cloud = as_cloud
        if custom_weights is not None and is_semivariance:
            # weighted semivariance first (classes/experimental_variogram.py:284-307); the covariance call
            # below takes no weights in the reference either (:240-262)
            wtab = weighted_semivariance(self.ds, custom_weights, step_size, max_range, direction, tolerance, self.lags)
            self.lags, self.semivariances, self.points_per_lag = wtab[:, 0], wtab[:, 1], wtab[:, 2]
            if is_covariance:
                _, _, cov, _ = experimental_curves(self.ds, step_size, max_range, direction, tolerance,
                                                   dir_neighbors_selection_method, custom_bins)
                self.covariances = cov
        elif is_semivariance or is_covariance:
            lags, sem, cov, npairs = experimental_curves(
                self.ds, step_size, max_range, direction, tolerance, dir_neighbors_selection_method,
                self.lags, None)
            self.lags = lags
            self.points_per_lag = npairs
            if is_semivariance:
                self.semivariances = sem
            if is_covariance:
                self.covariances = cov
        if as_cloud:
            self.point_cloud_semivariances = point_cloud_semivariance(
                self.ds, step_size, max_range, direction, tolerance, dir_neighbors_selection_method,
                self.lags, custom_weights)
            if self.lags is None:
                self.lags = core.get_lags(step_size, max_range, custom_bins)
        self.model = self.get_model_params()

    def get_model_params(self):
        return {
            "lags": self.lags, "points_per_lag": self.points_per_lag,
            "semivariances": self.semivariances, "covariances": self.covariances,
            "variance": self.variance, "direction": self.direction, "tolerance": self.tolerance,
            "max_range": self.max_range, "step_size": self.step_size,
            "custom_weights": self.custom_weights,
        }

    def __repr__(self):
        return f"ExperimentalVariogram(lags={None if self.lags is None else len(self.lags)}, direction={self.direction})"

    def __str__(self):
        if self.lags is None:
            return "Empty object"
        rows = ["lag | semivariance | covariance | var_cov_diff"]
        for i, lag in enumerate(self.lags):
            s = self.semivariances[i] if self.semivariances is not None else np.nan
            c = self.covariances[i] if self.covariances is not None else np.nan
            rows.append(f"{lag} | {s} | {c} | {self.variance - c}")
        return "\n".join(rows)


def build_experimental_variogram(ds, step_size=None, max_range=None, direction=None, tolerance=None,
                                 dir_neighbors_selection_method="t", custom_bins=None, custom_weights=None,
                                 is_semivariance=True, is_covariance=True, as_cloud=False):
    """classes/experimental_variogram.py:405-496."""
    return ExperimentalVariogram(ds, step_size, max_range, direction, tolerance, dir_neighbors_selection_method,
                                 custom_bins, custom_weights, is_semivariance, is_covariance, as_cloud)


class DirectionalVariogram:
    """classes/directional_variogram.py:8-209 — ISO + NS (90), WE (0), NE-SW (45), NW-SE (135).
    The four directional variograms come from ONE library call (one sort, four kernel passes)."""

    def __init__(self, ds, step_size, max_range, tolerance=0.2, dir_neighbors_selection_method="t",
                 custom_weights=None, custom_bins=None):
        self.ds = ds
        self.custom_bins = custom_bins
        self.step_size = step_size
        self.max_range = max_range
        self.tolerance = tolerance
        self.custom_weights = custom_weights
        self.dir_neighbors_selection_method = dir_neighbors_selection_method
        self.possible_variograms = ['ISO', 'NS', 'WE', 'NE-SW', 'NW-SE']
        self.directions = {'NS': 90, 'WE': 0, 'NE-SW': 45, 'NW-SE': 135}
        self.directional_variograms = {}
        self._build_experimental_variograms()

    def _build_experimental_variograms(self):
        _check_method(0, self.dir_neighbors_selection_method, self.custom_weights)
        if self.custom_weights is not None:
            # the reference hands custom_weights to all five ExperimentalVariograms (directional_variogram.py:126-143):
            # the weighted path, one variogram at a time
            self.directional_variograms['ISO'] = ExperimentalVariogram(
                self.ds, self.step_size, self.max_range, custom_bins=self.custom_bins, custom_weights=self.custom_weights)
            for name, angle in self.directions.items():
                self.directional_variograms[name] = ExperimentalVariogram(
                    self.ds, self.step_size, self.max_range, direction=angle, tolerance=self.tolerance,
                    dir_neighbors_selection_method=self.dir_neighbors_selection_method, custom_bins=self.custom_bins,
                    custom_weights=self.custom_weights)
            return
        self.directional_variograms['ISO'] = ExperimentalVariogram(
            self.ds, self.step_size, self.max_range, custom_bins=self.custom_bins)
        names = list(self.directions)
        lags, sem, cov, npairs = experimental_curves(
            self.ds, self.step_size, self.max_range, tolerance=self.tolerance, direction=None,
            dir_neighbors_selection_method=self.dir_neighbors_selection_method, custom_bins=self.custom_bins,
            directions=[self.directions[k] for k in names])
        for t, name in enumerate(names):
            ev = ExperimentalVariogram.__new__(ExperimentalVariogram)
            iso = self.directional_variograms['ISO']
            ev.__dict__.update(iso.__dict__)
            ev.lags, ev.semivariances, ev.covariances, ev.points_per_lag = lags, sem[t], cov[t], npairs[t]
            ev.direction, ev.tolerance, ev.method = self.directions[name], self.tolerance, self.dir_neighbors_selection_method
            ev.model = ev.get_model_params()
            self.directional_variograms[name] = ev

    def get(self, direction=None):
        if direction is None:
            return self.directional_variograms
        if direction in self.possible_variograms:
            return self.directional_variograms[direction]
        raise KeyError(f'Given direction is not possible to retrieve, pass one direction from a '
                       f'possible_variograms: {self.possible_variograms} or leave None to get a dictionary with all '
                       f'possible variograms.')


class VariogramCloud:
    """classes/variogram_cloud.py:82-504 (data side: ``semivariances``, ``lags``, ``points_per_lag``,
    ``average_semivariance``, ``describe``, ``experimental_semivariances``, ``remove_outliers``); plots stay
    with the reference."""

    def __init__(self, ds, step_size=None, max_range=None, direction=None, tolerance=None,
                 dir_neighbors_selection_method='t', custom_bins=None, custom_weights=None):
        self._experimental_variogram = ExperimentalVariogram(
            ds=ds, step_size=step_size, max_range=max_range, direction=direction, tolerance=tolerance,
            dir_neighbors_selection_method=dir_neighbors_selection_method, custom_bins=custom_bins,
            custom_weights=custom_weights, is_semivariance=True, is_covariance=False, as_cloud=True)
        self.semivariances = self._experimental_variogram.point_cloud_semivariances
        self.lags = self._experimental_variogram.lags
        self.direction = direction
        self.tolerance = tolerance
        self.points_per_lag = [len(self.semivariances[lag]) for lag in self.lags]

    def experimental_semivariances(self):
        return self._experimental_variogram

    def average_semivariance(self):
        """``[lag, mean(cloud) / 2, count]`` per lag (variogram_cloud.py:213-232)."""
        return np.array([[lag, np.mean(self.semivariances[lag]) / 2, len(self.semivariances[lag])]
                         for lag in self.lags])

    def describe(self, as_dataframe=False):
        """variogram_cloud.py:234-294."""
        from scipy.stats import kurtosis, skew
        stats = {}
        for lag in self.lags:
            data = self.semivariances[lag]
            stats[lag] = {'count': len(data), 'mean': np.mean(data) / 2, 'std': np.std(data), 'min': np.min(data),
                          'max': np.max(data), '25%': np.quantile(data, q=0.25), 'median': np.median(data),
                          '75%': np.quantile(data, q=0.75), 'skewness': skew(data), 'kurtosis': kurtosis(data),
                          'lag': lag}
        if as_dataframe:
            import pandas as pd
            return pd.DataFrame(stats)
        return stats

    def remove_outliers(self, method='zscore', z_lower_limit=-3, z_upper_limit=3, iqr_lower_limit=1.5,
                        iqr_upper_limit=1.5, inplace=False):
        """variogram_cloud.py:334-397 with the detectors of transform/statistical.py:14-95: per lag, drop the cloud
        values whose z-score leaves ``[z_lower_limit, z_upper_limit]`` or that lie more than ``iqr_*_limit`` standard
        deviations outside the quartiles."""
        if method == 'zscore':
            if z_lower_limit >= 0:
                raise ValueError('The parameter z_lower must be a float lesser than zero.')
            if z_upper_limit <= 0:
                raise ValueError('The parameter z_upper must be a float greater than zero.')
        elif method == 'iqr':
            if iqr_upper_limit < 0 or iqr_lower_limit < 0:
                raise ValueError('Parameters iqr_lower and iqr_upper must be floats greater or equal to 0.')
        else:
            raise KeyError(f'Given detection method: {method} is not available. '
                           f'Available methods are "zscore" and "iqr".')
        cleaned = {}
        for lag, values in self.semivariances.items():
            values = np.asarray(values)
            if method == 'zscore':
                with np.errstate(invalid='ignore', divide='ignore'):
                    score = (values - values.mean()) / values.std()          # scipy.stats.zscore, ddof = 0
                bad = (score > z_upper_limit) | (score < z_lower_limit)
            else:
                spread = np.std(values)
                bad = ((values > np.quantile(values, q=0.75) + spread * iqr_upper_limit)
                       | (values < np.quantile(values, q=0.25) - spread * iqr_lower_limit))
            cleaned[lag] = values[~bad]
        if inplace:
            self.semivariances = cleaned
            return None
        import copy
        other = copy.copy(self)
        other.semivariances = cleaned
        return other
