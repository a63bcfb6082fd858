"""TEST INFRASTRUCTURE ONLY — ctypes front-end of oracle/libpyinterp_oracle.so.

Mirrors the reference's call shapes (``calculate_semivariance``-style ``[lag, value, n]``
tables; ``ordinary_kriging``-style ``[zhat, sigma2, x, y]`` rows) so the parity tests read
like the reference's own tests.  Only tests/, ``__graft_entry__.smoke()`` and the CPU legs of
``bench.py`` import this module.  The host-side pieces below (lags, whitening matrix) restate

  * semivariogram/lags/lags.py:6-31           get_lags
  * distance/angular.py:221-250, :620-638     define_whitening_matrix, _rotation_matrix

with numpy/scipy exactly as the reference calls them.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpyinterp_oracle.so")

MODEL_IDS = {"circular": 0, "cubic": 1, "exponential": 2, "gaussian": 3,
             "linear": 4, "power": 5, "spherical": 6}

_lib = None


def build(force=False):
    """Compile the C restatement (gcc, a second or two)."""
    src = os.path.join(_HERE, "pyinterp_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int64)
        L.oracle_num_threads.restype = ctypes.c_int
        L.oracle_set_threads.argtypes = [ctypes.c_int]
        L.oracle_variogram_omni_rows.argtypes = [dp, ctypes.c_int64, dp, ctypes.c_int,
                                                 ctypes.c_int64, ctypes.c_int64, dp, dp, dp, ip]
        L.oracle_variogram_ellipse_rows.argtypes = [dp, ctypes.c_int64, dp, dp, ctypes.c_int,
                                                    dp, ctypes.c_int, ctypes.c_int64, ctypes.c_int64,
                                                    dp, dp, dp, ip]
        L.oracle_finalize.argtypes = [dp, dp, dp, ip, ctypes.c_int, dp, dp, dp]
        L.oracle_gamma.argtypes = [ctypes.c_int] + [ctypes.c_double] * 4
        L.oracle_gamma.restype = ctypes.c_double
        L.oracle_krige.argtypes = [dp, ctypes.c_int64, dp, ctypes.c_int64, ctypes.c_int,
                                   ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                   ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                   dp, ip, ctypes.POINTER(ctypes.c_uint8)]
        _lib = L
    return _lib


def num_threads():
    return lib().oracle_num_threads()


def set_threads(n):
    """Use n OpenMP threads (torchrun exports OMP_NUM_THREADS=1; the CPU baseline wants every core)."""
    lib().oracle_set_threads(int(n))


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


def get_lags(step_size, max_range, custom_bins=None):
    """semivariogram/lags/lags.py:6-31."""
    if custom_bins is not None:
        if custom_bins[0] == 0:
            return np.array(custom_bins)[1:]
        return np.array(custom_bins)
    return np.arange(step_size, max_range, step_size)


def whitening_matrix(theta, minor_axis_size):
    """distance/angular.py:221-250 + :620-638, same library calls as the reference."""
    from scipy.linalg import fractional_matrix_power
    p_lambda = np.array([[1, 0], [0, minor_axis_size]])
    frac = fractional_matrix_power(p_lambda, -0.5)
    t = np.radians(theta)
    rot = np.array([[np.cos(t), np.sin(t)], [-np.sin(t), np.cos(t)]])
    return np.matmul(frac, rot)


def ellipse_lower_limits(lags):
    """lower = lag - step_size with step_size = cur - prev (directional.py:432, angular.py:672-673)."""
    lags = np.asarray(lags, dtype=np.float64)
    prev = np.concatenate([[0.0], lags[:-1]])
    return lags - (lags - prev)


def variogram_sums(points, lags, directions=None, tolerance=None, rows=None):
    """Unordered-pair sums per lag.  Returns dict(sum_sq, sum_prod, sum_val, count), each
    ``[L]`` (omni) or ``[D, L]`` (ellipse, ``directions`` a sequence of angles in degrees)."""
    xyz = np.ascontiguousarray(np.asarray(points, dtype=np.float64))
    lags = np.ascontiguousarray(np.asarray(lags, dtype=np.float64))
    n, L = xyz.shape[0], lags.shape[0]
    r0, r1 = (0, n) if rows is None else rows
    if directions is None:
        shape = (L,)
    else:
        shape = (len(directions), L)
    ssq, spr, sva = (np.zeros(shape) for _ in range(3))
    cnt = np.zeros(shape, dtype=np.int64)
    if directions is None:
        lib().oracle_variogram_omni_rows(_dp(xyz), n, _dp(lags), L, r0, r1,
                                         _dp(ssq), _dp(spr), _dp(sva), _ip(cnt))
    else:
        W = np.ascontiguousarray(np.stack([whitening_matrix(d, tolerance) for d in directions])
                                 .reshape(len(directions), 4))
        lower = np.ascontiguousarray(ellipse_lower_limits(lags))
        lib().oracle_variogram_ellipse_rows(_dp(xyz), n, _dp(lags), _dp(lower), L, _dp(W),
                                            len(directions), r0, r1,
                                            _dp(ssq), _dp(spr), _dp(sva), _ip(cnt))
    return dict(sum_sq=ssq, sum_prod=spr, sum_val=sva, count=cnt)


def finalize(sums):
    """-> (semivariance, covariance, n_pairs) with the reference's empty-lag conventions."""
    ssq, spr, sva, cnt = (np.ascontiguousarray(sums[k]) for k in ("sum_sq", "sum_prod", "sum_val", "count"))
    shape = cnt.shape
    L = shape[-1]
    sem, cov, npairs = (np.zeros(shape) for _ in range(3))
    for idx in np.ndindex(*shape[:-1]):
        s, c, p = np.zeros(L), np.zeros(L), np.zeros(L)
        lib().oracle_finalize(_dp(np.ascontiguousarray(ssq[idx])), _dp(np.ascontiguousarray(spr[idx])),
                              _dp(np.ascontiguousarray(sva[idx])), _ip(np.ascontiguousarray(cnt[idx])),
                              L, _dp(s), _dp(c), _dp(p))
        sem[idx], cov[idx], npairs[idx] = s, c, p
    return sem, cov, npairs


def calculate_semivariance(points, step_size=None, max_range=None, direction=None, tolerance=None,
                           custom_bins=None):
    """``[lag, semivariance, n_pairs]`` like experimental_semivariogram.py:19-229 (omni or 'e')."""
    lags = get_lags(step_size, max_range, custom_bins)
    dirs = None if direction is None else [direction]
    sem, _, npairs = finalize(variogram_sums(points, lags, dirs, tolerance))
    if dirs is not None:
        sem, npairs = sem[0], npairs[0]
    return np.column_stack([lags, sem, npairs])


def calculate_covariance(points, step_size=None, max_range=None, direction=None, tolerance=None,
                         custom_bins=None):
    """``[lag, covariance, n_pairs]`` like experimental_covariogram.py:17-172."""
    lags = get_lags(step_size, max_range, custom_bins)
    dirs = None if direction is None else [direction]
    _, cov, npairs = finalize(variogram_sums(points, lags, dirs, tolerance))
    if dirs is not None:
        cov, npairs = cov[0], npairs[0]
    return np.column_stack([lags, cov, npairs])


def gamma(model, h, nugget, sill, rang):
    mid = MODEL_IDS[model]
    h = np.asarray(h, dtype=np.float64)
    return np.array([lib().oracle_gamma(mid, float(v), nugget, sill, rang) for v in h.ravel()]).reshape(h.shape)


def krige(known, unknown, model, nugget, sill, rang, no_neighbors, mode="ok", process_mean=0.0):
    """-> (out[M,4] = [zhat, sigma2, x, y], nbr_idx[M,k], status[M]) — ordinary.py:134-221 /
    simple.py:129-223 batch semantics."""
    known = np.ascontiguousarray(np.asarray(known, dtype=np.float64))
    unknown = np.ascontiguousarray(np.asarray(unknown, dtype=np.float64).reshape(-1, 2))
    n, m, k = known.shape[0], unknown.shape[0], int(no_neighbors)
    out = np.zeros((m, 4))
    idx = np.zeros((m, k), dtype=np.int64)
    status = np.zeros(m, dtype=np.uint8)
    lib().oracle_krige(_dp(known), n, _dp(unknown), m, MODEL_IDS[model], float(nugget), float(sill),
                       float(rang), k, 0 if mode == "ok" else 1, float(process_mean),
                       _dp(out), _ip(idx), status.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    return out, idx, status


def krige_with_fallback(known, unknown, model, nugget, sill, rang, no_neighbors, mode="ok", process_mean=0.0,
                        allow_lsa=True):
    """``krige`` plus the reference's singular-system fallbacks (kriging/utils/point_kriging_solve.py:134-147):
    a system whose LU meets a zero pivot gets zero weights when mean(A) == 0 or mean(b) == 0 (status 5), the
    ``np.linalg.lstsq(A, b, rcond=None)`` solution when ``allow_lsa`` (status 4), else stays status 1.
    Plain numpy per singular row — small cases only."""
    known = np.ascontiguousarray(np.asarray(known, dtype=np.float64))
    out, idx, status = krige(known, unknown, model, nugget, sill, rang, no_neighbors, mode, process_mean)
    for u in np.nonzero(status == 1)[0]:
        nb = known[idx[u][idx[u] >= 0]]
        kk = len(nb)
        dx = nb[:, None, 0] - nb[None, :, 0]
        dy = nb[:, None, 1] - nb[None, :, 1]
        big = gamma(model, np.sqrt(dx * dx + dy * dy), nugget, sill, rang)
        ux, uy = out[u, 2], out[u, 3]
        ddx, ddy = ux - nb[:, 0], uy - nb[:, 1]
        kvec = gamma(model, np.sqrt(ddx * ddx + ddy * ddy), nugget, sill, rang)
        if mode == "ok":                                         # ordinary.py:108-115
            a = np.zeros((kk + 1, kk + 1))
            a[:kk, :kk] = big
            a[kk, :kk] = 1.0
            a[:kk, kk] = 1.0
            b = np.append(kvec, 1.0)
        else:
            a, b = big, kvec
        if np.mean(a) == 0 or np.mean(b) == 0:
            w = np.zeros(len(b))
            status[u] = 5
        elif allow_lsa:
            w = np.linalg.lstsq(a, b, rcond=None)[0]
            status[u] = 4
        else:
            continue
        z = nb[:, 2]
        zhat = z.dot(w[:kk]) if mode == "ok" else (z - process_mean).dot(w[:kk]) + process_mean
        sigma = w.dot(b)
        out[u, 0] = zhat
        out[u, 1] = np.nan if sigma < 0 else sigma
    return out, idx, status
