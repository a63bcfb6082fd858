"""ctypes binding of libpyinterp_b200.so (include/pyinterp_b200.h).

There is no CPU fallback: if the CUDA library is missing or a call fails, an exception is
raised.  The library is built in-tree by ``__graft_entry__.build()`` / ``csrc/Makefile``.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PIB_LIB_PATH: load another build of the same library (kernel A/B measurements)
LIB_PATH = os.environ.get("PIB_LIB_PATH") or os.path.join(_HERE, "libpyinterp_b200.so")

MODEL_IDS = {"circular": 0, "cubic": 1, "exponential": 2, "gaussian": 3,
             "linear": 4, "power": 5, "spherical": 6}
KRIGE_ORDINARY, KRIGE_SIMPLE, KRIGE_BOTH = 0, 1, 2
STATUS_OK, STATUS_SINGULAR, STATUS_NEGATIVE_VARIANCE = 0, 1, 2

STATUS_TOO_MANY_NEIGHBORS = 3
STATUS_LSA, STATUS_ZERO_MATRIX = 4, 5

EXPORTS = ("pib_trim", "pib_last_kernel_name", "pib_krige_loo_host", "pib_krige_ex_device", "pib_krige_ex_host", "pib_last_error", "pib_version", "pib_device_sm_count", "pib_variogram_device",
           "pib_variogram_host", "pib_krige_device", "pib_krige_host", "pib_gamma_host",
           "pib_fp64_peak_tflops", "pib_last_kernel_ms", "pib_variogram_cloud_host",
           "pib_variogram_weighted_host", "pib_variogram_triangle_host", "pib_variogram_triangle_cloud_host",
           "pib_timing_accumulate", "pib_krige_column_host")

_lib = None


class PibError(RuntimeError):
    """A call into libpyinterp_b200.so failed."""


def load():
    """Load the shared library (no CUDA call is made until an entry point is used)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PibError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "or `make -C pyinterpolate_b200/csrc` — there is no CPU fallback")
    L = ctypes.CDLL(LIB_PATH)
    dp = ctypes.c_void_p
    i64, i32, f64 = ctypes.c_int64, ctypes.c_int, ctypes.c_double
    L.pib_last_error.restype = ctypes.c_char_p
    L.pib_version.restype = i32
    L.pib_device_sm_count.restype = i32
    vg_args = [dp, i64, dp, dp, i32, dp, i32, i32, i32, dp, dp, dp, dp, dp]
    L.pib_variogram_device.argtypes = vg_args + [ctypes.c_void_p]
    L.pib_variogram_host.argtypes = vg_args
    kr_args = [dp, i64, dp, i32, dp, dp, i64, i32, i32, f64, i32, dp, dp, dp]
    L.pib_krige_device.argtypes = kr_args + [ctypes.c_void_p]
    L.pib_krige_host.argtypes = kr_args
    kr_ex = kr_args[:11] + [f64, f64, f64, i32, i32] + kr_args[11:]
    L.pib_krige_ex_device.argtypes = kr_ex + [ctypes.c_void_p]
    L.pib_krige_ex_host.argtypes = kr_ex
    L.pib_krige_column_host.argtypes = kr_args[:11] + [f64, f64, f64, i32, i32, i32, dp, dp]
    L.pib_krige_loo_host.argtypes = [dp, i64, dp, i32, i32, f64, i32, f64, f64, f64, i32, i32, dp, dp]
    L.pib_gamma_host.argtypes = [i32, f64, f64, f64, dp, i64, dp]
    L.pib_fp64_peak_tflops.argtypes = [dp]
    L.pib_last_kernel_ms.argtypes = [i32, dp, dp]
    L.pib_timing_accumulate.argtypes = [i32]
    L.pib_last_kernel_name.argtypes = [i32, ctypes.c_char_p, i32]
    L.pib_variogram_cloud_host.argtypes = [dp, i64, dp, dp, i32, dp, i32, dp, dp]
    L.pib_variogram_weighted_host.argtypes = [dp, dp, i64, dp, dp, i32, dp, i32, dp, dp]
    L.pib_variogram_triangle_host.argtypes = [dp, i64, dp, i32, f64, f64, f64, dp, dp]
    L.pib_variogram_triangle_cloud_host.argtypes = [dp, i64, dp, i32, f64, f64, f64, dp, dp]
    for name in EXPORTS:
        getattr(L, name).restype = getattr(L, name).restype or i32
    L.pib_last_error.restype = ctypes.c_char_p
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        msg = load().pib_last_error()
        raise PibError(f"{what} failed (code {rc}): {msg.decode() if msg else 'unknown error'}")


def _host_ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def variogram_host(xyz, edges, lower=None, W=None, shard=(0, 1)):
    """Host-buffer call. Returns dict(sum_sq, sum_prod, sum_val, count, pairs_evaluated);
    arrays are ``[L]`` (omni) or ``[D, L]`` (ellipse)."""
    L = load()
    xyz, edges = _f64(xyz), _f64(edges)
    n, n_lags = xyz.shape[0], edges.shape[0]
    n_dirs = 0 if W is None else int(np.asarray(W).reshape(-1, 4).shape[0])
    Wc = None if W is None else _f64(np.asarray(W).reshape(-1, 4))
    lower_c = None if lower is None else _f64(lower)
    shape = (n_lags,) if n_dirs == 0 else (n_dirs, n_lags)
    ssq, spr, sva = (np.zeros(shape) for _ in range(3))
    cnt = np.zeros(shape, dtype=np.int64)
    evaluated = ctypes.c_int64(0)
    rc = L.pib_variogram_host(_host_ptr(xyz), n, _host_ptr(edges), _host_ptr(lower_c), n_lags, _host_ptr(Wc), n_dirs,
                              int(shard[0]), int(shard[1]), _host_ptr(ssq), _host_ptr(spr), _host_ptr(sva),
                              _host_ptr(cnt), ctypes.cast(ctypes.byref(evaluated), ctypes.c_void_p))
    check(rc, "pib_variogram_host")
    return dict(sum_sq=ssq, sum_prod=spr, sum_val=sva, count=cnt, pairs_evaluated=int(evaluated.value))


def variogram_weighted_host(xyz, weights, edges, lower=None, W=None):
    """Weighted pair sums (see the header).  Returns (sums[4, L], count[L]) over unordered pairs."""
    L = load()
    xyz, edges, weights = _f64(xyz), _f64(edges), _f64(weights)
    n, n_lags = xyz.shape[0], edges.shape[0]
    Wc = None if W is None else _f64(np.asarray(W).reshape(1, 4))
    lower_c = None if lower is None else _f64(lower)
    sums = np.zeros((4, n_lags))
    cnt = np.zeros(n_lags, dtype=np.int64)
    rc = L.pib_variogram_weighted_host(_host_ptr(xyz), _host_ptr(weights), n, _host_ptr(edges), _host_ptr(lower_c),
                                       n_lags, _host_ptr(Wc), 0 if W is None else 1, _host_ptr(sums), _host_ptr(cnt))
    check(rc, "pib_variogram_weighted_host")
    return sums, cnt


def variogram_triangle_host(xyz, tri_lags, cos_t, sin_t, tolerance):
    """Triangle-method pair sums over ordered pairs (see the header).  Returns (sums[3, L], count[L])."""
    L = load()
    xyz, tri_lags = _f64(xyz), _f64(tri_lags).reshape(-1, 9)
    n, n_lags = xyz.shape[0], tri_lags.shape[0]
    sums = np.zeros((3, n_lags))
    cnt = np.zeros(n_lags, dtype=np.int64)
    rc = L.pib_variogram_triangle_host(_host_ptr(xyz), n, _host_ptr(tri_lags), n_lags, float(cos_t), float(sin_t),
                                       float(tolerance), _host_ptr(sums), _host_ptr(cnt))
    check(rc, "pib_variogram_triangle_host")
    return sums, cnt


def variogram_triangle_cloud_host(xyz, tri_lags, cos_t, sin_t, tolerance):
    """Triangle-method point cloud: per lag the (z_p - z_q)^2 of the ordered pairs, in the reference's order
    (functions/directional.py:213-273).  Returns a list of float64 arrays, one per lag."""
    L = load()
    xyz, tri_lags = _f64(xyz), _f64(tri_lags).reshape(-1, 9)
    n, n_lags = xyz.shape[0], tri_lags.shape[0]
    _, counts = variogram_triangle_host(xyz, tri_lags, cos_t, sin_t, tolerance)
    offsets = np.zeros(n_lags + 1, dtype=np.int64)
    np.cumsum(counts, out=offsets[1:])
    values = np.empty(int(offsets[-1]), dtype=np.float64)
    rc = L.pib_variogram_triangle_cloud_host(_host_ptr(xyz), n, _host_ptr(tri_lags), n_lags, float(cos_t), float(sin_t),
                                             float(tolerance), _host_ptr(offsets), _host_ptr(values))
    check(rc, "pib_variogram_triangle_cloud_host")
    return [values[offsets[b]:offsets[b + 1]] for b in range(n_lags)]


def variogram_cloud_host(xyz, edges, lower=None, W=None):
    """Per-lag lists of (z_i - z_j)^2 over ordered pairs, in the reference's (i, j) order.
    One direction at most (``W`` = 4 doubles).  Returns a list of float64 arrays, one per lag."""
    L = load()
    xyz, edges = _f64(xyz), _f64(edges)
    n, n_lags = xyz.shape[0], edges.shape[0]
    Wc = None if W is None else _f64(np.asarray(W).reshape(1, 4))
    lower_c = None if lower is None else _f64(lower)
    counts = variogram_host(xyz, edges, lower=lower_c, W=Wc)["count"].reshape(-1)
    offsets = np.zeros(n_lags + 1, dtype=np.int64)
    np.cumsum(2 * counts, out=offsets[1:])
    values = np.empty(int(offsets[-1]), dtype=np.float64)
    rc = L.pib_variogram_cloud_host(_host_ptr(xyz), n, _host_ptr(edges), _host_ptr(lower_c), n_lags, _host_ptr(Wc),
                                    0 if W is None else 1, _host_ptr(offsets), _host_ptr(values))
    check(rc, "pib_variogram_cloud_host")
    return [values[offsets[b]:offsets[b + 1]] for b in range(n_lags)]


def variogram_device(xyz_t, edges, lower=None, W=None, shard=(0, 1), want_pairs=False):
    """Device-resident call: ``xyz_t`` is a CUDA float64 torch tensor ``[n, 3]``; returns CUDA tensors
    (enqueued on torch's current stream; nothing is copied to the host unless ``want_pairs``)."""
    import torch
    L = load()
    assert xyz_t.is_cuda and xyz_t.dtype == torch.float64 and xyz_t.is_contiguous()
    edges = _f64(edges)
    n, n_lags = xyz_t.shape[0], edges.shape[0]
    n_dirs = 0 if W is None else int(np.asarray(W).reshape(-1, 4).shape[0])
    Wc = None if W is None else _f64(np.asarray(W).reshape(-1, 4))
    lower_c = None if lower is None else _f64(lower)
    shape = (n_lags,) if n_dirs == 0 else (n_dirs, n_lags)
    sums = torch.empty((3,) + shape, dtype=torch.float64, device=xyz_t.device)
    cnt = torch.empty(shape, dtype=torch.int64, device=xyz_t.device)
    evaluated = ctypes.c_int64(0)
    with torch.cuda.device(xyz_t.device):
        stream = torch.cuda.current_stream().cuda_stream
        rc = L.pib_variogram_device(xyz_t.data_ptr(), n, _host_ptr(edges), _host_ptr(lower_c), n_lags, _host_ptr(Wc),
                                    n_dirs, int(shard[0]), int(shard[1]), sums[0].data_ptr(), sums[1].data_ptr(),
                                    sums[2].data_ptr(), cnt.data_ptr(),
                                    ctypes.cast(ctypes.byref(evaluated), ctypes.c_void_p) if want_pairs else None,
                                    ctypes.c_void_p(stream))
    check(rc, "pib_variogram_device")
    return dict(sum_sq=sums[0], sum_prod=sums[1], sum_val=sums[2], count=cnt, pairs_evaluated=int(evaluated.value))


def krige_host(known, unknown, model, params, mode, k, process_mean=0.0, values=None,
               want_neighbors=False, neighbors_range=0.0, direction=None, max_tick=5.0, use_all=False,
               allow_lsa=False):
    """Host-buffer call. ``params`` is ``[n_sets, 3]`` (nugget, sill, rang).  ``direction`` (degrees) switches to
    the directional neighbour selection within ``neighbors_range``; ``allow_lsa`` solves singular systems by least
    squares (status 4) instead of leaving them marked singular (status 1).
    Returns (out[n_sets, m, 4], nbr_idx[m, k] or None, status[n_sets, m]); with ``mode=KRIGE_BOTH`` the arrays get a
    leading axis of 2: ordinary kriging first, then simple kriging (one neighbour search and one elimination for both)."""
    L = load()
    known, unknown = _f64(known), _f64(unknown).reshape(-1, 2)
    params = _f64(params).reshape(-1, 3)
    n, m, n_sets = known.shape[0], unknown.shape[0], params.shape[0]
    values_c = None if values is None else _f64(values).reshape(n_sets, n)
    lead = (2,) if int(mode) == KRIGE_BOTH else ()
    out = np.empty(lead + (n_sets, m, 4))
    status = np.zeros(lead + (n_sets, m), dtype=np.uint8)
    idx = np.full((m, k), -1, dtype=np.int32) if want_neighbors else None
    rc = L.pib_krige_ex_host(_host_ptr(known), n, _host_ptr(values_c), n_sets, _host_ptr(params), _host_ptr(unknown), m,
                             int(model), int(mode), float(process_mean), int(k), float(neighbors_range),
                             float("nan") if direction is None else float(direction), float(max_tick),
                             1 if use_all else 0, 1 if allow_lsa else 0, _host_ptr(out), _host_ptr(idx), _host_ptr(status))
    check(rc, "pib_krige_ex_host")
    return out, idx, status


def krige_column_host(known, unknown, model, params, mode, k, column, process_mean=0.0, values=None, neighbors_range=0.0,
                      direction=None, max_tick=5.0, use_all=False, allow_lsa=False):
    """``krige_host`` that brings back one column of the result rows: ``column`` 0 = predictions, 1 = kriging variances.
    Returns (col[n_sets, m], status[n_sets, m]) (a leading axis of 2 with ``mode=KRIGE_BOTH``)."""
    L = load()
    known, unknown = _f64(known), _f64(unknown).reshape(-1, 2)
    params = _f64(params).reshape(-1, 3)
    n, m, n_sets = known.shape[0], unknown.shape[0], params.shape[0]
    values_c = None if values is None else _f64(values).reshape(n_sets, n)
    lead = (2,) if int(mode) == KRIGE_BOTH else ()
    col = np.empty(lead + (n_sets, m))
    status = np.zeros(lead + (n_sets, m), dtype=np.uint8)
    rc = L.pib_krige_column_host(_host_ptr(known), n, _host_ptr(values_c), n_sets, _host_ptr(params), _host_ptr(unknown), m,
                                 int(model), int(mode), float(process_mean), int(k), float(neighbors_range),
                                 float("nan") if direction is None else float(direction), float(max_tick),
                                 1 if use_all else 0, 1 if allow_lsa else 0, int(column), _host_ptr(col), _host_ptr(status))
    check(rc, "pib_krige_column_host")
    return col, status


def krige_loo_host(known, model, params, mode, k, process_mean=0.0, neighbors_range=0.0, direction=None, max_tick=5.0,
                   use_all=False, allow_lsa=False):
    """Leave-one-out kriging of every known row from the others.  Returns (out[n, 4], status[n])."""
    L = load()
    known = _f64(known)
    params = _f64(params).reshape(1, 3)
    n = known.shape[0]
    out = np.zeros((n, 4))
    status = np.zeros(n, dtype=np.uint8)
    rc = L.pib_krige_loo_host(_host_ptr(known), n, _host_ptr(params), int(model), int(mode), float(process_mean), int(k),
                              float(neighbors_range), float("nan") if direction is None else float(direction),
                              float(max_tick), 1 if use_all else 0, 1 if allow_lsa else 0, _host_ptr(out),
                              _host_ptr(status))
    check(rc, "pib_krige_loo_host")
    return out, status


def krige_device(known_t, unknown_t, model, params, mode, k, process_mean=0.0, values_t=None,
                 want_neighbors=False, want_status=True):
    """Device-resident call on torch CUDA tensors; returns CUDA tensors, enqueued on the current stream."""
    import torch
    L = load()
    for t in (known_t, unknown_t):
        assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
    params = _f64(params).reshape(-1, 3)
    n, m, n_sets = known_t.shape[0], unknown_t.shape[0], params.shape[0]
    dev = known_t.device
    lead = (2,) if int(mode) == KRIGE_BOTH else ()
    out = torch.empty(lead + (n_sets, m, 4), dtype=torch.float64, device=dev)
    status = torch.empty(lead + (n_sets, m), dtype=torch.uint8, device=dev) if want_status else None
    idx = torch.empty((m, k), dtype=torch.int32, device=dev) if want_neighbors else None
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        rc = L.pib_krige_device(known_t.data_ptr(), n, None if values_t is None else values_t.data_ptr(), n_sets,
                                _host_ptr(params), unknown_t.data_ptr(), m, int(model), int(mode), float(process_mean),
                                int(k), out.data_ptr(), None if idx is None else idx.data_ptr(),
                                None if status is None else status.data_ptr(), ctypes.c_void_p(stream))
    check(rc, "pib_krige_device")
    return out, idx, status


def gamma_host(model, nugget, sill, rang, h):
    L = load()
    h = _f64(h)
    out = np.zeros_like(h)
    check(L.pib_gamma_host(int(model), float(nugget), float(sill), float(rang), _host_ptr(h), h.size, _host_ptr(out)),
          "pib_gamma_host")
    return out


def fp64_peak_tflops():
    L = load()
    v = ctypes.c_double(0.0)
    check(L.pib_fp64_peak_tflops(ctypes.cast(ctypes.byref(v), ctypes.c_void_p)), "pib_fp64_peak_tflops")
    return float(v.value)


KERNEL_PAIR, KERNEL_KNN, KERNEL_SOLVE = 0, 1, 2


def last_kernel_ms(which):
    """(milliseconds, launches) of one hot kernel during the last library call (CUDA events)."""
    L = load()
    ms, launches = ctypes.c_double(0.0), ctypes.c_int(0)
    check(L.pib_last_kernel_ms(int(which), ctypes.cast(ctypes.byref(ms), ctypes.c_void_p),
                               ctypes.cast(ctypes.byref(launches), ctypes.c_void_p)), "pib_last_kernel_ms")
    return float(ms.value), int(launches.value)


def timing_accumulate(on):
    """Keep the kernel-time records across calls (``last_kernel_ms`` then covers every call since) or go back to per-call
    records; either way the records are cleared."""
    check(load().pib_timing_accumulate(1 if on else 0), "pib_timing_accumulate")


def last_kernel_name(which):
    """Name of the kernel variant the last library call launched for timing slot ``which``."""
    L = load()
    buf = ctypes.create_string_buffer(128)
    check(L.pib_last_kernel_name(int(which), buf, 128), "pib_last_kernel_name")
    return buf.value.decode()


def trim():
    """Return the library's cached device scratch to the device (synchronises)."""
    check(load().pib_trim(), "pib_trim")
