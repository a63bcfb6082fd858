/*
 * pyinterp_b200.h — C ABI of libpyinterp_b200.so (hand-written CUDA for sm_100a).
 *
 * The reference (DataverseLabs/pyinterpolate 1.0.3) is pure Python and has no FFI of its own;
 * its drop-in boundary is the set of Python call signatures re-exported in
 * src/pyinterpolate/__init__.py:15-24.  Each entry point below replaces the NumPy/SciPy
 * internals behind one of those calls (file:line relative to src/pyinterpolate/).  The
 * Python binding a maintainer adds is a ctypes stub — see INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary
 *   - `*_device` entry points take DEVICE pointers and enqueue on `stream`
 *     (a cudaStream_t passed as void*, NULL = legacy default stream); results are ready when the
 *     stream reaches the end of the enqueued work (the call itself only waits for a 32-byte
 *     bounding-box read-back while planning)
 *   - `*_host` entry points take HOST pointers, do their own H2D / D2H copies and return
 *     after the results are in host memory
 *   - every function returns PIB_OK (0) or a negative error; pib_last_error() has the text
 *   - inputs are borrowed and never modified; outputs are caller-allocated
 *   - all floating point is IEEE fp64; pair counts are int64; neighbour indices int32
 */
#ifndef PYINTERP_B200_H
#define PYINTERP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PIB_API __attribute__((visibility("default")))
#else
#define PIB_API
#endif

#define PIB_OK 0
#define PIB_ERR_INVALID -1      /* bad argument (NULL pointer, n < 0, non-ascending lags, ...) */
#define PIB_ERR_CUDA -2         /* a CUDA runtime call failed */
#define PIB_ERR_UNSUPPORTED -3  /* size outside what the kernels are built for */

/* theoretical variogram model ids — alphabetical, as in
 * semivariogram/theoretical/variogram_models/functions.py:8-14 (ALL_MODELS) */
#define PIB_MODEL_CIRCULAR 0
#define PIB_MODEL_CUBIC 1
#define PIB_MODEL_EXPONENTIAL 2
#define PIB_MODEL_GAUSSIAN 3
#define PIB_MODEL_LINEAR 4
#define PIB_MODEL_POWER 5
#define PIB_MODEL_SPHERICAL 6

#define PIB_KRIGE_ORDINARY 0    /* kriging/point/ordinary.py:30-131  (ok_calc)  */
#define PIB_KRIGE_SIMPLE 1      /* kriging/point/simple.py:26-126    (sk_calc)  */
#define PIB_KRIGE_BOTH 2        /* both from ONE neighbour search, assembly and elimination: the outputs hold the ordinary-kriging
                                   block [n_sets][m][..] first, then the simple-kriging block */

/* per-location status written by the kriging entry points */
#define PIB_STATUS_OK 0
#define PIB_STATUS_SINGULAR 1        /* zero pivot: the reference raises RuntimeError (kriging/utils/errors.py:1-15) */
#define PIB_STATUS_NEGATIVE_VARIANCE 2 /* sigma^2 < 0 -> NaN, like ordinary.py:125-126 */
#define PIB_STATUS_TOO_MANY_NEIGHBORS 3 /* use_all_neighbors_in_range selected more than 128 rows (unsupported) */
#define PIB_STATUS_LSA 4             /* singular system solved by least squares (allow_lsa): the reference warns
                                        LeastSquaresApproximationWarning (kriging/utils/point_kriging_solve.py:142-145) */
#define PIB_STATUS_ZERO_MATRIX 5     /* singular system whose matrix or right-hand side is all zero: weights = 0 and
                                        ZerosMatrixWarning (point_kriging_solve.py:138-140) */

PIB_API const char *pib_last_error(void);
PIB_API int pib_version(void);
/* number of SMs of the current device (grid sizing evidence for callers / tests) */
PIB_API int pib_device_sm_count(void);
/* The library takes its scratch from the device's stream-ordered default memory pool and keeps up to 4 GiB of it between
 * calls; pib_trim() synchronises the device and returns all cached scratch to it. */
PIB_API int pib_trim(void);

/* ------------------------------------------------------------------------------------------
 * Experimental variogram — one fused pass over point pairs.
 *
 * Replaces, for omnidirectional variograms (n_dirs == 0):
 *   distance/point.py:14-58            point_distance (scipy cdist, N x N matrix)
 *   distance/point.py:61-92            select_values_in_range (one masked pass per lag)
 *   semivariogram/experimental/functions/general.py:78-147, :235-303
 *                                      omnidirectional_variogram / _calc_omnidirectional_values
 *   semivariogram/experimental/functions/semivariance.py:4-24, covariance.py:4-26
 * and for ellipse-directional variograms (n_dirs >= 1, method 'e'):
 *   semivariogram/experimental/functions/directional.py:11-80, :395-465   from_ellipse
 *   distance/angular.py:412-456, :641-681   select_points_within_ellipse
 *
 * xyz        [n][3] row-major rows [x, y, value]
 * edges      [n_lags] ascending lag upper limits (semivariogram/lags/lags.py:6-31, host-made so
 *            they are bit-identical to the reference's np.arange); lag b is (edges[b-1], edges[b]],
 *            edges[-1] := 0                                           (HOST pointer)
 * lower      NULL, or [n_lags] lower limits for the ellipse path computed the reference's way,
 *            edges[b] - (edges[b] - edges[b-1]) (directional.py:432, angular.py:672-673);
 *            NULL means lower[b] = edges[b-1]                         (HOST pointer)
 * W          [n_dirs][4] row-major 2x2 whitening matrices (distance/angular.py:221-250), made on
 *            the host with the reference's own formula; ignored when n_dirs == 0   (HOST pointer)
 * shard_index, shard_count   this call evaluates the shard_index-th of shard_count interleaved
 *            slices of the tile-pair list (multi-GPU: one slice per rank, then a SUM all-reduce
 *            of the four outputs); (0, 1) = everything
 * outputs    [max(n_dirs,1)][n_lags], sums over UNORDERED pairs (i < j) that fall in the lag:
 *            sum_sq = sum (z_i - z_j)^2, sum_prod = sum z_i z_j, sum_val = sum (z_i + z_j),
 *            count = number of pairs.  The reference's ordered-pair outputs follow as
 *            semivariance = sum_sq / (2 count), covariance = sum_prod / count - (sum_val / (2 count))^2,
 *            n_pairs = 2 count.
 * pairs_evaluated  optional (may be NULL): number of point pairs whose distance was actually
 *            computed (tile pairs farther apart than the last lag are skipped)    (HOST pointer)
 * ---------------------------------------------------------------------------------------- */
PIB_API int pib_variogram_device(const double *d_xyz, int64_t n,
                         const double *edges, const double *lower, int n_lags,
                         const double *W, int n_dirs,
                         int shard_index, int shard_count,
                         double *d_sum_sq, double *d_sum_prod, double *d_sum_val, int64_t *d_count,
                         int64_t *pairs_evaluated, void *stream);

PIB_API int pib_variogram_host(const double *xyz, int64_t n,
                       const double *edges, const double *lower, int n_lags,
                       const double *W, int n_dirs,
                       int shard_index, int shard_count,
                       double *sum_sq, double *sum_prod, double *sum_val, int64_t *count,
                       int64_t *pairs_evaluated);

/* ------------------------------------------------------------------------------------------
 * Weighted semivariance (custom_weights).  Replaces
 *   semivariogram/weights/experimental/weighting.py:4-54          (omnidirectional, n_dirs == 0)
 *   semivariogram/experimental/functions/directional.py:276-392   (directional = ellipse, n_dirs == 1)
 * weights [n] per-point weights (non-zero).  sums [4][n_lags] (out), over UNORDERED pairs in the lag, with
 * w_p = w_i w_j / (w_i + w_j):
 *   omni:        sum w_p (z_i/w_i - z_j/w_j)^2,  sum w_p,  sum (z_i/w_i + z_j/w_j),  sum (w_i + w_j)
 *   directional: sum w_p (z_i - z_j)^2,          sum w_p,  sum (w_i z_i + w_j z_j),  sum (w_i + w_j)
 * count [n_lags] unordered pairs.  Small-data path (global atomics; sums are not bit-reproducible run to run).
 * ---------------------------------------------------------------------------------------- */
PIB_API int pib_variogram_weighted_host(const double *xyz, const double *weights, int64_t n,
                                const double *edges, const double *lower, int n_lags,
                                const double *W, int n_dirs,
                                double *sums, int64_t *count);

/* ------------------------------------------------------------------------------------------
 * Triangle ("t") directional selection, the reference's default directional method.  Replaces
 *   semivariogram/experimental/functions/directional.py:137-210, :536-565   from_triangle
 *   distance/angular.py:9-64, :188-218, :283-339, :375-409, :459-494        triangle masks
 * tri_lags [n_lags][9] (HOST), per lag h with a = radians(direction), made with numpy like the reference
 * (distance/angular.py:283-339, :599-617):  h, h cos a, h sin a, (-h) cos a, (-h) sin a,
 *   (h tol) cos(a + pi/2), (h tol) sin(a + pi/2), (h tol) cos(a - pi/2), (h tol) sin(a - pi/2)
 * cos_t, sin_t = cos a, sin a (only used to locate a pair between the lag edges; membership near an edge is
 * decided with the reference's own operation sequence).
 * sums [3][n_lags] (out) over ORDERED pairs (centre p, neighbour q) of the cleaned masks:
 *   sum (z_p - z_q)^2, sum z_p z_q, sum z_q;   count [n_lags] ordered pairs (the reference includes q == p
 *   in the first lag).  Small-data path (one thread per centre, global atomics).
 * ---------------------------------------------------------------------------------------- */
PIB_API int pib_variogram_triangle_host(const double *xyz, int64_t n,
                                const double *tri_lags, int n_lags,
                                double cos_t, double sin_t, double tolerance,
                                double *sums, int64_t *count);

/* ------------------------------------------------------------------------------------------
 * Variogram point cloud: per lag, (z_i - z_j)^2 of every ORDERED pair in the lag, in the reference's
 * order (row i ascending, then j ascending = np.where on the N x N mask).  Replaces
 *   semivariogram/experimental/functions/general.py:13-75, :150-233   omnidirectional_semivariogram_cloud
 *   semivariogram/experimental/functions/directional.py:83-134, :468-530   from_ellipse_cloud (n_dirs == 1)
 * lag_offsets [n_lags + 1] (HOST): start of each lag's block in `values`; lag b holds
 * lag_offsets[b+1] - lag_offsets[b] = 2 * count[b] values (count from pib_variogram_host).
 * values [lag_offsets[n_lags]] (HOST, out).  Output is O(pairs): meant for small N.
 * ---------------------------------------------------------------------------------------- */
PIB_API int pib_variogram_cloud_host(const double *xyz, int64_t n,
                             const double *edges, const double *lower, int n_lags,
                             const double *W, int n_dirs,
                             const int64_t *lag_offsets, double *values);

/* ------------------------------------------------------------------------------------------
 * Triangle-method point cloud — what the reference's point_cloud_semivariance returns for a direction with its DEFAULT
 * dir_neighbors_selection_method='t'.  Replaces
 *   semivariogram/experimental/functions/directional.py:213-273   from_triangle_cloud
 *   semivariogram/experimental/functions/directional.py:583-620   _get_values_triangle_cloud
 * tri_lags, cos_t, sin_t, tolerance: as for pib_variogram_triangle_host.
 * lag_offsets [n_lags + 1] (HOST): start of each lag's block in `values`; lag b holds count[b] values, count from
 * pib_variogram_triangle_host (ORDERED pairs, q == p of the first lag included).  The call returns PIB_ERR_INVALID
 * without writing if its own membership pass disagrees with the offsets.
 * values [lag_offsets[n_lags]] (HOST, out): (z_p - z_q)^2, centre p ascending, neighbour q ascending within a centre.
 * ---------------------------------------------------------------------------------------- */
PIB_API int pib_variogram_triangle_cloud_host(const double *xyz, int64_t n,
                                      const double *tri_lags, int n_lags,
                                      double cos_t, double sin_t, double tolerance,
                                      const int64_t *lag_offsets, double *values);

/* ------------------------------------------------------------------------------------------
 * Point kriging — batched neighbour search + system assembly + solve.
 *
 * Replaces the per-location Python loop of
 *   kriging/point/ordinary.py:134-221 (ordinary_kriging) / simple.py:129-223 (simple_kriging)
 * and everything below it:
 *   transform/select_points.py:42-101          select_kriging_data (cdist + full argsort)
 *   kriging/utils/point_kriging_solve.py:20-101  get_predictions (k vector, Gamma matrix)
 *   semivariogram/theoretical/variogram_models/models.py:41-440   gamma(h)
 *   kriging/utils/point_kriging_solve.py:104-149  solve_weights (np.linalg.solve = dgesv)
 *
 * known      [n][3] rows [x, y, value]
 * values     NULL, or [n_sets][n] alternative value columns sharing the known coordinates
 *            (indicator kriging: one 0/1 column per threshold, kriging/point/indicator.py:269-291);
 *            NULL means one set taken from known[:, 2]
 * params     [n_sets][3] = nugget, (partial) sill, range per set            (HOST pointer)
 * unknown    [m][2] rows [x, y]
 * model      PIB_MODEL_*; mode PIB_KRIGE_*; process_mean only used by simple kriging
 * k          no_neighbors; the k nearest known points by (distance, index) are used, all n when
 *            n < k (select_points.py:95-101 with use_all_neighbors_in_range=False)
 * out        [n_sets][m][4] rows [zhat, sigma^2 (NaN if negative), x, y]   ([2][n_sets][m][4] for PIB_KRIGE_BOTH)
 * nbr_idx    NULL, or [m][k] indices into `known`, ascending distance, -1 padded
 * status     NULL, or [n_sets][m] PIB_STATUS_*                               ([2][n_sets][m] for PIB_KRIGE_BOTH)
 * shard: locations are independent, so multi-GPU callers simply pass their own slice of `unknown`.
 * ---------------------------------------------------------------------------------------- */
PIB_API int pib_krige_device(const double *d_known, int64_t n,
                     const double *d_values, int n_sets, const double *params,
                     const double *d_unknown, int64_t m,
                     int model, int mode, double process_mean, int k,
                     double *d_out, int32_t *d_nbr_idx, uint8_t *d_status, void *stream);

PIB_API int pib_krige_host(const double *known, int64_t n,
                   const double *values, int n_sets, const double *params,
                   const double *unknown, int64_t m,
                   int model, int mode, double process_mean, int k,
                   double *out, int32_t *nbr_idx, uint8_t *status);

/* Same with the reference's remaining neighbour-selection options:
 * neighbors_range  radius of the directional search (kriging/utils/point_kriging_solve.py:68-69: model range by default)
 * direction        NaN = isotropic selection (above); otherwise the model direction in degrees: neighbours within
 *                  neighbors_range are filtered by their angular difference to it, the tolerance is widened 1, 2, ...
 *                  degrees until k qualify or it exceeds max_tick (transform/select_points.py:104-257); a location
 *                  may end up with fewer than k neighbours, or none (-> NaN row, like the reference)
 * use_all          use_all_neighbors_in_range (at most 128 rows, else status 3)
 * allow_lsa        allow_approximate_solutions: a singular system (zero pivot, what makes np.linalg.solve raise) is
 *                  solved like np.linalg.lstsq(A, b, rcond=None) — minimum-norm least squares, singular values below
 *                  eps * dim * s_max dropped — and marked PIB_STATUS_LSA (point_kriging_solve.py:134-147).
 *                  Independently of the flag, a singular system with an all-zero matrix or right-hand side gets zero
 *                  weights and PIB_STATUS_ZERO_MATRIX. */
PIB_API int pib_krige_ex_device(const double *d_known, int64_t n,
                        const double *d_values, int n_sets, const double *params,
                        const double *d_unknown, int64_t m,
                        int model, int mode, double process_mean, int k,
                        double neighbors_range, double direction, double max_tick, int use_all, int allow_lsa,
                        double *d_out, int32_t *d_nbr_idx, uint8_t *d_status, void *stream);

PIB_API int pib_krige_ex_host(const double *known, int64_t n,
                      const double *values, int n_sets, const double *params,
                      const double *unknown, int64_t m,
                      int model, int mode, double process_mean, int k,
                      double neighbors_range, double direction, double max_tick, int use_all, int allow_lsa,
                      double *out, int32_t *nbr_idx, uint8_t *status);

/* pib_krige_ex_host that hands back ONE column of the result rows: column 0 = predictions, 1 = kriging variances (NaN where
 * negative), out_column [n_sets][m] ([2][n_sets][m] for PIB_KRIGE_BOTH).  A quarter of the device-to-host traffic and of the
 * host-side copies — for callers that keep one column per parameter set, e.g. the reference's IndicatorKriging
 * (kriging/point/indicator.py:290-313 stores _pred_arr[0][1] per threshold). */
PIB_API int pib_krige_column_host(const double *known, int64_t n,
                          const double *values, int n_sets, const double *params,
                          const double *unknown, int64_t m,
                          int model, int mode, double process_mean, int k,
                          double neighbors_range, double direction, double max_tick,
                          int use_all, int allow_lsa,
                          int column, double *out_column, uint8_t *status);

/* Leave-one-out cross-validation in one batched call — replaces the loop of evaluate/cross_validation.py:78-131
 * (np.delete(points, i) + ordinary_kriging / simple_kriging at point i, for every i): location i is known row i,
 * and the neighbour search skips that row.  out [n][4] = [zhat_i, sigma^2_i, x_i, y_i]; status as above.
 * One parameter set: params = {nugget, sill, rang}.  Selection arguments as in pib_krige_ex_host. */
PIB_API int pib_krige_loo_host(const double *known, int64_t n, const double *params, int model, int mode,
                       double process_mean, int k, double neighbors_range, double direction, double max_tick,
                       int use_all, int allow_lsa, double *out, uint8_t *status);

/* gamma(h) for h[0..n) with the device implementation of the seven models (for tests) */
PIB_API int pib_gamma_host(int model, double nugget, double sill, double rang,
                   const double *h, int64_t n, double *out);

/* Device time (CUDA events on the launching stream) spent in one hot kernel during the LAST
 * variogram / kriging call of this process: which = 0 pair kernel, 1 neighbour-search kernel,
 * 2 assemble+solve kernel.  launches_out (may be NULL) receives how many times it was launched.
 * Synchronises on the recorded events. */
PIB_API int pib_last_kernel_ms(int which, double *ms_out, int *launches_out);

/* on != 0: the kernel-time records are kept ACROSS calls from now on (pib_last_kernel_ms then returns the sum and the launch
 * count over every call since), so a benchmark can time K back-to-back calls without reading — and synchronising on — the
 * events after each one; on == 0: back to per-call records.  Either value clears the records. */
PIB_API int pib_timing_accumulate(int on);

/* Name of the kernel variant behind pib_last_kernel_ms(which) in the last call (e.g. "pair_ring6_kernel<16,512,1,8>"),
 * copied into buf (NUL terminated, at most len bytes). */
PIB_API int pib_last_kernel_name(int which, char *buf, int len);

/* FP64 throughput probe: runs a dependent-chain-free DFMA kernel and returns TFLOP/s
 * (2 flop per DFMA) measured with CUDA events.  Used by bench.py for the roofline
 * denominator because MEASURED_PEAKS.json carries no FP64 figure. */
PIB_API int pib_fp64_peak_tflops(double *tflops_out);

#ifdef __cplusplus
}
#endif
#endif /* PYINTERP_B200_H */
