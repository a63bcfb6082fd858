// Point kriging (ordinary / simple), batched over unknown locations, sm_100a.
//
// What the reference does per unknown location, in a Python loop
// (src/pyinterpolate/kriging/point/ordinary.py:205-219, simple.py:206-221):
//   cdist to ALL known points + a full argsort (transform/select_points.py:83-101), gamma(h) for
//   the k neighbours and their k x k distance matrix (kriging/utils/point_kriging_solve.py:90-99),
//   a dense (k+1)^2 (OK, ordinary.py:108-115) or k^2 (SK, simple.py:102-113) system through
//   np.linalg.solve = LAPACK dgesv (point_kriging_solve.py:135-136), then zhat and sigma^2.
//
// What this file does instead:
//   knn_kernel    one warp per location; exact k nearest by (distance, original index) from a
//                 uniform grid (spatial.cu), expanding square rings until the k-th distance is
//                 covered; candidates are filtered against the running k-th best and merged with a
//                 warp bitonic sort in shared memory.  Distances are sqrt(dx*dx + dy*dy) un-fused,
//                 exactly the cdist values the reference sorts.
//   solve_kernel  one warp per location; assembles the system in shared memory in the reference's
//                 row order (neighbours by ascending distance, ones border last) and runs LU with
//                 partial pivoting like dgesv, so the pivot sequence follows the reference.
#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"

namespace pib {

constexpr int KNN_WARPS = 8;
constexpr int MAX_K = 128;
constexpr int MAX_SMEM_K = 232448;

struct __align__(16) Cand {
    double d;
    int32_t orig;   // row in the caller's known array (tie-break)
    int32_t pos;    // row in the cell-sorted arrays
};

static __device__ __forceinline__ bool cand_less(const Cand &a, const Cand &b)
{
    return a.d < b.d || (a.d == b.d && a.orig < b.orig);
}

// ascending bitonic sort of buf[0..n2) (n2 a power of two) by one warp
static __device__ void warp_bitonic_sort(Cand *buf, int n2, int lane)
{
    for (int size = 2; size <= n2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncwarp();
            for (int t = lane; t < (n2 >> 1); t += 32) {
                const int lo = ((t / stride) * stride * 2) + (t % stride);
                const int hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const Cand a = buf[lo], b = buf[hi];
                if (cand_less(b, a) == up) { buf[lo] = b; buf[hi] = a; }
            }
        }
    }
    __syncwarp();
}

struct KnnParams {
    GridIndex g;
    const double *unknown;      // [m][2]
    const int32_t *order;       // [count] location ids to process (cell-sorted)
    int64_t count;
    int kk;                     // min(k, n)
    int k2;                     // power of two >= kk
    int r0;                     // first ring radius in cells
    int32_t *nbr_pos;           // [count][kk]
    double *nbr_d;              // [count][kk]
    int32_t *nbr_idx;           // optional [m][k] (original rows, -1 padded), indexed by location id
    int k;
};

__global__ void __launch_bounds__(KNN_WARPS * 32) knn_kernel(const KnnParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t slot = (int64_t)blockIdx.x * KNN_WARPS + warp;
    if (slot >= p.count) return;
    const int k2 = p.k2, kk = p.kk;
    Cand *best = reinterpret_cast<Cand *>(smem_raw) + (size_t)warp * 2 * k2;   // [0,k2) sorted best, [k2,2k2) staging
    const GridIndex &g = p.g;

    const int32_t q = p.order[slot];
    const double ux = p.unknown[2 * (int64_t)q], uy = p.unknown[2 * (int64_t)q + 1];
    const int cx = clampi((int)floor((ux - g.x0) * g.inv_h), 0, g.gx - 1);
    const int cy = clampi((int)floor((uy - g.y0) * g.inv_h), 0, g.gy - 1);

    Cand inf_c; inf_c.d = INFINITY; inf_c.orig = 0x7fffffff; inf_c.pos = -1;
    for (int t = lane; t < 2 * k2; t += 32) best[t] = inf_c;
    __syncwarp();

    int staged = 0;                 // entries in the staging half
    Cand thr = inf_c;               // current k-th best (best[kk-1])
    int r_done = -1;                // rings [0, r_done] scanned
    int r = p.r0;
    const double slack = 1e-9 * (fabs(ux) + fabs(uy) + fabs(g.x0) + fabs(g.y0) + g.h * (g.gx + g.gy));

    while (true) {
        const int x_lo = max(cx - r, 0), x_hi = min(cx + r, g.gx - 1);
        const int y_lo = max(cy - r, 0), y_hi = min(cy + r, g.gy - 1);
        for (int yy = y_lo; yy <= y_hi; ++yy) {
            // rows inside the already scanned square only contribute their two new side strips
            const bool inner_row = (r_done >= 0) && (yy >= cy - r_done) && (yy <= cy + r_done);
            for (int part = 0; part < (inner_row ? 2 : 1); ++part) {
                int xa, xb;
                if (!inner_row) { xa = x_lo; xb = x_hi; }
                else if (part == 0) { xa = x_lo; xb = min(cx - r_done - 1, x_hi); }
                else { xa = max(cx + r_done + 1, x_lo); xb = x_hi; }
                if (xa > xb) continue;
                const int s0 = g.cell_start[yy * g.gx + xa], s1 = g.cell_start[yy * g.gx + xb + 1];
                for (int base = s0; base < s1; base += 32) {
                    const int pos = base + lane;
                    Cand c = inf_c;
                    if (pos < s1) {
                        // cdist: sqrt(dx*dx + dy*dy), nothing fused (transform/select_points.py:85)
                        const double dx = __dsub_rn(ux, g.x[pos]), dy = __dsub_rn(uy, g.y[pos]);
                        c.d = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
                        c.orig = g.orig[pos];
                        c.pos = pos;
                    }
                    const bool pass = (pos < s1) && cand_less(c, thr);
                    const unsigned mask = __ballot_sync(0xffffffffu, pass);
                    const int n_pass = __popc(mask);
                    if (n_pass == 0) continue;
                    if (staged + n_pass > k2) {                      // staging full: merge first
                        warp_bitonic_sort(best, 2 * k2, lane);
                        for (int t = k2 + lane; t < 2 * k2; t += 32) best[t] = inf_c;
                        __syncwarp();
                        staged = 0;
                        thr = best[kk - 1];
                        // re-test against the tighter threshold
                        const bool pass2 = pass && cand_less(c, thr);
                        const unsigned mask2 = __ballot_sync(0xffffffffu, pass2);
                        if (pass2) best[k2 + __popc(mask2 & ((1u << lane) - 1u))] = c;
                        staged = __popc(mask2);
                    } else {
                        if (pass) best[k2 + staged + __popc(mask & ((1u << lane) - 1u))] = c;
                        staged += n_pass;
                    }
                    __syncwarp();
                }
            }
        }
        r_done = r;
        if (staged > 0) {
            warp_bitonic_sort(best, 2 * k2, lane);
            for (int t = k2 + lane; t < 2 * k2; t += 32) best[t] = inf_c;
            __syncwarp();
            staged = 0;
            thr = best[kk - 1];
        }
        const bool whole_grid = (x_lo == 0 && y_lo == 0 && x_hi == g.gx - 1 && y_hi == g.gy - 1);
        if (whole_grid) break;
        // radius around the location that the scanned square is guaranteed to cover
        double rho = INFINITY;
        if (cx - r > 0) rho = fmin(rho, ux - (g.x0 + (cx - r) * g.h));
        if (cx + r < g.gx - 1) rho = fmin(rho, (g.x0 + (cx + r + 1) * g.h) - ux);
        if (cy - r > 0) rho = fmin(rho, uy - (g.y0 + (cy - r) * g.h));
        if (cy + r < g.gy - 1) rho = fmin(rho, (g.y0 + (cy + r + 1) * g.h) - uy);
        rho -= slack;
        if (thr.d <= rho) break;                                 // k-th neighbour is inside the covered disc
        if (isinf(thr.d)) r = 2 * r + 1;                          // fewer than k found so far
        else r = max(r + 1, min((int)ceil(thr.d * g.inv_h) + 1, 2 * r + 1));
    }

    for (int t = lane; t < kk; t += 32) {
        p.nbr_pos[slot * kk + t] = best[t].pos;
        p.nbr_d[slot * kk + t] = best[t].d;
    }
    if (p.nbr_idx)
        for (int t = lane; t < p.k; t += 32) p.nbr_idx[(int64_t)q * p.k + t] = t < kk ? best[t].orig : -1;
}

// ---------------------------------------------------------------------------------------------
// gamma(h): semivariogram/theoretical/variogram_models/models.py:41-440 (sill = partial sill).
// h == 0 gives the nugget for every model (what _get_zero_lag_value + the formulas evaluate to).
// ---------------------------------------------------------------------------------------------
static __device__ __forceinline__ double gamma_model(int model, double h, double nugget, double sill, double rang)
{
    if (h == 0.0) return nugget;
    const double ar = h / rang;
    switch (model) {
    case PIB_MODEL_CIRCULAR: {
        if (!(h <= rang)) return nugget + sill;
        const double pic = 2.0 / 3.14159265358979323846;
        return nugget + sill * (1.0 - (pic * acos(ar)) + (pic * ar) * sqrt(1.0 - ar * ar));
    }
    case PIB_MODEL_CUBIC: {
        if (!(h <= rang)) return nugget + sill;
        const double a2 = ar * ar, a3 = a2 * ar, a5 = a3 * a2, a7 = a5 * a2;
        return nugget + sill * (a2 * 7.0 + a3 * -8.75 + a5 * 3.5 + a7 * -0.75);
    }
    case PIB_MODEL_EXPONENTIAL: return nugget + sill * (1.0 - exp(-ar));
    case PIB_MODEL_GAUSSIAN: return nugget + sill * (1.0 - exp(-((h * h) / (rang * rang))));
    case PIB_MODEL_LINEAR: return (h <= rang) ? nugget + sill * ar : nugget + sill;
    case PIB_MODEL_POWER: return (h <= rang) ? nugget + sill * (ar * ar) : nugget + sill;
    case PIB_MODEL_SPHERICAL: {
        if (!(h <= rang)) return nugget + sill;
        return nugget + sill * (1.5 * ar - 0.5 * (ar * ar * ar));
    }
    default: return nan("");
    }
}

__global__ void gamma_kernel(int model, double nugget, double sill, double rang, const double *h, int64_t n, double *out)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = gamma_model(model, h[i], nugget, sill, rang);
}

struct SolveParams {
    GridIndex g;
    const double *unknown;
    const int32_t *order;
    int64_t count;
    int64_t m, n;
    int kk, k;
    const int32_t *nbr_pos;
    const double *nbr_d;
    const double *values;       // NULL or [n_sets][n] by original row
    const double *params;       // device [n_sets][3]
    int n_sets;
    int model, mode;
    double process_mean;
    double *out;                // [n_sets][m][4]
    uint8_t *status;            // NULL or [n_sets][m]
    int warps;                  // warps per CTA
    int stride;                 // row stride of the augmented matrix in doubles (odd)
};

// per-warp shared memory: A[dim][stride] | rhs0[dim] | dinv[dim] | nx[kk] ny[kk] nd[kk] | npos[kk] (int)
static size_t solve_smem_per_warp(int kk, int dim, int stride)
{
    size_t b = (size_t)dim * stride * 8 + (size_t)dim * 8 * 2 + (size_t)kk * 8 * 3 + (size_t)((kk + 1) / 2 * 2) * 4;
    return (b + 15) / 16 * 16;
}

__global__ void solve_kernel(const SolveParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t slot = (int64_t)blockIdx.x * p.warps + warp;
    if (slot >= p.count) return;
    const int kk = p.kk, mode = p.mode;
    const int dim = (mode == PIB_KRIGE_ORDINARY) ? kk + 1 : kk;
    const int st = p.stride;
    const size_t per_warp = ((size_t)dim * st * 8 + (size_t)dim * 16 + (size_t)kk * 24 + (size_t)((kk + 1) / 2 * 2) * 4 + 15) / 16 * 16;
    double *A = reinterpret_cast<double *>(smem_raw + per_warp * warp);
    double *rhs0 = A + (size_t)dim * st;
    double *dinv = rhs0 + dim;
    double *nx = dinv + dim, *ny = nx + kk, *nd = ny + kk;
    int32_t *npos = reinterpret_cast<int32_t *>(nd + kk);
    const GridIndex &g = p.g;

    const int32_t q = p.order[slot];
    const double ux = p.unknown[2 * (int64_t)q], uy = p.unknown[2 * (int64_t)q + 1];
    for (int t = lane; t < kk; t += 32) {
        const int pos = p.nbr_pos[slot * kk + t];
        npos[t] = pos; nx[t] = g.x[pos]; ny[t] = g.y[pos]; nd[t] = p.nbr_d[slot * kk + t];
    }
    __syncwarp();

    for (int set = 0; set < p.n_sets; ++set) {
        const double nugget = p.params[3 * set], sill = p.params[3 * set + 1], rang = p.params[3 * set + 2];
        // ---- assemble [Gamma 1; 1^T 0 | k 1] in the reference's row order (ordinary.py:108-115) ----
        for (int i = 0; i < kk; ++i) {
            for (int j = i + 1 + lane; j < kk; j += 32) {
                const double dx = __dsub_rn(nx[i], nx[j]), dy = __dsub_rn(ny[i], ny[j]);
                const double h = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
                const double gm = gamma_model(p.model, h, nugget, sill, rang);
                A[i * st + j] = gm; A[j * st + i] = gm;
            }
        }
        for (int i = lane; i < kk; i += 32) {
            A[i * st + i] = nugget;                                   // gamma(0)
            const double kv = gamma_model(p.model, nd[i], nugget, sill, rang);
            A[i * st + dim] = kv; rhs0[i] = kv;
            if (mode == PIB_KRIGE_ORDINARY) { A[i * st + kk] = 1.0; A[kk * st + i] = 1.0; }
        }
        if (mode == PIB_KRIGE_ORDINARY && lane == 0) { A[kk * st + kk] = 0.0; A[kk * st + dim] = 1.0; rhs0[kk] = 1.0; }
        __syncwarp();

        // ---- LU with partial pivoting on the augmented matrix (dgesv: dgetrf + dgetrs) ----------
        bool singular = false;
        for (int j = 0; j < dim; ++j) {
            double bv = -1.0; int bi = j;
            for (int r = j + lane; r < dim; r += 32) {
                const double v = fabs(A[r * st + j]);
                if (v > bv) { bv = v; bi = r; }
            }
            for (int o = 16; o; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (!(bv > 0.0) || isinf(bv)) { singular = true; break; }   // zero / NaN pivot column
            if (bi != j) {
                for (int c = lane; c <= dim; c += 32) {
                    const double t = A[j * st + c]; A[j * st + c] = A[bi * st + c]; A[bi * st + c] = t;
                }
            }
            __syncwarp();
            const double inv = 1.0 / A[j * st + j];
            if (lane == 0) dinv[j] = inv;
            for (int r = j + 1 + lane; r < dim; r += 32) A[r * st + j] *= inv;   // dgetf2 scales by the reciprocal
            __syncwarp();
            // trailing update, lanes own columns (including the right-hand side column `dim`)
            for (int c0 = j + 1; c0 <= dim; c0 += 32) {
                const int c = c0 + lane;
                if (c <= dim) {
                    const double ajc = A[j * st + c];
                    for (int r = j + 1; r < dim; ++r) A[r * st + c] = __fma_rn(-A[r * st + j], ajc, A[r * st + c]);
                }
            }
            __syncwarp();
        }
        double zhat = nan(""), sigma = nan("");
        uint8_t stat = PIB_STATUS_SINGULAR;
        if (!singular) {
            // back substitution, column oriented: x_i = b_i / U_ii, then b_r -= U_ri x_i for r < i
            for (int i = dim - 1; i >= 0; --i) {
                const double xi = A[i * st + dim] * dinv[i];
                __syncwarp();
                if (lane == 0) A[i * st + dim] = xi;
                for (int r = lane; r < i; r += 32) A[r * st + dim] = __fma_rn(-A[r * st + i], xi, A[r * st + dim]);
                __syncwarp();
            }
            // zhat = z . w (ordinary.py:121) or (z - mean) . w + mean (simple.py:114-116); sigma = w . k (:123 / :118)
            double zs = 0.0, ss = 0.0;
            for (int i = lane; i < dim; i += 32) {
                const double w = A[i * st + dim];
                ss = __fma_rn(w, rhs0[i], ss);
                if (i < kk) {
                    double zv = p.values ? p.values[(size_t)set * p.n + g.orig[npos[i]]] : g.z[npos[i]];
                    if (mode == PIB_KRIGE_SIMPLE) zv -= p.process_mean;
                    zs = __fma_rn(zv, w, zs);
                }
            }
            for (int o = 16; o; o >>= 1) {
                zs += __shfl_xor_sync(0xffffffffu, zs, o);
                ss += __shfl_xor_sync(0xffffffffu, ss, o);
            }
            zhat = (mode == PIB_KRIGE_SIMPLE) ? zs + p.process_mean : zs;
            sigma = ss;
            stat = PIB_STATUS_OK;
            if (sigma < 0.0) { sigma = nan(""); stat = PIB_STATUS_NEGATIVE_VARIANCE; }   // ordinary.py:125-126
        }
        if (lane == 0) {
            double *o = p.out + ((size_t)set * p.m + q) * 4;
            o[0] = zhat; o[1] = sigma; o[2] = ux; o[3] = uy;
            if (p.status) p.status[(size_t)set * p.m + q] = stat;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
int krige_device(const double *d_known, int64_t n, const double *d_values, int n_sets, const double *params,
                 const double *d_unknown, int64_t m, int model, int mode, double process_mean, int k,
                 double *d_out, int32_t *d_nbr_idx, uint8_t *d_status, cudaStream_t stream)
{
    PIB_CHECK(n >= 1 && m >= 0 && k >= 1, PIB_ERR_INVALID, "krige: need n >= 1, m >= 0, k >= 1");
    PIB_CHECK(n_sets >= 1 && params != nullptr, PIB_ERR_INVALID, "krige: need at least one parameter set");
    PIB_CHECK(model >= 0 && model <= 6, PIB_ERR_INVALID, "krige: unknown model id");
    PIB_CHECK(mode == PIB_KRIGE_ORDINARY || mode == PIB_KRIGE_SIMPLE, PIB_ERR_INVALID, "krige: unknown mode");
    PIB_CHECK(m < ((int64_t)1 << 31) && n < ((int64_t)1 << 31), PIB_ERR_UNSUPPORTED, "krige: more than 2^31-1 rows");
    const int kk = (int)std::min<int64_t>(k, n);
    PIB_CHECK(kk <= MAX_K, PIB_ERR_UNSUPPORTED, "krige: no_neighbors above 128 is not supported");
    timing_reset(PIB_T_KNN);
    timing_reset(PIB_T_SOLVE);
    if (m == 0) return PIB_OK;

    Scratch ws(stream);
    GridIndex g;
    PIB_TRY(build_grid_index(d_known, n, 4.0, ws, stream, &g));
    int32_t *order;
    PIB_TRY(sort_queries_by_cell(d_unknown, m, g, ws, stream, &order));
    double *d_params;
    PIB_TRY(ws.alloc(&d_params, (size_t)n_sets * 3));
    PIB_CUDA(cudaMemcpyAsync(d_params, params, sizeof(double) * 3 * n_sets, cudaMemcpyHostToDevice, stream));

    int k2 = 32;                       // staging must hold a full warp batch
    while (k2 < kk) k2 <<= 1;
    const double per_cell = (double)n / ((double)g.gx * g.gy);
    int r0 = (int)std::ceil(std::sqrt(1.3 * kk / (3.14159265358979 * std::max(per_cell, 1e-9))));
    r0 = std::max(1, std::min(r0, std::max(g.gx, g.gy)));

    const int dim = (mode == PIB_KRIGE_ORDINARY) ? kk + 1 : kk;
    const int stride = (dim + 1) | 1;
    const size_t per_warp = solve_smem_per_warp(kk, dim, stride);
    int solve_warps = (int)std::min<size_t>(8, (size_t)(MAX_SMEM_K - 1024) / per_warp);
    PIB_CHECK(solve_warps >= 1, PIB_ERR_UNSUPPORTED, "krige: system does not fit in shared memory");
    PIB_CUDA(cudaFuncSetAttribute(solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per_warp * solve_warps)));
    const size_t knn_smem = (size_t)KNN_WARPS * 2 * k2 * sizeof(Cand);
    PIB_CUDA(cudaFuncSetAttribute(knn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)knn_smem));

    const int64_t chunk = std::min<int64_t>(m, (int64_t)1 << 21);
    int32_t *nbr_pos;
    double *nbr_d;
    PIB_TRY(ws.alloc(&nbr_pos, (size_t)chunk * kk));
    PIB_TRY(ws.alloc(&nbr_d, (size_t)chunk * kk));

    for (int64_t begin = 0; begin < m; begin += chunk) {
        const int64_t count = std::min(chunk, m - begin);
        KnnParams kp;
        kp.g = g; kp.unknown = d_unknown; kp.order = order + begin; kp.count = count; kp.kk = kk; kp.k2 = k2; kp.r0 = r0;
        kp.nbr_pos = nbr_pos; kp.nbr_d = nbr_d; kp.nbr_idx = d_nbr_idx; kp.k = k;
        timing_begin(PIB_T_KNN, stream);
        knn_kernel<<<(int)((count + KNN_WARPS - 1) / KNN_WARPS), KNN_WARPS * 32, knn_smem, stream>>>(kp);
        timing_end(PIB_T_KNN, stream);
        PIB_CUDA(cudaGetLastError());
        SolveParams sp;
        sp.g = g; sp.unknown = d_unknown; sp.order = order + begin; sp.count = count; sp.m = m; sp.n = n; sp.kk = kk; sp.k = k;
        sp.nbr_pos = nbr_pos; sp.nbr_d = nbr_d; sp.values = d_values; sp.params = d_params; sp.n_sets = n_sets;
        sp.model = model; sp.mode = mode; sp.process_mean = process_mean; sp.out = d_out; sp.status = d_status;
        sp.warps = solve_warps; sp.stride = stride;
        timing_begin(PIB_T_SOLVE, stream);
        solve_kernel<<<(int)((count + solve_warps - 1) / solve_warps), solve_warps * 32, per_warp * solve_warps, stream>>>(sp);
        timing_end(PIB_T_SOLVE, stream);
        PIB_CUDA(cudaGetLastError());
    }
    return PIB_OK;
}

int gamma_device(int model, double nugget, double sill, double rang, const double *d_h, int64_t n, double *d_out,
                 cudaStream_t stream)
{
    if (n == 0) return PIB_OK;
    gamma_kernel<<<(int)((n + 255) / 256), 256, 0, stream>>>(model, nugget, sill, rang, d_h, n, d_out);
    PIB_CUDA(cudaGetLastError());
    return PIB_OK;
}

}  // namespace pib
