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
#include <cstdlib>
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

// Running "k smallest by (distance, original index)" of one warp: best[0, k2) sorted, best[k2, 2 k2) staging.
struct TopK {
    Cand *best;
    int k2, kk, staged;
    Cand thr;                       // current kk-th best
};

static __device__ __forceinline__ Cand cand_inf()
{
    Cand c; c.d = INFINITY; c.orig = 0x7fffffff; c.pos = -1;
    return c;
}

static __device__ void topk_init(TopK &t, Cand *buf, int k2, int kk, int lane)
{
    t.best = buf; t.k2 = k2; t.kk = kk; t.staged = 0; t.thr = cand_inf();
    for (int i = lane; i < 2 * k2; i += 32) buf[i] = cand_inf();
    __syncwarp();
}

static __device__ void topk_merge(TopK &t, int lane)
{
    warp_bitonic_sort(t.best, 2 * t.k2, lane);
    for (int i = t.k2 + lane; i < 2 * t.k2; i += 32) t.best[i] = cand_inf();
    __syncwarp();
    t.staged = 0;
    t.thr = t.best[t.kk - 1];
}

// warp-collective: every lane offers one candidate (valid = false for idle lanes)
static __device__ void topk_offer(TopK &t, const Cand &c, bool valid, int lane)
{
    bool pass = valid && cand_less(c, t.thr);
    unsigned mask = __ballot_sync(0xffffffffu, pass);
    if (mask == 0u) return;
    if (t.staged + __popc(mask) > t.k2) {                       // staging full: merge, then re-test
        topk_merge(t, lane);
        pass = pass && cand_less(c, t.thr);
        mask = __ballot_sync(0xffffffffu, pass);
    }
    if (pass) t.best[t.k2 + t.staged + __popc(mask & ((1u << lane) - 1u))] = c;
    t.staged += __popc(mask);
    __syncwarp();
}

static __device__ void topk_finish(TopK &t, int lane)
{
    if (t.staged > 0) topk_merge(t, lane);
}

struct KnnParams {
    GridIndex g;
    const double *unknown;      // [m][2]
    const int32_t *order;       // [count] location ids to process (cell-sorted)
    int64_t count;
    int kk;                     // min(k, n) (row stride of nbr_pos / nbr_d)
    int k2;                     // power of two >= max(kk, 32)
    int r0;                     // first ring radius in cells
    double r_try;               // fast path: radius expected to hold ~1.8 kk known points (the first disc tried)
    int32_t *nbr_pos;           // [count][kk]
    double *nbr_d;              // [count][kk]
    int32_t *nbr_idx;           // optional [m][k] (original rows, -1 padded), indexed by location id
    int k;
    // directional / range-limited selection (transform/select_points.py:104-257)
    int32_t *nbr_count;         // [count] neighbours found (directional kernel)
    double range, direction;
    int last_tol;               // largest angular tolerance tried (degrees)
    int use_all;
    int stride;                 // row stride of nbr_pos / nbr_d (>= kk)
    int isotropic;              // range kernel without the angular filter (use_all_neighbors_in_range, isotropic model)
    const int32_t *exclude;     // NULL, or [m]: row of `known` that location q must not use (leave-one-out), -1 = none
};

static __device__ __forceinline__ Cand make_cand(const GridIndex &g, int pos, double ux, double uy)
{
    // cdist: sqrt(dx*dx + dy*dy), nothing fused (transform/select_points.py:85)
    Cand c;
    const double dx = __dsub_rn(ux, g.x[pos]), dy = __dsub_rn(uy, g.y[pos]);
    c.d = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    c.orig = g.orig[pos];
    c.pos = pos;
    return c;
}

// E > 0: fast path — when the first search window holds at most 32 * E points, every lane takes E of them
// in registers and the warp sorts all of them at once with a bitonic network (shuffles for strides >= E,
// register compare-exchanges below), instead of filtering batches through the shared-memory top-k (whose
// 16-byte compare-exchanges saturate the shared-memory pipe: ncu showed 98 %).  Falls back to the incremental
// path when the window is fuller (clustered data) or the k-th distance is not yet covered by the window.
// Ascending bitonic sort of 32 * E candidates held E per lane (element index = lane * E + b) by (distance, original row).
// Branch-free compare-exchanges: the comparison is three predicates combined without short-circuit evaluation and the exchange
// is field-wise selects (written as `if (less(o, c)) c = o` with || and && it compiled to ~26 instructions per exchange, two of
// them branches; the direction of a shuffle stage depends on the lane alone, so it is formed once per stage).
template <int E>
static __device__ __forceinline__ void bitonic_sort_regs(Cand (&c)[E], int lane)
{
    auto less_nb = [](const Cand &a, const Cand &b) -> bool {
        const bool lt = a.d < b.d, eq = a.d == b.d, ol = a.orig < b.orig;
        return lt | (eq & ol);
    };
#pragma unroll
    for (int size = 2; size <= 32 * E; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= E) {                        // partner lives in another lane
                const int lane_stride = stride / E;
                // ascending block (bit `size` of the element index clear) and lower element of the pair (bit `stride` clear):
                // both bits lie in the lane part of the index
                const bool up = size >= 32 * E ? true : ((lane & (size / E)) == 0);
                const bool want_small = (((lane & lane_stride) == 0) == up);
#pragma unroll
                for (int b = 0; b < E; ++b) {
                    Cand o;
                    o.d = __shfl_xor_sync(0xffffffffu, c[b].d, lane_stride);
                    o.orig = __shfl_xor_sync(0xffffffffu, c[b].orig, lane_stride);
                    o.pos = __shfl_xor_sync(0xffffffffu, c[b].pos, lane_stride);
                    // keep the smaller one if I am the lower element of an ascending pair (or the upper of a descending one)
                    const bool take = (less_nb(o, c[b]) == want_small);
                    c[b].d = take ? o.d : c[b].d;
                    c[b].orig = take ? o.orig : c[b].orig;
                    c[b].pos = take ? o.pos : c[b].pos;
                }
            } else {                                  // both elements are mine
#pragma unroll
                for (int b = 0; b < E; ++b) {
                    if ((b & stride) == 0) {
                        const bool up = size >= 32 * E ? true : size >= E ? ((lane & (size / E)) == 0) : ((b & size) == 0);
                        const Cand lo = c[b], hi = c[b + stride];
                        const bool sw = (less_nb(hi, lo) == up);
                        c[b].d = sw ? hi.d : lo.d; c[b].orig = sw ? hi.orig : lo.orig; c[b].pos = sw ? hi.pos : lo.pos;
                        c[b + stride].d = sw ? lo.d : hi.d; c[b + stride].orig = sw ? lo.orig : hi.orig;
                        c[b + stride].pos = sw ? lo.pos : hi.pos;
                    }
                }
            }
        }
    }
}

// the first kk of the sorted candidates are the location's neighbours
template <int E>
static __device__ __forceinline__ void knn_emit(const KnnParams &p, const Cand (&c)[E], int64_t slot, int32_t q, int kk, int lane)
{
#pragma unroll
    for (int b = 0; b < E; ++b) {
        const int t = lane * E + b;
        if (t < kk) {
            p.nbr_pos[slot * p.stride + t] = c[b].pos;
            p.nbr_d[slot * p.stride + t] = c[b].d;
        }
        if (p.nbr_idx && t < p.k) p.nbr_idx[(int64_t)q * p.k + t] = t < kk ? c[b].orig : -1;
    }
    if (p.nbr_idx)
        for (int t = 32 * E + lane; t < p.k; t += 32) p.nbr_idx[(int64_t)q * p.k + t] = -1;
}

template <int E>
__global__ void __launch_bounds__(KNN_WARPS * 32) knn_kernel(const KnnParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t slot = (int64_t)blockIdx.x * KNN_WARPS + warp;
    if (slot >= p.count) return;
    const int k2 = p.k2, kk = p.kk;
    const GridIndex &g = p.g;

    const int32_t q = p.order[slot];
    const double ux = p.unknown[2 * (int64_t)q], uy = p.unknown[2 * (int64_t)q + 1];
    const int cx = clampi((int)floor((ux - g.x0) * g.inv_h), 0, g.gx - 1);
    const int cy = clampi((int)floor((uy - g.y0) * g.inv_h), 0, g.gy - 1);
    const double slack = 1e-9 * (fabs(ux) + fabs(uy) + fabs(g.x0) + fabs(g.y0) + g.h * (g.gx + g.gy));
    const int ex = p.exclude ? p.exclude[q] : -1;               // leave-one-out: this row is not a neighbour

    if constexpr (E > 0) {
        const int r = p.r0;
        const int x_lo = max(cx - r, 0), x_hi = min(cx + r, g.gx - 1);
        const int y_lo = max(cy - r, 0), y_hi = min(cy + r, g.gy - 1);
        const int n_rows = y_hi - y_lo + 1;
        // start / exclusive prefix of each row segment (lane = row)
        int seg0 = 0, len = 0;
        if (lane < n_rows) {
            seg0 = g.cell_start[(y_lo + lane) * g.gx + x_lo];
            len = g.cell_start[(y_lo + lane) * g.gx + x_hi + 1] - seg0;
        }
        int pre = len;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, pre, o);
            if (lane >= o) pre += t;
        }
        const int total = __shfl_sync(0xffffffffu, pre, 31);
        pre -= len;
        if (n_rows <= 32 && total <= 32 * E && total - (ex >= 0 ? 1 : 0) >= kk) {
            Cand c[E];
#pragma unroll
            for (int b = 0; b < E; ++b) {
                const int f = b * 32 + lane;                  // flat index into the concatenated row segments
                c[b] = cand_inf();
                // the row whose [pre, pre + len) holds f
                int pos = -1;
                for (int row = 0; row < n_rows; ++row) {
                    const int rp = __shfl_sync(0xffffffffu, pre, row), rl = __shfl_sync(0xffffffffu, len, row);
                    const int rs = __shfl_sync(0xffffffffu, seg0, row);
                    if (f >= rp && f < rp + rl) pos = rs + (f - rp);
                }
                if (pos >= 0) {
                    c[b] = make_cand(g, pos, ux, uy);
                    if (c[b].orig == ex) c[b] = cand_inf();
                }
            }
            // the radius the scanned window is guaranteed to cover
            double rho = INFINITY;
            if (cx - r > 0) rho = fmin(rho, ux - (g.x0 + (cx - r) * g.h));
            if (cx + r < g.gx - 1) rho = fmin(rho, (g.x0 + (cx + r + 1) * g.h) - ux);
            if (cy - r > 0) rho = fmin(rho, uy - (g.y0 + (cy - r) * g.h));
            if (cy + r < g.gy - 1) rho = fmin(rho, (g.y0 + (cy + r + 1) * g.h) - uy);
            rho -= slack;
            // Only candidates inside a disc that lies in the window can be among the kk nearest, and a disc holding at least kk of
            // them holds all kk: try the radius expected to hold ~1.8 kk points first, then the whole guaranteed disc.  The kept
            // candidates are packed through shared memory and sorted in the smallest register network that takes them (a 64-slot
            // network is 42 compare-exchanges, the 128-slot one 112): k = 32 sorts ~58 instead of ~100 candidates.
            auto count_within = [&](double R) {
                int cnt = 0;
#pragma unroll
                for (int b = 0; b < E; ++b) cnt += __popc(__ballot_sync(0xffffffffu, c[b].d <= R));
                return cnt;
            };
            double R = fmin(rho, p.r_try);
            int cnt = count_within(R);
            if (cnt < kk && R < rho) { R = rho; cnt = count_within(R); }
            if (cnt >= kk) {                                  // (else: the k-th neighbour is not covered yet — incremental path)
                Cand *stage = reinterpret_cast<Cand *>(smem_raw) + (size_t)warp * 2 * k2;       // 2 k2 >= 64 (>= 128 for E = 8) slots
                auto pack = [&](int slots) {
                    int base = 0;
#pragma unroll
                    for (int b = 0; b < E; ++b) {
                        const bool keep = c[b].d <= R;
                        const unsigned m = __ballot_sync(0xffffffffu, keep);
                        if (keep) {
                            PIB_DBG_ASSERT(base + __popc(m & ((1u << lane) - 1u)) < slots);
                            stage[base + __popc(m & ((1u << lane) - 1u))] = c[b];
                        }
                        base += __popc(m);
                    }
                    for (int sl = base + lane; sl < slots; sl += 32) stage[sl] = cand_inf();
                    __syncwarp();
                };
                if (E >= 2 && cnt <= 32) {
                    pack(32);
                    Cand d1[1] = {stage[lane]};
                    __syncwarp();
                    bitonic_sort_regs<1>(d1, lane);
                    knn_emit<1>(p, d1, slot, q, kk, lane);
                } else if (E >= 4 && cnt <= 64) {
                    pack(64);
                    Cand d2[2] = {stage[2 * lane], stage[2 * lane + 1]};
                    __syncwarp();
                    bitonic_sort_regs<2>(d2, lane);
                    knn_emit<2>(p, d2, slot, q, kk, lane);
                } else if (E >= 8 && cnt <= 128) {
                    pack(128);
                    Cand d4[4] = {stage[4 * lane], stage[4 * lane + 1], stage[4 * lane + 2], stage[4 * lane + 3]};
                    __syncwarp();
                    bitonic_sort_regs<4>(d4, lane);
                    knn_emit<4>(p, d4, slot, q, kk, lane);
                } else {
                    bitonic_sort_regs<E>(c, lane);            // (candidates beyond R sort behind the kk nearest)
                    knn_emit<E>(p, c, slot, q, kk, lane);
                }
                return;
            }
        }
    }

    TopK top;
    topk_init(top, reinterpret_cast<Cand *>(smem_raw) + (size_t)warp * 2 * k2, k2, kk, lane);

    int r_done = -1;                // rings [0, r_done] scanned
    int r = p.r0;

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
                    Cand c = cand_inf();
                    if (pos < s1) c = make_cand(g, pos, ux, uy);
                    topk_offer(top, c, pos < s1 && c.orig != ex, lane);
                }
            }
        }
        r_done = r;
        topk_finish(top, lane);
        const bool whole_grid = (x_lo == 0 && y_lo == 0 && x_hi == g.gx - 1 && y_hi == g.gy - 1);
        if (whole_grid) break;
        // radius around the location that the scanned square is guaranteed to cover
        double rho = INFINITY;
        if (cx - r > 0) rho = fmin(rho, ux - (g.x0 + (cx - r) * g.h));
        if (cx + r < g.gx - 1) rho = fmin(rho, (g.x0 + (cx + r + 1) * g.h) - ux);
        if (cy - r > 0) rho = fmin(rho, uy - (g.y0 + (cy - r) * g.h));
        if (cy + r < g.gy - 1) rho = fmin(rho, (g.y0 + (cy + r + 1) * g.h) - uy);
        rho -= slack;
        if (top.thr.d <= rho) break;                             // k-th neighbour is inside the covered disc
        if (isinf(top.thr.d)) r = 2 * r + 1;                      // fewer than k found so far
        else r = max(r + 1, min((int)ceil(top.thr.d * g.inv_h) + 1, 2 * r + 1));
    }

    for (int t = lane; t < kk; t += 32) {
        p.nbr_pos[slot * p.stride + t] = top.best[t].pos;
        p.nbr_d[slot * p.stride + t] = top.best[t].d;
    }
    if (p.nbr_idx)
        for (int t = lane; t < p.k; t += 32) p.nbr_idx[(int64_t)q * p.k + t] = t < kk ? top.best[t].orig : -1;
}

// ---------------------------------------------------------------------------------------------
// Directional neighbour selection (models fitted on a directional variogram carry a direction):
//   transform/select_points.py:104-168  select_kriging_data_from_direction
//   transform/select_points.py:171-257  select_possible_neighbors_angular
//   distance/angular.py:106-185         calc_angles, calculate_angular_difference
// Points within `range` are kept, sorted by distance, and filtered by their angular difference to the
// model direction; the tolerance starts at 1 degree and is widened by 1 degree until at least k
// points qualify or it exceeds max_tick.  The result has min(k, qualifying) rows (all qualifying rows
// with use_all), possibly none.  Two passes over the cells that intersect the range disc:
// pass 1 histograms the smallest integer tolerance each in-range point needs, pass 2 selects.
// ---------------------------------------------------------------------------------------------
static __device__ __forceinline__ double py_mod_2pi(double x)
{
    const double two_pi = 2.0 * 3.141592653589793;
    double m = fmod(x, two_pi);
    if (m != 0.0 && m < 0.0) m += two_pi;                       // numpy's floored modulo
    return m;
}

static __device__ __forceinline__ double angular_difference(double ux, double uy, double kx, double ky, double direction)
{
    const double rad2deg = 180.0 / 3.141592653589793, deg2rad = 3.141592653589793 / 180.0;
    const double ang = py_mod_2pi(atan2(ky - uy, kx - ux)) * rad2deg;            // _calc_angle_from_origin
    const double e = direction * deg2rad, r = ang * deg2rad;                       // calculate_angular_difference
    const double a = fabs(py_mod_2pi(r - e) * rad2deg), b = fabs(py_mod_2pi(e - r) * rad2deg);
    return fmin(a, b);
}

constexpr int MAX_TOL = 64;

__global__ void __launch_bounds__(KNN_WARPS * 32) knn_dir_kernel(const KnnParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_hist[KNN_WARPS][MAX_TOL + 2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t slot = (int64_t)blockIdx.x * KNN_WARPS + warp;
    if (slot >= p.count) return;
    const int k2 = p.k2, kk = p.kk;
    const GridIndex &g = p.g;
    const int32_t q = p.order[slot];
    const double ux = p.unknown[2 * (int64_t)q], uy = p.unknown[2 * (int64_t)q + 1];
    // cells that can hold a point within the range (one extra cell of margin for rounding)
    const int x_lo = clampi((int)floor((ux - p.range - g.x0) * g.inv_h) - 1, 0, g.gx - 1);
    const int x_hi = clampi((int)floor((ux + p.range - g.x0) * g.inv_h) + 1, 0, g.gx - 1);
    const int y_lo = clampi((int)floor((uy - p.range - g.y0) * g.inv_h) - 1, 0, g.gy - 1);
    const int y_hi = clampi((int)floor((uy + p.range - g.y0) * g.inv_h) + 1, 0, g.gy - 1);
    const int ex = p.exclude ? p.exclude[q] : -1;
    int *hist = s_hist[warp];
    for (int t = lane; t < MAX_TOL + 2; t += 32) hist[t] = 0;
    __syncwarp();

    TopK top;
    topk_init(top, reinterpret_cast<Cand *>(smem_raw) + (size_t)warp * 2 * k2, k2, kk, lane);
    int tol = p.last_tol;
    for (int pass = 0; pass < 2; ++pass) {
        for (int yy = y_lo; yy <= y_hi; ++yy) {
            const int s0 = g.cell_start[yy * g.gx + x_lo], s1 = g.cell_start[yy * g.gx + x_hi + 1];
            for (int base = s0; base < s1; base += 32) {
                const int pos = base + lane;
                Cand c = cand_inf();
                bool ok = false;
                double ad = 0.0;
                if (pos < s1) {
                    c = make_cand(g, pos, ux, uy);
                    if (c.d <= p.range && c.orig != ex) {        // distances <= max_range (select_points.py:201 / :93)
                        ad = p.isotropic ? 0.0 : angular_difference(ux, uy, g.x[pos], g.y[pos], p.direction);
                        ok = true;
                    }
                }
                if (pass == 0) {
                    if (ok) {                                    // smallest integer tolerance with ad <= tolerance
                        int need = max(1, (int)ceil(ad));
                        if ((double)need < ad) ++need;
                        if (need <= p.last_tol) atomicAdd(&hist[need], 1);
                    }
                } else {
                    topk_offer(top, c, ok && ad <= (double)tol, lane);
                }
            }
        }
        __syncwarp();
        if (pass == 0) {
            // widen 1, 2, ... until >= k qualify or the tolerance passes max_tick (select_points.py:232-240)
            int cum = 0;
            tol = p.last_tol;
            for (int t = 1; t <= p.last_tol; ++t) {
                cum += hist[t];
                if (cum >= p.k) { tol = t; break; }
            }
        }
    }
    topk_finish(top, lane);
    int found = 0;
    for (int t = lane; t < kk; t += 32) {
        const Cand c = top.best[t];
        found += (c.pos >= 0) ? 1 : 0;
    }
    for (int o = 16; o; o >>= 1) found += __shfl_xor_sync(0xffffffffu, found, o);
    if (p.isotropic) {
        // select_points.py:95-101: with at least k points in range ALL of them are used; otherwise the k nearest
        // overall, which knn_kernel already wrote for this location — leave them
        int in_range = 0;
        for (int t = 1; t <= p.last_tol; ++t) in_range += hist[t];
        if (in_range < p.k) {
            if (lane == 0) p.nbr_count[slot] = min(p.k, (int)g.n - (ex >= 0 ? 1 : 0));
            return;
        }
    }
    for (int t = lane; t < kk; t += 32) {
        const Cand c = top.best[t];
        p.nbr_pos[slot * p.stride + t] = c.pos;
        p.nbr_d[slot * p.stride + t] = c.d;
    }
    if (p.use_all) {                                             // all qualifying rows are wanted: did they fit?
        int qualifying = 0;
        for (int t = 1; t <= tol; ++t) qualifying += hist[t];
        if (qualifying > kk) found = -1;
        else if (qualifying < p.k) found = min(found, qualifying);
    } else found = min(found, p.k);
    if (lane == 0) p.nbr_count[slot] = found;
    if (p.nbr_idx)
        for (int t = lane; t < p.k; t += 32) p.nbr_idx[(int64_t)q * p.k + t] = (t < kk && top.best[t].pos >= 0) ? top.best[t].orig : -1;
}

// ---------------------------------------------------------------------------------------------
// gamma(h): semivariogram/theoretical/variogram_models/models.py:41-440 (sill = partial sill).
// h == 0 gives the nugget for every model (what _get_zero_lag_value + the formulas evaluate to).
// ---------------------------------------------------------------------------------------------
// a / b with y = RN(1 / b) at hand: Markstein's correction step, q' = RN(q + (a - b q) y) with q = RN(a y) and the residual
// exact in an FMA, IS the correctly rounded quotient (3 instructions instead of the ~25 of a full fp64 division; compared
// with a / b on 2e8 random pairs: no difference).  y == 0 asks for the plain division (parameters near the ends of the
// exponent range, where y or the residual would over- / underflow).
static __device__ __forceinline__ double div_by(double a, double b, double y)
{
    if (y == 0.0) return a / b;
    const double q = __dmul_rn(a, y);
    return __fma_rn(__fma_rn(-q, b, a), y, q);
}
// the reciprocal div_by may use for a divisor b, or 0
static __device__ __forceinline__ double safe_rcp(double b)
{
    return (fabs(b) > 1e-100 && fabs(b) < 1e100) ? __drcp_rn(b) : 0.0;
}

// rinv = safe_rcp(rang), rinv2 = safe_rcp(rang * rang) (or 0, 0): the quotients h / rang and (h h) / (rang rang) are the same bits either way
static __device__ __forceinline__ double gamma_model(int model, double h, double nugget, double sill, double rang,
                                                     double rinv = 0.0, double rinv2 = 0.0)
{
    if (h == 0.0) return nugget;
    const double ar = div_by(h, rang, rinv);
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
    case PIB_MODEL_GAUSSIAN: return nugget + sill * (1.0 - exp(-div_by(h * h, rang * rang, rinv2)));
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
    const int32_t *nbr_count;   // NULL, or [count] neighbours actually selected (<= kk; directional selection)
    int only_status;            // 0: every system; else only the (set, location) entries whose status byte equals it
    size_t mode_stride;         // PIB_KRIGE_BOTH: offset (in systems) of the simple-kriging block behind the ordinary one
    int fast_div;               // 1: quotients by the model range through div_by (same bits, 3 instructions); 0 (PIB_FAST_DIV=0): plain divisions
    int *singular_flag;         // NULL, or a device word set by every solver that marks a system PIB_STATUS_SINGULAR (singular_kernel
                                // leaves at once while it is 0)
    int *retry_flag;            // NULL, or a device word the Gauss-Jordan kernels set when they mark a system PIB_STATUS_RETRY:
                                // the bordered pass behind them (only_status == PIB_STATUS_RETRY) leaves at once while it is 0
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
    if (p.only_status && p.retry_flag && *p.retry_flag == 0) return;            // nothing was marked for this pass
    const int kmax = p.kk, mode = p.mode;                       // shared-memory layout is sized for kmax neighbours
    const int dmax = (mode == PIB_KRIGE_ORDINARY) ? kmax + 1 : kmax;
    const int st = p.stride;
    const size_t per_warp = ((size_t)dmax * st * 8 + (size_t)dmax * 16 + (size_t)kmax * 24 + (size_t)((kmax + 1) / 2 * 2) * 4 + 15) / 16 * 16;
    double *A = reinterpret_cast<double *>(smem_raw + per_warp * warp);
    double *rhs0 = A + (size_t)dmax * st;
    double *dinv = rhs0 + dmax;
    double *nx = dinv + dmax, *ny = nx + kmax, *nd = ny + kmax;
    int32_t *npos = reinterpret_cast<int32_t *>(nd + kmax);
    const GridIndex &g = p.g;

    const int32_t q = p.order[slot];
    const double ux = p.unknown[2 * (int64_t)q], uy = p.unknown[2 * (int64_t)q + 1];
    // a non-finite location has no neighbours (every candidate distance is NaN): same NaN row as an empty selection
    const int kk = !(isfinite(ux) && isfinite(uy)) ? 0 : p.nbr_count ? p.nbr_count[slot] : kmax;      // neighbours of THIS location
    const int dim = (mode == PIB_KRIGE_ORDINARY) ? kk + 1 : kk;
    if (kk <= 0) {                       // 0: no neighbour qualified, the reference kriges a NaN row; -1: too many
        if (lane == 0)
            for (int set = 0; set < p.n_sets; ++set) {
                double *o = p.out + ((size_t)set * p.m + q) * 4;
                // the reference solves a 1-row system built from a NaN row: zhat = NaN and sigma^2 = gamma(NaN),
                // i.e. nugget + sill for the models written with `lags <= rang` masks, NaN for the exponential ones
                o[0] = nan(""); o[2] = ux; o[3] = uy;
                o[1] = kk < 0 ? nan("") : gamma_model(p.model, nan(""), p.params[3 * set], p.params[3 * set + 1], p.params[3 * set + 2]);
                if (p.status) p.status[(size_t)set * p.m + q] = kk < 0 ? PIB_STATUS_TOO_MANY_NEIGHBORS : PIB_STATUS_OK;
            }
        return;
    }
    for (int t = lane; t < kk; t += 32) {
        const int pos = p.nbr_pos[slot * kmax + t];
        PIB_DBG_ASSERT(pos >= 0 && pos < g.n);
        npos[t] = pos; nx[t] = g.x[pos]; ny[t] = g.y[pos]; nd[t] = p.nbr_d[slot * kmax + t];
    }
    __syncwarp();

    for (int set = 0; set < p.n_sets; ++set) {
        if (p.only_status && p.status[(size_t)set * p.m + q] != p.only_status) continue;     // warp-uniform
        const double nugget = p.params[3 * set], sill = p.params[3 * set + 1], rang = p.params[3 * set + 2];
        // ---- assemble [Gamma 1; 1^T 0 | k 1] in the reference's row order (ordinary.py:108-115) ----
        bool dup = false;
        for (int i = 0; i < kk; ++i) {
            for (int j = i + 1 + lane; j < kk; j += 32) {
                const double dx = __dsub_rn(nx[i], nx[j]), dy = __dsub_rn(ny[i], ny[j]);
                const double h = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
                dup |= (h == 0.0);
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
        // duplicated coordinates among the neighbours = two identical rows: always singular (see solve_reg_kernel)
        bool singular = __any_sync(0xffffffffu, dup);
        for (int j = 0; j < dim && !singular; ++j) {
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
            if (stat == PIB_STATUS_SINGULAR && p.singular_flag) *p.singular_flag = 1;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// Register-resident solver (k <= 64): the system never touches shared memory during elimination.
//
// solve_kernel above keeps the matrix in shared memory, so every DFMA of the trailing update is an
// LDS + STS and the kernel is shared-memory bound (2-3 % of the FP64 peak).  Here ONE THREAD OWNS ONE
// ROW of Gamma in registers (K doubles + border column + right-hand side); K = 8/16/32/64 is a template
// parameter and the whole LU is unrolled, so every register index is static.  Per elimination step
// the threads of a system agree on the pivot (partial pivoting = argmax over the rows not yet used,
// like dgetf2), the pivot row is broadcast through shared memory (one LDS.128 per two DFMAs) and every
// other row updates itself.  Rows are never swapped: a row simply retires when it is chosen as pivot,
// which is the same elimination order as LAPACK's explicit swaps.  The ones border of ordinary kriging
// is the (K+1)-th row; it lives in shared memory, one element per thread (thread c owns column c).
// Systems with fewer than K neighbours are padded with identity rows / columns.
// ---------------------------------------------------------------------------------------------
template <int TPS>
static __device__ __forceinline__ void sys_sync(int sid)
{
    if (TPS <= 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(sid + 1), "n"(TPS) : "memory");
}

template <int K, bool ORD>
struct RegSysSmem {
    double prow[K + 4];        // pivot row: columns 0..K-1, [K] border column, [K+1] right-hand side
    double brow[K + 4];        // border row (ORD): columns 0..K-1, [K] corner, [K+1] its right-hand side
    double xs[K + 4];          // solution
    double2 nxy[K];
    double G[K * (K + 1)];     // Gamma, row stride K + 1 (odd: column reads are conflict free); each entry computed once
    double redv[4];
    long long redk[2];
    int redi[4];
    int dup;                   // a neighbour pair at distance 0 (duplicated coordinates): the system is singular
};

// Threads per CTA / CTAs per SM.  The kernel is latency bound (a chain of dependent elimination steps per system),
// so what matters is how many systems are in flight per SM.  ptxas takes every register the launch bounds allow
// (240 at 256 threads for K = 64) although the row itself needs 128: capping it at 168 (384 threads) costs a
// 48-byte spill and puts 6 systems on an SM instead of 4; K = 32 is fastest as 2 x 256 threads at 128 registers (2 x 320 at 96 spills and is 12 % slower).
#ifndef PIB_SOLVE64_THREADS
#define PIB_SOLVE64_THREADS 384
#endif
#ifndef PIB_SOLVE32_THREADS
#define PIB_SOLVE32_THREADS 256
#endif
template <int K>
__host__ __device__ constexpr int solve_reg_threads() { return K == 64 ? PIB_SOLVE64_THREADS : K == 32 ? PIB_SOLVE32_THREADS : 256; }
template <int K>
__host__ __device__ constexpr int solve_reg_min_ctas() { return K == 32 ? 2 : 1; }

template <int K, bool ORD>
__global__ void __launch_bounds__(solve_reg_threads<K>(), solve_reg_min_ctas<K>()) solve_reg_kernel(const SolveParams p)
{
    constexpr int TPS = K < 32 ? 32 : K;          // threads per system
    constexpr int SPC = solve_reg_threads<K>() / TPS;   // systems per CTA
    constexpr int NSTEP = ORD ? K + 1 : K;
    constexpr int WARPS = TPS / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RegSysSmem<K, ORD> *smem = reinterpret_cast<RegSysSmem<K, ORD> *>(smem_raw);
    const int sid = threadIdx.x / TPS, t = threadIdx.x % TPS, lane = threadIdx.x & 31, wis = t >> 5;
    const int64_t slot = (int64_t)blockIdx.x * SPC + sid;
    if (slot >= p.count) return;                   // whole systems leave together (barriers are per system)
    if (p.only_status && p.retry_flag && *p.retry_flag == 0) return;            // nothing was marked for this pass
    RegSysSmem<K, ORD> &S = smem[sid];
    const GridIndex &g = p.g;
    const int32_t q = p.order[slot];
    if (p.only_status) {                           // a pass over marked systems: leave before the neighbour gathers (uniform per system)
        bool wanted = false;
        for (int set = 0; set < p.n_sets; ++set) wanted |= p.status[(size_t)set * p.m + q] == p.only_status;
        if (!wanted) return;
    }
    const double ux = p.unknown[2 * (int64_t)q], uy = p.unknown[2 * (int64_t)q + 1];
    // neighbours of THIS location (<= p.kk <= K); a non-finite location has none (every candidate distance is NaN)
    const int kk = !(isfinite(ux) && isfinite(uy)) ? 0 : p.nbr_count ? p.nbr_count[slot] : p.kk;
    if (kk <= 0) {                       // 0: no neighbour qualified, the reference kriges a NaN row; -1: too many
        if (t == 0)
            for (int set = 0; set < p.n_sets; ++set) {
                double *o = p.out + ((size_t)set * p.m + q) * 4;
                // the reference solves a 1-row system built from a NaN row: zhat = NaN and sigma^2 = gamma(NaN),
                // i.e. nugget + sill for the models written with `lags <= rang` masks, NaN for the exponential ones
                o[0] = nan(""); o[2] = ux; o[3] = uy;
                o[1] = kk < 0 ? nan("") : gamma_model(p.model, nan(""), p.params[3 * set], p.params[3 * set + 1], p.params[3 * set + 2]);
                if (p.status) p.status[(size_t)set * p.m + q] = kk < 0 ? PIB_STATUS_TOO_MANY_NEIGHBORS : PIB_STATUS_OK;
            }
        return;
    }
    int my_pos = -1;
    double my_d = 0.0;
    if (t < kk) {
        my_pos = p.nbr_pos[slot * p.kk + t];
        PIB_DBG_ASSERT(my_pos >= 0 && my_pos < g.n);
        my_d = p.nbr_d[slot * p.kk + t];
        S.nxy[t] = make_double2(g.x[my_pos], g.y[my_pos]);
    } else if (t < K) {
        S.nxy[t] = make_double2(0.0, 0.0);
    }
    if (t == 0) S.dup = 0;
    sys_sync<TPS>(sid);
    const double2 me = (t < K) ? S.nxy[t] : make_double2(0.0, 0.0);

    for (int set = 0; set < p.n_sets; ++set) {
        if (p.only_status && p.status[(size_t)set * p.m + q] != p.only_status) continue;     // uniform over the system
        const double nugget = p.params[3 * set], sill = p.params[3 * set + 1], rang = p.params[3 * set + 2];
        // ---- my row of [Gamma | 1 | k] (ordinary.py:108-115 row order: neighbours by ascending distance) ----
        // Gamma is symmetric: thread t evaluates the K/2 entries (t, t+1 .. t+K/2) cyclically and writes both
        // (t, c) and (c, t) into shared memory; afterwards every thread loads its row with static indices
        if (t < K) {
            for (int i = 0; i < K / 2; ++i) {
                int c = t + 1 + i;
                if (c >= K) c -= K;
                double v;
                if (t < kk && c < kk) {
                    const double2 o = S.nxy[c];
                    const double dx = __dsub_rn(me.x, o.x), dy = __dsub_rn(me.y, o.y);
                    const double h = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
                    if (h == 0.0) S.dup = 1;
                    v = gamma_model(p.model, h, nugget, sill, rang);
                } else v = 0.0;                                                 // identity padding (off diagonal)
                S.G[t * (K + 1) + c] = v;
                S.G[c * (K + 1) + t] = v;
            }
            S.G[t * (K + 1) + t] = (t < kk) ? nugget : 1.0;                     // gamma(0) / identity padding
        }
        sys_sync<TPS>(sid);
        double a[K];
#pragma unroll
        for (int c = 0; c < K; ++c) a[c] = (t < K) ? S.G[t * (K + 1) + c] : 0.0;
        const double rhs0 = (t < kk) ? gamma_model(p.model, my_d, nugget, sill, rang) : 0.0;
        double rhs = rhs0;
        double ab = (ORD && t < kk) ? 1.0 : 0.0;                                // border column entry of my row
        if (ORD) {
            if (t < K) S.brow[t] = (t < kk) ? 1.0 : 0.0;
            if (t == 0) { S.brow[K] = 0.0; S.brow[K + 1] = 1.0; }
        }
        bool row_active = (t < K);
        bool border_active = ORD;
        int my_step = -1, border_step = -1;
        double my_inv = 0.0, border_inv = 0.0;
        // two identical rows: singular in exact arithmetic.  LAPACK only notices when rounding happens to leave an
        // exact zero pivot (about half of such systems with OpenBLAS, garbage for the rest); here it is always caught
        bool singular = S.dup != 0;
        sys_sync<TPS>(sid);

        // ---- LU with partial pivoting, fully unrolled (column j) ------------------------------------
#pragma unroll
        for (int j = 0; j < NSTEP; ++j) {
            if (singular) break;
            // candidates: my row's entry in column j, and (ORD) the border row's entry offered by its owner
            double bv = row_active ? fabs(j < K ? a[j < K ? j : 0] : ab) : -1.0;
            int bi = t;
            if (ORD && border_active) {
                const int owner = (j < K) ? j : 0;
                if (t == owner) {
                    const double v = fabs(S.brow[j]);
                    if (v > bv) { bv = v; bi = K; }                             // ties keep the smaller row id
                }
            }
            {
                // warp argmax with the hardware integer reductions: |v| >= 0, so the fp64 order is the order of
                // the bit pattern; first the high words, then the low words among the leaders, then the lowest lane
                // (NaN candidates have the largest pattern and are caught by the check below)
                const long long key = __double_as_longlong(bv < 0.0 ? -1.0 : bv);
                const unsigned hi = (bv < 0.0) ? 0u : (unsigned)(key >> 32) + 1u, lo = (unsigned)key;
                const unsigned hmax = __reduce_max_sync(0xffffffffu, hi);
                const unsigned lmax = __reduce_max_sync(0xffffffffu, hi == hmax ? lo : 0u);
                const unsigned lead = __ballot_sync(0xffffffffu, hi == hmax && lo == lmax);
                const int src = __ffs(lead) - 1;
                bi = __shfl_sync(0xffffffffu, bi, src);
                bv = (hmax == 0u) ? -1.0 : __longlong_as_double(((long long)(hmax - 1u) << 32) | lmax);
            }
            if (WARPS > 1) {
                if (lane == 0) { S.redv[wis] = bv; S.redi[wis] = bi; }
                sys_sync<TPS>(sid);
#pragma unroll
                for (int w = 0; w < WARPS; ++w) {
                    const double ov = S.redv[w];
                    const int oi = S.redi[w];
                    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                }
            }
            if (!(bv > 0.0) || isinf(bv)) { singular = true; break; }           // zero / NaN pivot column
            // publish the pivot row
            if (bi < K) {
                if (t == bi) {
#pragma unroll
                    for (int c = 0; c < K; ++c)
                        if (c >= j) S.prow[c] = a[c];
                    S.prow[K] = ab; S.prow[K + 1] = rhs;
                    // two-warp systems: the owner computes the reciprocal once, ahead of the barrier, instead of
                    // all 64 threads after it (it sits on every thread's dependent chain)
                    if (WARPS > 1) S.prow[K + 2] = __drcp_rn(j < K ? a[j < K ? j : 0] : ab);
                    row_active = false; my_step = j;
                }
            } else {
                if (t < K) S.prow[t] = S.brow[t];
                if (t == 0) {
                    S.prow[K] = S.brow[K]; S.prow[K + 1] = S.brow[K + 1];
                    if (WARPS > 1) S.prow[K + 2] = __drcp_rn(S.brow[j]);
                }
                border_active = false; border_step = j;
            }
            sys_sync<TPS>(sid);
            const double inv = WARPS > 1 ? S.prow[K + 2] : __drcp_rn(S.prow[j]);
            if (my_step == j) my_inv = inv;
            if (border_step == j) border_inv = inv;
            {
                // retired rows update with l = 0 (an exact no-op): no divergent branch around the DFMAs
                const double l = row_active ? (j < K ? a[j < K ? j : 0] : ab) * inv : 0.0;   // dgetf2 scales by the reciprocal
#pragma unroll
                for (int c = 0; c < K; ++c)
                    if (c > j) a[c] = __fma_rn(-l, S.prow[c], a[c]);
                if (ORD && j < K) ab = __fma_rn(-l, S.prow[K], ab);
                rhs = __fma_rn(-l, S.prow[K + 1], rhs);
            }
            if (ORD && border_active) {
                const double lb = S.brow[j] * inv;
                if (t < K && t > j) S.brow[t] = __fma_rn(-lb, S.prow[t], S.brow[t]);
                if (t == 0) {
                    if (j < K) S.brow[K] = __fma_rn(-lb, S.prow[K], S.brow[K]);
                    S.brow[K + 1] = __fma_rn(-lb, S.prow[K + 1], S.brow[K + 1]);
                }
            }
            if (WARPS == 1) __syncwarp();
        }
        sys_sync<TPS>(sid);

        double zhat = nan(""), sigma = nan("");
        uint8_t stat = PIB_STATUS_SINGULAR;
        if (!singular) {
            // ---- back substitution: column j of U belongs to the rows that retired before step j ----
#pragma unroll
            for (int j = NSTEP - 1; j >= 0; --j) {
                if (my_step == j) S.xs[j] = rhs * my_inv;
                if (ORD && border_step == j && t == 0) S.xs[j] = S.brow[K + 1] * border_inv;
                sys_sync<TPS>(sid);
                const double x = S.xs[j];
                if (my_step >= 0 && my_step < j) rhs = __fma_rn(-(j < K ? a[j < K ? j : 0] : ab), x, rhs);
                if (ORD && border_step >= 0 && border_step < j && t == 0)
                    S.brow[K + 1] = __fma_rn(-S.brow[j], x, S.brow[K + 1]);
            }
            sys_sync<TPS>(sid);
            // zhat = z . w (ordinary.py:121) or (z - mean) . w + mean (simple.py:114-116); sigma = w . k (+ mu)
            double zs = 0.0, ss = 0.0;
            if (t < kk) {
                const double w = S.xs[t];
                double zv = p.values ? p.values[(size_t)set * p.n + g.orig[my_pos]] : g.z[my_pos];
                if (!ORD) zv -= p.process_mean;
                zs = zv * w;
                ss = w * rhs0;
            }
            if (ORD && t == 0) ss += S.xs[K];                                   // mu * 1
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                zs += __shfl_xor_sync(0xffffffffu, zs, o);
                ss += __shfl_xor_sync(0xffffffffu, ss, o);
            }
            if (WARPS > 1) {
                if (lane == 0) { S.redv[wis] = zs; S.redv[2 + wis] = ss; }
                sys_sync<TPS>(sid);
                zs = S.redv[0] + S.redv[1]; ss = S.redv[2] + S.redv[3];
            }
            zhat = ORD ? zs : zs + p.process_mean;
            sigma = ss;
            stat = PIB_STATUS_OK;
            if (sigma < 0.0) { sigma = nan(""); stat = PIB_STATUS_NEGATIVE_VARIANCE; }   // ordinary.py:125-126
        }
        if (t == 0) {
            double *o = p.out + ((size_t)set * p.m + q) * 4;
            o[0] = zhat; o[1] = sigma; o[2] = ux; o[3] = uy;
            if (p.status) p.status[(size_t)set * p.m + q] = stat;
            if (stat == PIB_STATUS_SINGULAR && p.singular_flag) *p.singular_flag = 1;
        }
        sys_sync<TPS>(sid);
    }
}

template <int K>
static int launch_solve_reg(const SolveParams &sp, int mode, cudaStream_t stream)
{
    constexpr int NT = solve_reg_threads<K>();
    constexpr int SPC = NT / (K < 32 ? 32 : K);
    const int grid = (int)((sp.count + SPC - 1) / SPC);
    if (mode == PIB_KRIGE_ORDINARY) {
        const size_t bytes = sizeof(RegSysSmem<K, true>) * SPC;
        PIB_CUDA(cudaFuncSetAttribute(solve_reg_kernel<K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        solve_reg_kernel<K, true><<<grid, NT, bytes, stream>>>(sp);
    } else {
        const size_t bytes = sizeof(RegSysSmem<K, false>) * SPC;
        PIB_CUDA(cudaFuncSetAttribute(solve_reg_kernel<K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        solve_reg_kernel<K, false><<<grid, NT, bytes, stream>>>(sp);
    }
    PIB_CUDA(cudaGetLastError());
    return PIB_OK;
}

// ---------------------------------------------------------------------------------------------
// Gauss-Jordan solver on Gamma alone (k <= 64) — the default.
//
// solve_reg_kernel spends ~250 warp instructions per elimination step, 200 of them bookkeeping on one dependent
// chain (two barriers, the ones border kept in shared memory, reciprocal and pivot exchange after the barrier), and
// needs a 65-step back substitution: 5.5 % of the FP64 peak at k = 64 (profiles/r1_kriging_kernels_ncu_full.md).
// Here the same data layout (ONE THREAD OWNS ONE ROW in registers, fully unrolled) runs a leaner recurrence:
//   * Gamma only, two right-hand sides.  The ordinary system [Gamma 1; 1^T 0][w; mu] = [g0; 1] (ordinary.py:108-123) is
//     solved through its Schur complement: Gamma u = g0, Gamma v = 1, mu = (1^T u - 1) / (1^T v), w = u - mu v.
//     Simple kriging (simple.py:102-118) IS Gamma w = g0.  One elimination therefore serves ordinary kriging, simple
//     kriging, or both at once (PIB_KRIGE_BOTH: BASELINE config 4 asks for both).
//   * Gauss-Jordan with partial pivoting: retired rows keep eliminating, so the solution is rhs / pivot per row and
//     there is no back substitution.  In this layout that is free — the retired rows executed the DFMAs of every
//     step anyway (with a zero multiplier, to avoid divergence).
//   * ONE barrier per column: every warp publishes its own best candidate row (with |pivot| and the reciprocal,
//     computed by the row's owner BEFORE the barrier) into a double-buffered slot; after the barrier every thread
//     picks the winner from two loads.
// The pivot order is dgetf2's (largest |a|, lowest row on ties) applied to Gamma; results agree with the reference's
// bordered LU to rounding (checked against the fixtures at 1e-9).  Systems this route cannot decide — Gamma singular
// although the bordered matrix may not be (k = 1 with a zero nugget), 1^T v = 0 — are marked PIB_STATUS_RETRY and
// re-solved by solve_reg_kernel on the bordered matrix.
// ---------------------------------------------------------------------------------------------
#define PIB_STATUS_RETRY 255

template <int K>
struct GjSmem {
    static constexpr int W = (K < 32 ? 32 : K) / 32;
    static constexpr int CS = K + 6;                       // candidate slot: row[K], rhs0, rhs1, |pivot|, 1/pivot, row id
    // Gamma (row stride K + 2: 16-byte aligned rows, conflict-free LDS.128 by row) is only needed until every thread
    // has its row in registers; the candidate slots reuse the same memory afterwards
    double G[K * (K + 2) > 2 * W * CS ? K * (K + 2) : 2 * W * CS];
    double2 nxy[K];
    double g0[K];                                          // gamma(distance to the location) per neighbour
    double zv[K];                                          // neighbour values of the current set
    double red[4 * W];
    int dup;
};

#ifndef PIB_GJ_LOOKAHEAD
#define PIB_GJ_LOOKAHEAD 1
#endif
#ifndef PIB_GJ32_THREADS
#define PIB_GJ32_THREADS 256
#endif
#ifndef PIB_GJ32_CTAS
#define PIB_GJ32_CTAS 2
#endif
#ifndef PIB_GJ64_THREADS
#define PIB_GJ64_THREADS 384
#endif
template <int K>
__host__ __device__ constexpr int gj_threads() { return K == 64 ? PIB_GJ64_THREADS : K == 32 ? PIB_GJ32_THREADS : 256; }
template <int K>
__host__ __device__ constexpr int gj_min_ctas() { return K == 32 ? PIB_GJ32_CTAS : 1; }

template <int K, int MODE>      // MODE: PIB_KRIGE_ORDINARY, PIB_KRIGE_SIMPLE or PIB_KRIGE_BOTH
__global__ void __launch_bounds__(gj_threads<K>(), gj_min_ctas<K>()) solve_gj_kernel(const SolveParams p)
{
    constexpr int TPS = K < 32 ? 32 : K;          // threads per system
    constexpr int SPC = gj_threads<K>() / TPS;    // systems per CTA
    constexpr int WARPS = TPS / 32;
    constexpr int CS = GjSmem<K>::CS;
    constexpr bool WANT_OK = MODE != PIB_KRIGE_SIMPLE, WANT_SK = MODE != PIB_KRIGE_ORDINARY;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GjSmem<K> *smem = reinterpret_cast<GjSmem<K> *>(smem_raw);
    const int sid = threadIdx.x / TPS, t = threadIdx.x % TPS, lane = threadIdx.x & 31, wis = t >> 5;
    const int64_t slot = (int64_t)blockIdx.x * SPC + sid;
    if (slot >= p.count) return;                   // whole systems leave together (barriers are per system)
    GjSmem<K> &S = smem[sid];
    const GridIndex &g = p.g;
    const int32_t q = p.order[slot];
    const double ux = p.unknown[2 * (int64_t)q], uy = p.unknown[2 * (int64_t)q + 1];
    // neighbours of THIS location (<= p.kk <= K); a non-finite location has none (every candidate distance is NaN)
    const int kk = !(isfinite(ux) && isfinite(uy)) ? 0 : p.nbr_count ? p.nbr_count[slot] : p.kk;
    const size_t sk_off = MODE == PIB_KRIGE_BOTH ? p.mode_stride : 0;       // where the simple-kriging rows go
    if (kk <= 1) {
        // 0 / -1: no (too many) neighbours; 1: Gamma = [gamma(0)] may be singular although the bordered matrix is not —
        // all of them go to the bordered solver
        if (t == 0)
            for (int set = 0; set < p.n_sets; ++set) {
                if (WANT_OK) p.status[(size_t)set * p.m + q] = PIB_STATUS_RETRY;
                if (WANT_SK) p.status[sk_off + (size_t)set * p.m + q] = PIB_STATUS_RETRY;
                if (p.retry_flag) *p.retry_flag = 1;
            }
        return;
    }
    int my_pos = -1;
    double my_d = 0.0;
    if (t < kk) {
        my_pos = p.nbr_pos[slot * p.kk + t];
        PIB_DBG_ASSERT(my_pos >= 0 && my_pos < g.n);
        my_d = p.nbr_d[slot * p.kk + t];
        S.nxy[t] = make_double2(g.x[my_pos], g.y[my_pos]);
    } else if (t < K) {
        S.nxy[t] = make_double2(0.0, 0.0);
    }
    if (t == 0) S.dup = 0;
    sys_sync<TPS>(sid);
    const double2 me = (t < K) ? S.nxy[t] : make_double2(0.0, 0.0);

    for (int set = 0; set < p.n_sets; ++set) {
        const double nugget = p.params[3 * set], sill = p.params[3 * set + 1], rang = p.params[3 * set + 2];
        const double rinv = p.fast_div ? safe_rcp(rang) : 0.0, rinv2 = p.fast_div ? safe_rcp(rang * rang) : 0.0;
        // ---- Gamma, every entry evaluated once: thread t takes (t, t+1 .. t+K/2) cyclically and writes both halves.
        // Four entries per iteration: one evaluation is a chain of ~100 dependent instructions (sqrt, division, model) ----
        if (t < K) {
            constexpr int AU = K >= 8 ? 4 : 1;
#pragma unroll 1
            for (int i0 = 0; i0 < K / 2; i0 += AU) {
                int cc[AU];
                double hh[AU], vv[AU];
#pragma unroll
                for (int u = 0; u < AU; ++u) {
                    int c = t + 1 + i0 + u;
                    if (c >= K) c -= K;
                    cc[u] = c;
                    const double2 o = S.nxy[c];
                    const double dx = __dsub_rn(me.x, o.x), dy = __dsub_rn(me.y, o.y);
                    hh[u] = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
                }
#pragma unroll
                for (int u = 0; u < AU; ++u) {
                    const bool real = (t < kk && cc[u] < kk);
                    if (real && hh[u] == 0.0) S.dup = 1;
                    vv[u] = real ? gamma_model(p.model, hh[u], nugget, sill, rang, rinv, rinv2) : 0.0;       // identity padding (off diagonal)
                }
#pragma unroll
                for (int u = 0; u < AU; ++u) {
                    S.G[t * (K + 2) + cc[u]] = vv[u];
                    S.G[cc[u] * (K + 2) + t] = vv[u];
                }
            }
            S.G[t * (K + 2) + t] = (t < kk) ? nugget : 1.0;                     // gamma(0) / identity padding
            const double kv = (t < kk) ? gamma_model(p.model, my_d, nugget, sill, rang, rinv, rinv2) : 0.0;
            S.g0[t] = kv;
            double zv = 0.0;
            if (t < kk) zv = p.values ? p.values[(size_t)set * p.n + g.orig[my_pos]] : g.z[my_pos];
            S.zv[t] = zv;
        }
        sys_sync<TPS>(sid);
        double a[K];
#pragma unroll
        for (int c = 0; c < K; c += 2) {
            const double2 v = (t < K) ? *reinterpret_cast<const double2 *>(&S.G[t * (K + 2) + c]) : make_double2(0.0, 0.0);
            a[c] = v.x; a[c + 1] = v.y;
        }
        double r0 = (t < K) ? S.g0[t] : 0.0;                                    // right-hand side g0
        double r1 = (WANT_OK && t < kk) ? 1.0 : 0.0;                            // right-hand side 1 (ordinary kriging)
        bool active = (t < K);
        int my_step = -1;
        double my_inv = 0.0;
        // two identical rows: singular in exact arithmetic.  LAPACK only notices when rounding happens to leave an
        // exact zero pivot (about half of such systems with OpenBLAS, garbage for the rest); here it is always caught
        const bool dup = S.dup != 0;
        bool singular = false;
        sys_sync<TPS>(sid);                                                     // G is dead: the candidate slots take its place
        double *const cand = S.G;

        if (!dup) {
            // Look-ahead: the search key of column j + 1 (and the reciprocal every row would publish as the pivot's) is formed as
            // soon as step j has updated that one column, AHEAD of the step's other K - j - 2 columns, so the warp reduction
            // and the reciprocal's dependent chain run in the shadow of the trailing update instead of heading the next step
            // (PIB_GJ_LOOKAHEAD=0 builds the plain order)
            // |v| >= 0, so the fp64 order is the order of the bit pattern: high words first (inactive rows 0, NaN on top)
            double av = fabs(a[0]);
            unsigned hi = active ? (unsigned)__double2hiint(av) + 1u : 0u;
            unsigned hmax = __reduce_max_sync(0xffffffffu, hi);
            double inv_own = __drcp_rn(a[0]);                                   // dgetf2 scales by the reciprocal of the pivot
#pragma unroll
            for (int j = 0; j < K; ++j) {
                // ---- my warp's candidate for column j: largest |a[j]| among its rows not yet used, lowest row on ties ----
                if (!PIB_GJ_LOOKAHEAD && j > 0) {
                    av = fabs(a[j]);
                    hi = active ? (unsigned)__double2hiint(av) + 1u : 0u;
                    hmax = __reduce_max_sync(0xffffffffu, hi);
                    inv_own = __drcp_rn(a[j]);
                }
                unsigned lead = __ballot_sync(0xffffffffu, hi == hmax);
                if (__popc(lead) > 1 && hmax != 0u) {                           // rare: equal high words, look at the low words
                    const unsigned lo = (hi == hmax) ? (unsigned)__double2loint(av) : 0u;
                    const unsigned lmax = __reduce_max_sync(0xffffffffu, lo);
                    lead = __ballot_sync(0xffffffffu, hi == hmax && lo == lmax);
                }
                const int src = __ffs(lead) - 1;
                double *const C = cand + ((j & 1) * WARPS + wis) * CS;
                if (lane == src) {
#pragma unroll
                    for (int c = 0; c < K; ++c)
                        if (c > j) C[c] = a[c];
                    C[K] = r0; C[K + 1] = r1;
                    C[K + 2] = hmax ? av : -1.0;
                    C[K + 3] = inv_own;
                    reinterpret_cast<int *>(C + K + 4)[0] = t;
                }
                sys_sync<TPS>(sid);
                // ---- the winner among the warps' candidates (rows of warp 0 come first: it wins ties) ----
                const double *P = cand + ((j & 1) * WARPS) * CS;
                double pv = P[K + 2];
                if (WARPS > 1) {
#pragma unroll
                    for (int w = 1; w < WARPS; ++w) {
                        const double *Pw = cand + ((j & 1) * WARPS + w) * CS;
                        const double ov = Pw[K + 2];
                        if (ov > pv || (ov != ov && pv == pv)) { pv = ov; P = Pw; }
                    }
                }
                if (!(pv > 0.0) || isinf(pv)) { singular = true; break; }       // zero / NaN pivot column (uniform over the system)
                const double inv = P[K + 3];
                const bool is_p = (t == reinterpret_cast<const int *>(P + K + 4)[0]);
                if (is_p) { active = false; my_step = j; my_inv = inv; }
                // Gauss-Jordan: EVERY other row eliminates column j, the rows already used included
                const double l = is_p ? 0.0 : a[j] * inv;
                if (PIB_GJ_LOOKAHEAD && j + 1 < K) {
                    a[j + 1] = __fma_rn(-l, P[j + 1], a[j + 1]);
                    av = fabs(a[j + 1]);
                    hi = active ? (unsigned)__double2hiint(av) + 1u : 0u;
                    hmax = __reduce_max_sync(0xffffffffu, hi);
                    inv_own = __drcp_rn(a[j + 1]);
                }
#pragma unroll
                for (int c = 0; c < K; ++c)
                    if (c > j + (PIB_GJ_LOOKAHEAD ? 1 : 0)) a[c] = __fma_rn(-l, P[c], a[c]);
                r0 = __fma_rn(-l, P[K], r0);
                if (WANT_OK) r1 = __fma_rn(-l, P[K + 1], r1);
            }
        }
        sys_sync<TPS>(sid);

        // ---- solution: row t was the pivot of column my_step, so u[my_step] = r0 / pivot, v[my_step] = r1 / pivot ----
        const bool real = (my_step >= 0 && my_step < kk);
        const double u = real ? r0 * my_inv : 0.0, v = real ? r1 * my_inv : 0.0;
        const double zn = real ? S.zv[my_step] : 0.0, gn = real ? S.g0[my_step] : 0.0;
        // sums over the system: 1^T u, 1^T v, u.z, v.z, u.g0, v.g0
        double s[6] = {u, v, u * zn, v * zn, u * gn, v * gn};
#pragma unroll
        for (int o = 16; o; o >>= 1)
#pragma unroll
            for (int i = 0; i < 6; ++i) s[i] += __shfl_xor_sync(0xffffffffu, s[i], o);
        if (WARPS > 1) {
            double *red = cand + 2 * WARPS * CS;                                // behind the candidate slots (G is larger)
            if (lane == 0)
#pragma unroll
                for (int i = 0; i < 6; ++i) red[wis * 6 + i] = s[i];
            sys_sync<TPS>(sid);
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                double acc = red[i];
#pragma unroll
                for (int w = 1; w < WARPS; ++w) acc += red[w * 6 + i];
                s[i] = acc;
            }
        }
        if (t == 0) {
            if (WANT_OK) {
                // mu = (1^T u - 1) / (1^T v), w = u - mu v:  zhat = w.z (ordinary.py:121), sigma^2 = w.g0 + mu (:123)
                double zhat = nan(""), sigma = nan("");
                uint8_t stat = PIB_STATUS_SINGULAR;
                if (!dup) {
                    const double mu = (s[0] - 1.0) / s[1];
                    if (singular || !(fabs(s[1]) > 0.0) || !isfinite(mu)) {                             // decide on the bordered matrix
                        stat = PIB_STATUS_RETRY;
                        if (p.retry_flag) *p.retry_flag = 1;
                    }
                    else {
                        zhat = s[2] - mu * s[3];
                        sigma = (s[4] - mu * s[5]) + mu;
                        stat = PIB_STATUS_OK;
                        if (sigma < 0.0) { sigma = nan(""); stat = PIB_STATUS_NEGATIVE_VARIANCE; }       // ordinary.py:125-126
                    }
                }
                double *o = p.out + ((size_t)set * p.m + q) * 4;
                o[0] = zhat; o[1] = sigma; o[2] = ux; o[3] = uy;
                p.status[(size_t)set * p.m + q] = stat;
                if (stat == PIB_STATUS_SINGULAR && p.singular_flag) *p.singular_flag = 1;
            }
            if (WANT_SK) {
                // w = u:  zhat = (z - mean).w + mean (simple.py:114-116), sigma^2 = w.g0 (:118)
                double zhat = nan(""), sigma = nan("");
                uint8_t stat = PIB_STATUS_SINGULAR;
                if (!dup && !singular) {
                    zhat = (s[2] - p.process_mean * s[0]) + p.process_mean;
                    sigma = s[4];
                    stat = PIB_STATUS_OK;
                    if (sigma < 0.0) { sigma = nan(""); stat = PIB_STATUS_NEGATIVE_VARIANCE; }
                }
                double *o = p.out + (sk_off + (size_t)set * p.m + q) * 4;
                o[0] = zhat; o[1] = sigma; o[2] = ux; o[3] = uy;
                p.status[sk_off + (size_t)set * p.m + q] = stat;
                if (stat == PIB_STATUS_SINGULAR && p.singular_flag) *p.singular_flag = 1;
            }
        }
        sys_sync<TPS>(sid);
    }
}

template <int K, int MODE>
static int launch_solve_gj_m(const SolveParams &sp, cudaStream_t stream)
{
    constexpr int NT = gj_threads<K>();
    constexpr int SPC = NT / (K < 32 ? 32 : K);
    const int grid = (int)((sp.count + SPC - 1) / SPC);
    const size_t bytes = sizeof(GjSmem<K>) * SPC;
    PIB_CUDA(cudaFuncSetAttribute(solve_gj_kernel<K, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    solve_gj_kernel<K, MODE><<<grid, NT, bytes, stream>>>(sp);
    PIB_CUDA(cudaGetLastError());
    return PIB_OK;
}

// ---------------------------------------------------------------------------------------------
// solve_gj2_kernel (k <= 32, the default there): the same Gauss-Jordan recurrence with TWO ROWS PER THREAD, so a system
// takes K / 2 lanes and a warp carries 2 (K = 32), 4 (K = 16) or 8 (K = 8) systems in lockstep.  solve_gj_kernel<32> spends
// ~110 warp instructions per elimination step of ONE system, only ~16 of them DFMAs (ncu: the pivot search, the reciprocal,
// the single-lane publication of the pivot row and its read-back are per step, not per row); here those instructions serve
// two rows per lane and several systems per warp, every pivot-row value read from shared memory feeds two DFMAs, and
// nothing crosses a warp (no named barriers).  Pivot rule, arithmetic per row and the outputs are those of
// solve_gj_kernel; only the order of the final dot-product reduction differs (fp64 rounding).
// Lanes never leave early (the pivot search is a warp-wide butterfly): systems beyond the batch, systems with fewer than
// two neighbours, duplicated neighbours and systems that met a zero pivot ride along masked.
// ---------------------------------------------------------------------------------------------
#ifndef PIB_GJ2_32_THREADS
#define PIB_GJ2_32_THREADS 64
#endif
#ifndef PIB_GJ2_SMALL_CTAS
#define PIB_GJ2_SMALL_CTAS 8
#endif
#ifndef PIB_GJ2_32_CTAS
#define PIB_GJ2_32_CTAS 5
#endif
template <int K>
__host__ __device__ constexpr int gj2_threads() { return K == 32 ? PIB_GJ2_32_THREADS : 64; }
template <int K>
// (K = 16: 8 CTAs of 64 threads at 128 registers — 10 / 12 CTAs spill: 2.44 -> 3.30 / 4.47 ms; K = 8: 12 CTAs at 80 registers, 0.84 -> 0.74 ms)
__host__ __device__ constexpr int gj2_min_ctas() { return K == 32 ? PIB_GJ2_32_CTAS : K == 8 ? 12 : PIB_GJ2_SMALL_CTAS; }

template <int K, int MODE>
__global__ void __launch_bounds__(gj2_threads<K>(), gj2_min_ctas<K>()) solve_gj2_kernel(const SolveParams p)
{
    constexpr int TPS = K / 2;                    // lanes per system: lane t owns rows t and t + TPS
    constexpr int SPC = gj2_threads<K>() / TPS;   // systems per CTA
    constexpr int CS = GjSmem<K>::CS;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr bool WANT_OK = MODE != PIB_KRIGE_SIMPLE, WANT_SK = MODE != PIB_KRIGE_ORDINARY;
    static_assert(K == 8 || K == 16 || K == 32, "two rows per thread: 4, 8 or 16 lanes per system");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GjSmem<K> *smem = reinterpret_cast<GjSmem<K> *>(smem_raw);
    const int sid = threadIdx.x / TPS, t = threadIdx.x % TPS, lane = threadIdx.x & 31;
    const unsigned seg = ((1u << TPS) - 1u) << (lane & ~(TPS - 1));             // the lanes of my system
    const int64_t slot_raw = (int64_t)blockIdx.x * SPC + sid;
    const int64_t slot = min(slot_raw, p.count - 1);                            // (systems beyond the batch redo the last one, unstored)
    GjSmem<K> &S = smem[sid];
    const GridIndex &g = p.g;
    const int32_t q = p.order[slot];
    const double ux = p.unknown[2 * (int64_t)q], uy = p.unknown[2 * (int64_t)q + 1];
    const int kk = !(isfinite(ux) && isfinite(uy)) ? 0 : p.nbr_count ? p.nbr_count[slot] : p.kk;
    const size_t sk_off = MODE == PIB_KRIGE_BOTH ? p.mode_stride : 0;
    const bool store = slot_raw < p.count && kk > 1;
    if (slot_raw < p.count && kk <= 1 && t == 0) {
        // 0 / -1: no (too many) neighbours; 1: Gamma = [gamma(0)] may be singular although the bordered matrix is not
        for (int set = 0; set < p.n_sets; ++set) {
            if (WANT_OK) p.status[(size_t)set * p.m + q] = PIB_STATUS_RETRY;
            if (WANT_SK) p.status[sk_off + (size_t)set * p.m + q] = PIB_STATUS_RETRY;
        }
        if (p.retry_flag) *p.retry_flag = 1;
    }
    int my_pos[2];
    double my_d[2];
    double2 me[2];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
        const int r = t + rr * TPS;
        my_pos[rr] = -1; my_d[rr] = 0.0;
        me[rr] = make_double2(0.0, 0.0);
        if (r < kk) {
            my_pos[rr] = p.nbr_pos[slot * p.kk + r];
            PIB_DBG_ASSERT(my_pos[rr] >= 0 && my_pos[rr] < g.n);
            my_d[rr] = p.nbr_d[slot * p.kk + r];
            me[rr] = make_double2(g.x[my_pos[rr]], g.y[my_pos[rr]]);
        }
        S.nxy[r] = me[rr];
    }
    if (t == 0) S.dup = 0;
    __syncwarp();

    for (int set = 0; set < p.n_sets; ++set) {
        const double nugget = p.params[3 * set], sill = p.params[3 * set + 1], rang = p.params[3 * set + 2];
        const double rinv = p.fast_div ? safe_rcp(rang) : 0.0, rinv2 = p.fast_div ? safe_rcp(rang * rang) : 0.0;
        // ---- Gamma, every entry evaluated once: row r takes the columns r + 1 .. r + K / 2 cyclically, both halves written ----
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int r = t + rr * TPS;
            constexpr int AU = 4;
#pragma unroll 1
            for (int i0 = 0; i0 < K / 2; i0 += AU) {
                int cc[AU];
                double hh[AU], vv[AU];
#pragma unroll
                for (int u = 0; u < AU; ++u) {
                    int c = r + 1 + i0 + u;
                    if (c >= K) c -= K;
                    cc[u] = c;
                    const double2 o = S.nxy[c];
                    const double dx = __dsub_rn(me[rr].x, o.x), dy = __dsub_rn(me[rr].y, o.y);
                    hh[u] = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
                }
#pragma unroll
                for (int u = 0; u < AU; ++u) {
                    const bool real = (r < kk && cc[u] < kk);
                    if (real && hh[u] == 0.0) S.dup = 1;
                    vv[u] = real ? gamma_model(p.model, hh[u], nugget, sill, rang, rinv, rinv2) : 0.0;       // identity padding (off diagonal)
                }
#pragma unroll
                for (int u = 0; u < AU; ++u) {
                    S.G[r * (K + 2) + cc[u]] = vv[u];
                    S.G[cc[u] * (K + 2) + r] = vv[u];
                }
            }
            S.G[r * (K + 2) + r] = (r < kk) ? nugget : 1.0;                     // gamma(0) / identity padding
            S.g0[r] = (r < kk) ? gamma_model(p.model, my_d[rr], nugget, sill, rang, rinv, rinv2) : 0.0;
            double zv = 0.0;
            if (r < kk) zv = p.values ? p.values[(size_t)set * p.n + g.orig[my_pos[rr]]] : g.z[my_pos[rr]];
            S.zv[r] = zv;
        }
        __syncwarp();
        double a[2][K], r0[2], r1[2], my_inv[2];
        int my_step[2];
        bool active[2];
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int r = t + rr * TPS;
#pragma unroll
            for (int c = 0; c < K; c += 2) {
                const double2 v = *reinterpret_cast<const double2 *>(&S.G[r * (K + 2) + c]);
                a[rr][c] = v.x; a[rr][c + 1] = v.y;
            }
            r0[rr] = S.g0[r];                                                   // right-hand side g0
            r1[rr] = (WANT_OK && r < kk) ? 1.0 : 0.0;                           // right-hand side 1 (ordinary kriging)
            active[rr] = true; my_step[rr] = -1; my_inv[rr] = 0.0;
        }
        const bool dup = S.dup != 0;              // two identical rows: singular in exact arithmetic (see solve_gj_kernel)
        bool singular = false;
        __syncwarp();                                                           // G is dead: the candidate slots take its place
        double *const cand = S.G;

#pragma unroll
        for (int j = 0; j < K; ++j) {
            // ---- pivot of column j: largest |a[j]| among my system's rows not yet used, lowest row on ties ----
            const bool go = !(dup || singular);
            const double av0 = fabs(a[0][j]), av1 = fabs(a[1][j]);
            // |v| >= 0, so the fp64 order is the order of the bit pattern: high words first (inactive rows 0, NaN on top)
            const unsigned hi0 = (go && active[0]) ? (unsigned)__double2hiint(av0) + 1u : 0u;
            const unsigned hi1 = (go && active[1]) ? (unsigned)__double2hiint(av1) + 1u : 0u;
            unsigned hmax = max(hi0, hi1);
#pragma unroll
            for (int o = TPS / 2; o; o >>= 1) hmax = max(hmax, __shfl_xor_sync(FULL, hmax, o));
            unsigned lead0 = __ballot_sync(FULL, hi0 == hmax) & seg, lead1 = __ballot_sync(FULL, hi1 == hmax) & seg;
            if (__any_sync(FULL, hmax != 0u && __popc(lead0) + __popc(lead1) > 1)) {      // rare: equal high words somewhere in the warp
                const unsigned lo0 = (hi0 == hmax) ? (unsigned)__double2loint(av0) : 0u;
                const unsigned lo1 = (hi1 == hmax) ? (unsigned)__double2loint(av1) : 0u;
                unsigned lmax = max(lo0, lo1);
#pragma unroll
                for (int o = TPS / 2; o; o >>= 1) lmax = max(lmax, __shfl_xor_sync(FULL, lmax, o));
                lead0 = __ballot_sync(FULL, hi0 == hmax && lo0 == lmax) & seg;
                lead1 = __ballot_sync(FULL, hi1 == hmax && lo1 == lmax) & seg;
            }
            // rows t come before rows t + TPS
            const int src = __ffs(lead0 ? lead0 : lead1) - 1;
            PIB_DBG_ASSERT(src >= (lane & ~(TPS - 1)) && src < (lane & ~(TPS - 1)) + TPS);          // a lane of my own system
            double *const C = cand + (j & 1) * CS;
            if (lane == src) {
                if (lead0) {
#pragma unroll
                    for (int c = 0; c < K; ++c)
                        if (c > j) C[c] = a[0][c];
                    C[K] = r0[0]; C[K + 1] = r1[0];
                    C[K + 2] = hmax ? av0 : -1.0;
                    C[K + 3] = __drcp_rn(a[0][j]);                              // dgetf2 scales by the reciprocal of the pivot
                    reinterpret_cast<int *>(C + K + 4)[0] = t;
                } else {
#pragma unroll
                    for (int c = 0; c < K; ++c)
                        if (c > j) C[c] = a[1][c];
                    C[K] = r0[1]; C[K + 1] = r1[1];
                    C[K + 2] = hmax ? av1 : -1.0;
                    C[K + 3] = __drcp_rn(a[1][j]);
                    reinterpret_cast<int *>(C + K + 4)[0] = t + TPS;
                }
            }
            __syncwarp();
            const double pv = C[K + 2];
            if (go && (!(pv > 0.0) || isinf(pv))) singular = true;              // zero / NaN pivot column (uniform over the system)
            if (go && !singular) {
                const double inv = C[K + 3];
                const int prow = reinterpret_cast<const int *>(C + K + 4)[0];
                PIB_DBG_ASSERT(prow >= 0 && prow < K);
                const bool is_p0 = (prow == t), is_p1 = (prow == t + TPS);
                if (is_p0) { active[0] = false; my_step[0] = j; my_inv[0] = inv; }
                if (is_p1) { active[1] = false; my_step[1] = j; my_inv[1] = inv; }
                // Gauss-Jordan: EVERY other row eliminates column j, the rows already used included
                const double l0 = is_p0 ? 0.0 : a[0][j] * inv, l1 = is_p1 ? 0.0 : a[1][j] * inv;
#pragma unroll
                for (int c = 0; c < K; ++c)
                    if (c > j) {
                        const double pc = C[c];
                        a[0][c] = __fma_rn(-l0, pc, a[0][c]);
                        a[1][c] = __fma_rn(-l1, pc, a[1][c]);
                    }
                const double p0 = C[K];
                r0[0] = __fma_rn(-l0, p0, r0[0]);
                r0[1] = __fma_rn(-l1, p0, r0[1]);
                if (WANT_OK) {
                    const double p1 = C[K + 1];
                    r1[0] = __fma_rn(-l0, p1, r1[0]);
                    r1[1] = __fma_rn(-l1, p1, r1[1]);
                }
            }
        }
        __syncwarp();

        // ---- solution: a row that was the pivot of column s gives u[s] = r0 / pivot, v[s] = r1 / pivot ----
        double s[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};       // 1^T u, 1^T v, u.z, v.z, u.g0, v.g0
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const bool real = (my_step[rr] >= 0 && my_step[rr] < kk);
            const double u = real ? r0[rr] * my_inv[rr] : 0.0, v = real ? r1[rr] * my_inv[rr] : 0.0;
            PIB_DBG_ASSERT(my_step[rr] >= -1 && my_step[rr] < K);
            const double zn = real ? S.zv[my_step[rr]] : 0.0, gn = real ? S.g0[my_step[rr]] : 0.0;
            s[0] += u; s[1] += v; s[2] += u * zn; s[3] += v * zn; s[4] += u * gn; s[5] += v * gn;
        }
#pragma unroll
        for (int o = TPS / 2; o; o >>= 1)
#pragma unroll
            for (int i = 0; i < 6; ++i) s[i] += __shfl_xor_sync(FULL, s[i], o);
        if (t == 0 && store) {
            if (WANT_OK) {
                // mu = (1^T u - 1) / (1^T v), w = u - mu v:  zhat = w.z (ordinary.py:121), sigma^2 = w.g0 + mu (:123)
                double zhat = nan(""), sigma = nan("");
                uint8_t stat = PIB_STATUS_SINGULAR;
                if (!dup) {
                    const double mu = (s[0] - 1.0) / s[1];
                    if (singular || !(fabs(s[1]) > 0.0) || !isfinite(mu)) {                             // decide on the bordered matrix
                        stat = PIB_STATUS_RETRY;
                        if (p.retry_flag) *p.retry_flag = 1;
                    }
                    else {
                        zhat = s[2] - mu * s[3];
                        sigma = (s[4] - mu * s[5]) + mu;
                        stat = PIB_STATUS_OK;
                        if (sigma < 0.0) { sigma = nan(""); stat = PIB_STATUS_NEGATIVE_VARIANCE; }       // ordinary.py:125-126
                    }
                }
                double *o = p.out + ((size_t)set * p.m + q) * 4;
                o[0] = zhat; o[1] = sigma; o[2] = ux; o[3] = uy;
                p.status[(size_t)set * p.m + q] = stat;
                if (stat == PIB_STATUS_SINGULAR && p.singular_flag) *p.singular_flag = 1;
            }
            if (WANT_SK) {
                // w = u:  zhat = (z - mean).w + mean (simple.py:114-116), sigma^2 = w.g0 (:118)
                double zhat = nan(""), sigma = nan("");
                uint8_t stat = PIB_STATUS_SINGULAR;
                if (!dup && !singular) {
                    zhat = (s[2] - p.process_mean * s[0]) + p.process_mean;
                    sigma = s[4];
                    stat = PIB_STATUS_OK;
                    if (sigma < 0.0) { sigma = nan(""); stat = PIB_STATUS_NEGATIVE_VARIANCE; }
                }
                double *o = p.out + (sk_off + (size_t)set * p.m + q) * 4;
                o[0] = zhat; o[1] = sigma; o[2] = ux; o[3] = uy;
                p.status[sk_off + (size_t)set * p.m + q] = stat;
                if (stat == PIB_STATUS_SINGULAR && p.singular_flag) *p.singular_flag = 1;
            }
        }
        __syncwarp();
    }
}

template <int K, int MODE>
static int launch_solve_gj2_m(const SolveParams &sp, cudaStream_t stream)
{
    constexpr int NT = gj2_threads<K>();
    constexpr int SPC = NT / (K / 2);
    const int grid = (int)((sp.count + SPC - 1) / SPC);
    const size_t bytes = sizeof(GjSmem<K>) * SPC;
    PIB_CUDA(cudaFuncSetAttribute(solve_gj2_kernel<K, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    solve_gj2_kernel<K, MODE><<<grid, NT, bytes, stream>>>(sp);
    PIB_CUDA(cudaGetLastError());
    return PIB_OK;
}

template <int K>
static int launch_solve_gj2(const SolveParams &sp, int mode, cudaStream_t stream)
{
    if (mode == PIB_KRIGE_ORDINARY) return launch_solve_gj2_m<K, PIB_KRIGE_ORDINARY>(sp, stream);
    if (mode == PIB_KRIGE_SIMPLE) return launch_solve_gj2_m<K, PIB_KRIGE_SIMPLE>(sp, stream);
    return launch_solve_gj2_m<K, PIB_KRIGE_BOTH>(sp, stream);
}

template <int K>
static int launch_solve_gj(const SolveParams &sp, int mode, cudaStream_t stream)
{
    if (mode == PIB_KRIGE_ORDINARY) return launch_solve_gj_m<K, PIB_KRIGE_ORDINARY>(sp, stream);
    if (mode == PIB_KRIGE_SIMPLE) return launch_solve_gj_m<K, PIB_KRIGE_SIMPLE>(sp, stream);
    return launch_solve_gj_m<K, PIB_KRIGE_BOTH>(sp, stream);
}

// ---------------------------------------------------------------------------------------------
// Singular systems: the reference's fallbacks in solve_weights (kriging/utils/point_kriging_solve.py:134-147)
//
//   try: np.linalg.solve(A, b)
//   except LinAlgError:
//       if mean(A) == 0 or mean(b) == 0: ZerosMatrixWarning, weights = 0
//       elif allow_lsa:                  LeastSquaresApproximationWarning, weights = lstsq(A, b, rcond=None)[0]
//       else:                            raise
//
// The solve kernels mark such systems PIB_STATUS_SINGULAR (exact zero pivot, the condition dgesv reports).
// This kernel runs after them: warps scan the status bytes, and every singular system is rebuilt and
// finished here.  lstsq(rcond=None) is the minimum-norm least-squares solution with singular values below
// eps * n * s_max dropped; A is symmetric, so its SVD is its eigen-decomposition A = V L V^T (s = |l|) and
// x = sum_{|l_i| > cut} v_i (v_i . b) / l_i.  The eigen-decomposition is a cyclic Jacobi iteration (warp per
// system, A in shared memory, V in shared memory when both fit, else in a global scratch slab).
// ---------------------------------------------------------------------------------------------
struct SingularParams {
    SolveParams s;
    int allow_lsa;
    double *vglobal;       // NULL (V lives in shared memory) or [gridDim.x][dmax * stride]
};

static __device__ int jacobi_eigen(double *A, double *V, int n, int st, int lane)
{
    for (int idx = lane; idx < n * n; idx += 32) {
        const int i = idx / n, j = idx - i * n;
        V[i * st + j] = (i == j) ? 1.0 : 0.0;
    }
    __syncwarp();
    int sweep = 0;
    for (; sweep < 60; ++sweep) {
        double off = 0.0, tot = 0.0;
        for (int idx = lane; idx < n * n; idx += 32) {
            const int i = idx / n, j = idx - i * n;
            const double a = A[i * st + j];
            tot += a * a;
            if (i != j) off += a * a;
        }
        for (int o = 16; o; o >>= 1) {
            off += __shfl_xor_sync(0xffffffffu, off, o);
            tot += __shfl_xor_sync(0xffffffffu, tot, o);
        }
        if (!(off > tot * 1e-36)) break;                         // off-diagonal mass below 1e-18 ||A||_F (or NaN)
        const double thr = sqrt(tot) * 1e-20;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = A[p * st + q];
                if (!(fabs(apq) > thr)) continue;                // warp-uniform
                const double app = A[p * st + p], aqq = A[q * st + q];
                const double theta = (aqq - app) / (2.0 * apq);
                double t;
                if (fabs(theta) > 1e150) t = 0.5 / theta;
                else t = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                __syncwarp();
                for (int i = lane; i < n; i += 32) {
                    if (i != p && i != q) {
                        const double aip = A[i * st + p], aiq = A[i * st + q];
                        const double np_ = c * aip - s * aiq, nq_ = s * aip + c * aiq;
                        A[i * st + p] = np_; A[p * st + i] = np_;
                        A[i * st + q] = nq_; A[q * st + i] = nq_;
                    }
                    const double vip = V[i * st + p], viq = V[i * st + q];
                    V[i * st + p] = c * vip - s * viq;
                    V[i * st + q] = s * vip + c * viq;
                }
                if (lane == 0) {
                    A[p * st + p] = app - t * apq; A[q * st + q] = aqq + t * apq;
                    A[p * st + q] = 0.0; A[q * st + p] = 0.0;
                }
                __syncwarp();
            }
    }
    return sweep;
}

__global__ void __launch_bounds__(32) singular_kernel(const SingularParams sp)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const SolveParams &p = sp.s;
    const int lane = threadIdx.x;
    const int kmax = p.kk, mode = p.mode;
    const int dmax = (mode == PIB_KRIGE_ORDINARY) ? kmax + 1 : kmax;
    const int st = p.stride;
    double *A = reinterpret_cast<double *>(smem_raw);
    double *V = sp.vglobal ? sp.vglobal + (size_t)blockIdx.x * dmax * st : A + (size_t)dmax * st;
    double *rhs = (sp.vglobal ? A : V) + (size_t)dmax * st;
    double *xv = rhs + dmax;
    double *nx = xv + dmax, *ny = nx + kmax, *nd = ny + kmax;
    int32_t *npos = reinterpret_cast<int32_t *>(nd + kmax);
    const GridIndex &g = p.g;

    if (p.singular_flag && *p.singular_flag == 0) return;                       // no solver marked a system singular
    for (int64_t base = (int64_t)blockIdx.x * 32; base < p.count; base += (int64_t)gridDim.x * 32) {
        const int64_t mine = base + lane;
        bool hit = false;
        int32_t myq = 0;
        if (mine < p.count) {
            myq = p.order[mine];
            for (int set = 0; set < p.n_sets; ++set) hit |= p.status[(size_t)set * p.m + myq] == PIB_STATUS_SINGULAR;
        }
        unsigned todo = __ballot_sync(0xffffffffu, hit);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int64_t slot = base + src;
            const int32_t q = __shfl_sync(0xffffffffu, myq, src);
            const int kk = p.nbr_count ? p.nbr_count[slot] : kmax;
            if (kk <= 0) continue;
            const int dim = (mode == PIB_KRIGE_ORDINARY) ? kk + 1 : kk;
            __syncwarp();
            for (int t = lane; t < kk; t += 32) {
                const int pos = p.nbr_pos[slot * kmax + t];
                PIB_DBG_ASSERT(pos >= 0 && pos < g.n);
                npos[t] = pos; nx[t] = g.x[pos]; ny[t] = g.y[pos]; nd[t] = p.nbr_d[slot * kmax + t];
            }
            __syncwarp();
            for (int set = 0; set < p.n_sets; ++set) {
                if (p.status[(size_t)set * p.m + q] != PIB_STATUS_SINGULAR) continue;
                const double nugget = p.params[3 * set], sill = p.params[3 * set + 1], rang = p.params[3 * set + 2];
                for (int i = 0; i < kk; ++i)
                    for (int j = i + 1 + lane; j < kk; j += 32) {
                        const double dx = __dsub_rn(nx[i], nx[j]), dy = __dsub_rn(ny[i], ny[j]);
                        const double h = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
                        const double gm = gamma_model(p.model, h, nugget, sill, rang);
                        A[i * st + j] = gm; A[j * st + i] = gm;
                    }
                for (int i = lane; i < kk; i += 32) {
                    A[i * st + i] = nugget;
                    rhs[i] = gamma_model(p.model, nd[i], nugget, sill, rang);
                    if (mode == PIB_KRIGE_ORDINARY) { A[i * st + kk] = 1.0; A[kk * st + i] = 1.0; }
                }
                if (mode == PIB_KRIGE_ORDINARY && lane == 0) { A[kk * st + kk] = 0.0; rhs[kk] = 1.0; }
                __syncwarp();
                // np.mean(A) == 0 or np.mean(b) == 0  (entries are >= 0 for valid models, so sum == 0 <=> all zero)
                double sa = 0.0, sb = 0.0;
                for (int idx = lane; idx < dim * dim; idx += 32) { const int i = idx / dim, j = idx - i * dim; sa += A[i * st + j]; }
                for (int i = lane; i < dim; i += 32) sb += rhs[i];
                for (int o = 16; o; o >>= 1) {
                    sa += __shfl_xor_sync(0xffffffffu, sa, o);
                    sb += __shfl_xor_sync(0xffffffffu, sb, o);
                }
                uint8_t stat;
                if (sa == 0.0 || sb == 0.0) {
                    for (int i = lane; i < dim; i += 32) xv[i] = 0.0;
                    stat = PIB_STATUS_ZERO_MATRIX;
                } else if (!sp.allow_lsa) {
                    continue;                                     // stays PIB_STATUS_SINGULAR: the caller raises
                } else {
                    jacobi_eigen(A, V, dim, st, lane);
                    double lmax = 0.0;
                    for (int i = lane; i < dim; i += 32) lmax = fmax(lmax, fabs(A[i * st + i]));
                    for (int o = 16; o; o >>= 1) lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
                    const double cut = 2.220446049250313e-16 * dim * lmax;   // lstsq(rcond=None): eps * max(M, N) * s_max
                    for (int i = lane; i < dim; i += 32) xv[i] = 0.0;
                    __syncwarp();
                    for (int e = 0; e < dim; ++e) {
                        const double lam = A[e * st + e];
                        if (!(fabs(lam) > cut)) continue;
                        double dot = 0.0;
                        for (int r = lane; r < dim; r += 32) dot = __fma_rn(V[r * st + e], rhs[r], dot);
                        for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
                        const double coef = dot / lam;
                        for (int r = lane; r < dim; r += 32) xv[r] = __fma_rn(coef, V[r * st + e], xv[r]);
                    }
                    stat = PIB_STATUS_LSA;
                }
                __syncwarp();
                double zs = 0.0, ss = 0.0;
                for (int i = lane; i < dim; i += 32) {
                    const double w = xv[i];
                    ss = __fma_rn(w, rhs[i], ss);
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
                if (lane == 0) {
                    double *o = p.out + ((size_t)set * p.m + q) * 4;
                    o[0] = (mode == PIB_KRIGE_SIMPLE) ? zs + p.process_mean : zs;
                    o[1] = ss < 0.0 ? nan("") : ss;               // ordinary.py:125-126
                    p.status[(size_t)set * p.m + q] = stat;
                }
                __syncwarp();
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
int krige_device(const double *d_known, int64_t n, const double *d_values, int n_sets, const double *params,
                 const double *d_unknown, int64_t m, int model, int mode, double process_mean, int k,
                 double neighbors_range, double direction, double max_tick, int use_all, int allow_lsa,
                 const int32_t *d_exclude, double *d_out, int32_t *d_nbr_idx, uint8_t *d_status, cudaStream_t stream,
                 const GridIndex *prebuilt)
{
    const bool directional = !std::isnan(direction);
    const bool ranged = directional || use_all;                 // the selection depends on neighbors_range
    PIB_CHECK(!ranged || (neighbors_range > 0 && std::isfinite(neighbors_range)), PIB_ERR_INVALID,
              "krige: directional / use_all selection needs a positive neighbors_range");
    PIB_CHECK(n >= 1 && m >= 0 && k >= 1, PIB_ERR_INVALID, "krige: need n >= 1, m >= 0, k >= 1");
    PIB_CHECK(n_sets >= 1 && params != nullptr, PIB_ERR_INVALID, "krige: need at least one parameter set");
    PIB_CHECK(model >= 0 && model <= 6, PIB_ERR_INVALID, "krige: unknown model id");
    PIB_CHECK(mode == PIB_KRIGE_ORDINARY || mode == PIB_KRIGE_SIMPLE || mode == PIB_KRIGE_BOTH, PIB_ERR_INVALID,
              "krige: unknown mode");
    PIB_CHECK(m < ((int64_t)1 << 31) && n < ((int64_t)1 << 31), PIB_ERR_UNSUPPORTED, "krige: more than 2^31-1 rows");
    PIB_CHECK(std::min<int64_t>(k, n) <= MAX_K, PIB_ERR_UNSUPPORTED, "krige: no_neighbors above 128 is not supported");
    const int64_t n_avail = d_exclude ? n - 1 : n;              // leave-one-out: every location loses one row
    PIB_CHECK(n_avail >= 1, PIB_ERR_INVALID, "krige: leave-one-out needs at least two known points");
    const int kk = use_all ? (int)std::min<int64_t>(MAX_K, n_avail) : (int)std::min<int64_t>(k, n_avail);
    int last_tol = 1;
    if (directional) {
        last_tol = std::max(2, (int)std::floor(max_tick) + 1);   // 1, 2, ... until the tolerance exceeds max_tick
        PIB_CHECK(last_tol <= MAX_TOL, PIB_ERR_UNSUPPORTED, "krige: max_tick above 63 degrees is not supported");
    }
    if (!prebuilt) {                   // (a chunked host call accumulates the kernel times of all its chunks)
        timing_reset(PIB_T_KNN);
        timing_reset(PIB_T_SOLVE);
    }
    if (m == 0) return PIB_OK;

    Scratch ws(stream);
    GridIndex g;
    if (prebuilt) g = *prebuilt;
    else PIB_TRY(build_grid_index(d_known, n, 4.0, ws, stream, &g));
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

    const int dim = (mode != PIB_KRIGE_SIMPLE) ? kk + 1 : kk;           // the old solvers' (bordered) system; BOTH: the larger one
    const int stride = (dim + 1) | 1;
    const size_t per_warp = solve_smem_per_warp(kk, dim, stride);
    const int n_modes = mode == PIB_KRIGE_BOTH ? 2 : 1;
    const size_t mode_stride = (size_t)n_sets * m;                        // BOTH: the simple-kriging block follows the ordinary one
    int solve_warps = (int)std::min<size_t>(8, (size_t)(MAX_SMEM_K - 1024) / per_warp);
    PIB_CHECK(solve_warps >= 1, PIB_ERR_UNSUPPORTED, "krige: system does not fit in shared memory");
    PIB_CUDA(cudaFuncSetAttribute(solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per_warp * solve_warps)));
    const size_t knn_smem = (size_t)KNN_WARPS * 2 * k2 * sizeof(Cand);
    PIB_CUDA(cudaFuncSetAttribute(knn_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)knn_smem));
    PIB_CUDA(cudaFuncSetAttribute(knn_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)knn_smem));
    PIB_CUDA(cudaFuncSetAttribute(knn_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)knn_smem));
    PIB_CUDA(cudaFuncSetAttribute(knn_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)knn_smem));
    const char *knn_env = std::getenv("PIB_KNN_FAST");
    const bool knn_fast = !(knn_env && knn_env[0] == '0');
    PIB_CUDA(cudaFuncSetAttribute(knn_dir_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)knn_smem));

    const char *force_smem = std::getenv("PIB_FORCE_SMEM_SOLVE");
    const bool use_reg = kk <= 64 && !(force_smem && force_smem[0] == '1');
    const char *solve_env = std::getenv("PIB_SOLVE_V");              // "1": the round-1 bordered LU instead of Gauss-Jordan
    const bool use_gj = use_reg && !(solve_env && solve_env[0] == '1');
    const int64_t chunk = std::min<int64_t>(m, (int64_t)1 << 21);
    int32_t *nbr_pos;
    double *nbr_d;
    PIB_TRY(ws.alloc(&nbr_pos, (size_t)chunk * kk));
    PIB_TRY(ws.alloc(&nbr_d, (size_t)chunk * kk));
    int32_t *nbr_count = nullptr;
    if (ranged) PIB_TRY(ws.alloc(&nbr_count, (size_t)chunk));
    const int ksel = (int)std::min<int64_t>(k, n_avail);        // what the plain k-nearest search selects
    if (!d_status) {                                            // the singular-system pass needs the status bytes
        PIB_TRY(ws.alloc(&d_status, (size_t)n_modes * n_sets * m));
    }
    // singular-system pass: one warp per CTA; V next to A in shared memory when both fit, else in global scratch
    const size_t mat_bytes = (size_t)dim * stride * 8;
    const size_t sing_tail = (size_t)dim * 16 + (size_t)kk * 24 + (size_t)((kk + 1) / 2 * 2) * 4 + 16;
    const bool v_in_smem = 2 * mat_bytes + sing_tail <= 100 * 1024;
    const size_t sing_smem = (v_in_smem ? 2 : 1) * mat_bytes + sing_tail;
    const int sing_grid = sm_count() * (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)(MAX_SMEM_K - 1024) / (sing_smem + 1024)));
    double *vglobal = nullptr;
    if (!v_in_smem) PIB_TRY(ws.alloc(&vglobal, (size_t)sing_grid * dim * stride));
    PIB_CUDA(cudaFuncSetAttribute(singular_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sing_smem));

    int *d_retry_flag = nullptr;                    // [0] systems marked PIB_STATUS_RETRY, [1] systems marked PIB_STATUS_SINGULAR
    PIB_TRY(ws.alloc(&d_retry_flag, 2));
    for (int64_t begin = 0; begin < m; begin += chunk) {
        const int64_t count = std::min(chunk, m - begin);
        KnnParams kp;
        kp.g = g; kp.unknown = d_unknown; kp.order = order + begin; kp.count = count; kp.kk = kk; kp.k2 = k2; kp.r0 = r0;
        kp.nbr_pos = nbr_pos; kp.nbr_d = nbr_d; kp.nbr_idx = d_nbr_idx; kp.k = k;
        kp.r_try = g.h * std::sqrt(1.8 * kk / (3.14159265358979 * std::max(per_cell, 1e-9)));
        kp.nbr_count = nbr_count; kp.range = neighbors_range; kp.direction = direction; kp.last_tol = last_tol;
        kp.use_all = use_all; kp.stride = kk; kp.isotropic = directional ? 0 : 1; kp.exclude = d_exclude;
        timing_begin(PIB_T_KNN, stream);
        if (!directional && use_all) {
            // k nearest first (the fallback rows), then the range kernel replaces them where >= k points are in range
            KnnParams kn = kp;
            kn.kk = ksel;
            knn_kernel<0><<<(int)((count + KNN_WARPS - 1) / KNN_WARPS), KNN_WARPS * 32, knn_smem, stream>>>(kn);
            knn_dir_kernel<<<(int)((count + KNN_WARPS - 1) / KNN_WARPS), KNN_WARPS * 32, knn_smem, stream>>>(kp);
        } else if (directional)
            knn_dir_kernel<<<(int)((count + KNN_WARPS - 1) / KNN_WARPS), KNN_WARPS * 32, knn_smem, stream>>>(kp);
        else if (knn_fast && kk <= 16 && r0 == 1)             // a 3 x 3 window (~36 points): 64 register slots are enough
            knn_kernel<2><<<(int)((count + KNN_WARPS - 1) / KNN_WARPS), KNN_WARPS * 32, knn_smem, stream>>>(kp);
        else if (knn_fast && kk <= 40)
            knn_kernel<4><<<(int)((count + KNN_WARPS - 1) / KNN_WARPS), KNN_WARPS * 32, knn_smem, stream>>>(kp);
        else if (knn_fast && kk <= 96)
            knn_kernel<8><<<(int)((count + KNN_WARPS - 1) / KNN_WARPS), KNN_WARPS * 32, knn_smem, stream>>>(kp);
        else
            knn_kernel<0><<<(int)((count + KNN_WARPS - 1) / KNN_WARPS), KNN_WARPS * 32, knn_smem, stream>>>(kp);
        timing_end(PIB_T_KNN, stream);
        PIB_CUDA(cudaGetLastError());
        SolveParams sp;
        sp.g = g; sp.unknown = d_unknown; sp.order = order + begin; sp.count = count; sp.m = m; sp.n = n; sp.kk = kk; sp.k = k;
        sp.nbr_pos = nbr_pos; sp.nbr_d = nbr_d; sp.values = d_values; sp.params = d_params; sp.n_sets = n_sets;
        sp.model = model; sp.mode = mode; sp.process_mean = process_mean; sp.out = d_out; sp.status = d_status;
        sp.warps = solve_warps; sp.stride = stride; sp.nbr_count = nbr_count;
        sp.only_status = 0; sp.mode_stride = mode_stride;
        sp.retry_flag = use_gj ? d_retry_flag : nullptr;
        sp.singular_flag = d_retry_flag + 1;
        { const char *fd = std::getenv("PIB_FAST_DIV"); sp.fast_div = !(fd && fd[0] == '0'); }
        PIB_CUDA(cudaMemsetAsync(d_retry_flag, 0, 2 * sizeof(int), stream));
        timing_begin(PIB_T_SOLVE, stream);
        int rc = PIB_OK;
        // the round-1 solvers on the bordered matrix, one kriging mode at a time (`only` = just the flagged systems)
        auto bordered = [&](int one_mode, int only) -> int {
            SolveParams so = sp;
            so.mode = one_mode; so.only_status = only;
            if (mode == PIB_KRIGE_BOTH && one_mode == PIB_KRIGE_SIMPLE) { so.out = d_out + mode_stride * 4; so.status = d_status + mode_stride; }
            if (use_reg && kk <= 8) return launch_solve_reg<8>(so, one_mode, stream);
            if (use_reg && kk <= 16) return launch_solve_reg<16>(so, one_mode, stream);
            if (use_reg && kk <= 32) return launch_solve_reg<32>(so, one_mode, stream);
            if (use_reg && kk <= 64) return launch_solve_reg<64>(so, one_mode, stream);
            // shared-memory solver: the layout is sized for the larger (ordinary) system when both modes run
            solve_kernel<<<(int)((count + solve_warps - 1) / solve_warps), solve_warps * 32, per_warp * solve_warps, stream>>>(so);
            if (cudaGetLastError() != cudaSuccess) { set_error("solve_kernel launch failed"); return PIB_ERR_CUDA; }
            return PIB_OK;
        };
        const char *gj2_env = std::getenv("PIB_GJ2");                    // "0": one row per thread for k <= 16 as well; "1": two rows up to k = 32
        const bool use_gj2 = use_gj && (kk <= 16 || (kk <= 32 && gj2_env && gj2_env[0] == '1')) && !(gj2_env && gj2_env[0] == '0');
        set_kernel_name(PIB_T_SOLVE, use_gj2 ? "solve_gj2_kernel" : use_gj ? "solve_gj_kernel" : use_reg ? "solve_reg_kernel" : "solve_kernel");
        if (use_gj2) {
            if (kk <= 8) rc = launch_solve_gj2<8>(sp, mode, stream);
            else if (kk <= 16) rc = launch_solve_gj2<16>(sp, mode, stream);
            else rc = launch_solve_gj2<32>(sp, mode, stream);
        } else if (use_gj) {
            if (kk <= 8) rc = launch_solve_gj<8>(sp, mode, stream);
            else if (kk <= 16) rc = launch_solve_gj<16>(sp, mode, stream);
            else if (kk <= 32) rc = launch_solve_gj<32>(sp, mode, stream);
            else rc = launch_solve_gj<64>(sp, mode, stream);
        }
        if (use_gj) {
            // systems Gauss-Jordan on Gamma could not decide (marked PIB_STATUS_RETRY): bordered LU; returns at once otherwise
            if (rc == PIB_OK && mode != PIB_KRIGE_SIMPLE) rc = bordered(PIB_KRIGE_ORDINARY, PIB_STATUS_RETRY);
            if (rc == PIB_OK && mode != PIB_KRIGE_ORDINARY) rc = bordered(PIB_KRIGE_SIMPLE, PIB_STATUS_RETRY);
        } else {
            if (mode != PIB_KRIGE_SIMPLE) rc = bordered(PIB_KRIGE_ORDINARY, 0);
            if (rc == PIB_OK && mode != PIB_KRIGE_ORDINARY) rc = bordered(PIB_KRIGE_SIMPLE, 0);
        }
        for (int one = 0; one < 2 && rc == PIB_OK; ++one) {
            const int one_mode = one == 0 ? PIB_KRIGE_ORDINARY : PIB_KRIGE_SIMPLE;
            if (mode != PIB_KRIGE_BOTH && mode != one_mode) continue;
            SingularParams sg;
            sg.s = sp; sg.s.mode = one_mode; sg.allow_lsa = allow_lsa; sg.vglobal = vglobal;
            if (mode == PIB_KRIGE_BOTH && one_mode == PIB_KRIGE_SIMPLE) { sg.s.out = d_out + mode_stride * 4; sg.s.status = d_status + mode_stride; }
            singular_kernel<<<sing_grid, 32, sing_smem, stream>>>(sg);
            if (cudaGetLastError() != cudaSuccess) { set_error("singular_kernel launch failed"); rc = PIB_ERR_CUDA; }
        }
        timing_end(PIB_T_SOLVE, stream);
        PIB_TRY(rc);
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

namespace pib {
// the grid index of the known points alone (a chunked host call builds it once and hands it to krige_device per chunk)
int krige_build_index(const double *d_known, int64_t n, Scratch &ws, cudaStream_t stream, GridIndex *out)
{
    PIB_CHECK(n >= 1 && n < ((int64_t)1 << 31), PIB_ERR_INVALID, "krige: need 1 <= n < 2^31 known points");
    return build_grid_index(d_known, n, 4.0, ws, stream, out);
}
}  // namespace pib
