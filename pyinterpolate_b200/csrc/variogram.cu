// Experimental variogram pair kernels (omnidirectional and ellipse-directional), sm_100a.
//
// What the reference does (src/pyinterpolate/): an N x N cdist matrix (distance/point.py:56),
// then one boolean-mask pass + np.where + gather PER LAG (distance/point.py:86-92,
// semivariogram/experimental/functions/general.py:235-303); the ellipse path loops over points
// and lags in Python (functions/directional.py:395-465, distance/angular.py:641-681).
//
// What this file does instead: ONE pass over the upper triangle of point pairs.
//   * points are Morton-sorted (spatial.cu) so 128 consecutive points form a compact tile;
//     a tile pair whose bounding boxes are farther apart than the last lag edge is skipped
//   * a CTA owns one A tile (thread t keeps point t of it in registers) and streams B tiles
//     through shared memory with TMA bulk copies (cp.async.bulk + mbarrier, double buffered)
//   * distances are computed in the reference's operation order (cdist: dx*dx + dy*dy with each
//     product and the sum rounded separately; ellipse: the dgemm FMA pattern followed by an
//     un-fused n0*n0 + n1*n1), so the value that is compared against a lag edge is bit-identical
//     to the reference's squared distance
//   * lag assignment is sqrt-free and exact: the host turns every edge e into the largest double
//     T with sqrt(T) <= e (IEEE sqrt is monotone and correctly rounded), so  d <= e  <=>  d2 <= T
//   * per-lag accumulation is the bottleneck (data-dependent bin per pair).  Shared-memory fp64
//     atomics are CAS loops on this architecture, so every THREAD owns a private histogram in
//     shared memory ([slot][thread] layout => bank-conflict free, no atomics).  Per pair it adds
//     dz, dz^2 and 1 to one slot; because a thread's own point z_i is fixed for a whole work item,
//     sum z_i z_j and sum (z_i + z_j) are recovered at the end of the item from (count, sum dz)
//   * private histograms are folded warp-wise into per-warp global partials after every work item
//     and a last kernel adds the partials in a fixed order: results are deterministic
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace pib {

constexpr int TILE = 128;        // points per tile
constexpr int CHUNK = 64;        // B tiles per work item
constexpr int LUT2_MAX = 2048;   // slot table: (high word of the fp64 squared distance - base) >> LUT2_SHIFT -> first candidate slot
constexpr int LUT2_SHIFT = 13;   // keep 7 of the 20 mantissa bits of the high word: 128 cells per octave
constexpr int MAX_SMEM = 232448; // opt-in dynamic shared memory per CTA on sm_100

struct PairParams {
    const double2 *xy;           // [n_pad]
    const double *z;             // [n_pad]
    const double4 *tile_box;     // [n_tiles]
    int64_t n;                   // real points
    int n_tiles;
    int n_slots;                 // n_lags + 2: slot 0 = "d2 <= 0 / masked", slot s = lag s-1, last = beyond range
    const double *upper;         // [n_slots] slot upper thresholds on d2: U[0] = 0, U[s] = T(edge[s-1]), U[last] = +inf
    const uint16_t *lut2;        // [lut2_n] (hi word of d2 - lut2_base) >> 13 -> first candidate slot
    int lut2_base, lut2_n;
    int lut2_base_f;             // the same base as fp32 bits (a normal float's bits are (high word - 896 * 2^20) << 3 | 3 more mantissa bits)
    double box_slack;            // ellipse ring kernel: slack on the whitened bounding boxes
    double max_d2;               // T(edges[n_lags-1])
    double w00, w01, w10, w11;   // whitening matrix (ellipse)
    int shard_index, shard_count;
    double *partial;             // [grid][warps][n_slots][4]  (sum_sq, sum_prod, sum_val, count as double-exact int64 bits)
    unsigned long long *pairs_evaluated;
    int group_path;              // ring kernel: one slot lookup per group of 8 pairs (off-diagonal tile pairs)
    const double2 *gsum;         // [n_pad / 8] (sum z, sum z^2) over each group of 8 consecutive (sorted) points; z centred
    const int32_t *item_list;    // ring kernel: work items ordered by descending cost (NULL: natural order)
    const unsigned *n_live_items;  // number of items with at least one live tile pair (device scalar)
    int chunk;                     // pair_ring6_kernel / item_cost_kernel: B tiles per work item (<= CHUNK)
    int hiw_ok;                    // pair_ring6_kernel: no threshold has an all-ones low word (high-word fast path is exact)
    const int *run_flag;           // NULL, or device flag: the kernel returns at once when it is 0 (conditional re-run)
    // pair_ring6_kernel, fp32 screening path (see sweep_f32): per tile TILE / 2 x {x0, x1, y0, y1} single-precision coordinates
    // relative to the tile's origin, then the origin itself (double2) — TF32_TILE_BYTES per tile; NULL = path off
    const unsigned char *tf32;
    const float2 *slotf;           // [n_slots + 4] (fp32 threshold, half-width of its undecided band) per slot
    double f32_rcap;               // tiles wider than 2 * f32_rcap (either axis) take the exact path
};
constexpr int TF32_TILE_BYTES = TILE * 8 + 16;

// ---- PTX helpers: mbarrier + 1D TMA bulk copy -------------------------------------------------
static __device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

static __device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
static __device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
static __device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
static __device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n"     // suspend-time hint: fewer polls from waiting warps
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
static __device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
static __device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// exact-safe lower bound of the squared distance between two boxes, same operation order as the pairs
static __device__ __forceinline__ double box_min_d2(const double4 a, const double4 b)
{
    const double dx = fmax(0.0, fmax(a.x - b.z, b.x - a.z));
    const double dy = fmax(0.0, fmax(a.y - b.w, b.y - a.w));
    return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
}

// Shared memory:
//   static : B tile double buffer (TMA destination), slot thresholds, distance->slot table, mbarriers
//   dynamic: double2 acc[n_slots][NT] private (sum dz, sum dz^2); uint32_t cnt[n_slots][NT] private counts
// The read-only tables are separate static objects so the compiler can prove that loads from them never
// alias the private histogram stores and may hoist them freely (that is what lets it overlap the
// distance / slot computation of one batch of pairs with the read-modify-writes of the previous batch).
constexpr int MAX_SLOTS = 352;
constexpr int STATIC_SMEM = 2 * TILE * 24 + MAX_SLOTS * 8 + LUT2_MAX * 2 + 64;

template <int NT>
static size_t pair_smem_bytes(int n_slots)
{
    return (size_t)n_slots * NT * (sizeof(double2) + sizeof(uint32_t));
}

// FIX = number of branch-free threshold steps after the table lookup (0: data-dependent loop)
template <int NT, bool ELLIPSE, int FIX>
__global__ void __launch_bounds__(NT, 1) pair_kernel(const PairParams p)
{
    static_assert(TILE % NT == 0, "an A sub-tile is NT consecutive points of one tile");
    if (p.run_flag && *p.run_flag == 0) return;    // conditional re-run (cancellation guard): nothing to do
    constexpr int A_PER_TILE = TILE / NT;          // thread t owns point t of the A sub-tile
    constexpr int U = 8;                           // pairs per batch
    static_assert(TILE % (2 * U) == 0, "tile must hold an even number of batches");
    __shared__ __align__(16) double2 s_txy[2][TILE];
    __shared__ __align__(16) double s_tz[2][TILE];
    __shared__ long long s_upper[MAX_SLOTS];       // thresholds as bit patterns: d2 >= 0, so integer order == fp order
    __shared__ uint16_t s_lut[LUT2_MAX];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ uint64_t s_live;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_slots = p.n_slots;

    double2 *acc = reinterpret_cast<double2 *>(smem_raw) + tid;            // acc[s * NT]
    uint32_t *cnt = reinterpret_cast<uint32_t *>(smem_raw + (size_t)n_slots * NT * sizeof(double2)) + tid;

    for (int s = 0; s < n_slots; ++s) {
        acc[s * NT] = make_double2(0.0, 0.0);
        cnt[s * NT] = 0u;
    }
    for (int s = tid; s < n_slots; s += NT) s_upper[s] = __double_as_longlong(p.upper[s]);
    for (int s = tid; s < p.lut2_n; s += NT) s_lut[s] = p.lut2[s];
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int n_tiles = p.n_tiles;
    const int n_chunks = (n_tiles + CHUNK - 1) / CHUNK;
    const int64_t n_items = (int64_t)n_tiles * A_PER_TILE * n_chunks;
    uint32_t seq = 0;                       // tiles staged so far by this CTA (buffer = seq & 1, phase = (seq >> 1) & 1)
    unsigned long long evaluated = 0;
    double *my_partial = p.partial + ((size_t)blockIdx.x * (NT / 32) + warp) * n_slots * 4;
    const int base_hi = p.lut2_base, lut_last = p.lut2_n - 1;
    const double w00 = p.w00, w01 = p.w01, w10 = p.w10, w11 = p.w11;

    for (int64_t item = ((int64_t)blockIdx.x) * p.shard_count + p.shard_index; item < n_items;
         item += (int64_t)gridDim.x * p.shard_count) {
        const int64_t a0 = (item / n_chunks) * NT;            // first point of my A sub-tile
        const int ta = (int)(a0 / TILE);                      // the tile that contains it
        const int tb0 = (int)(item % n_chunks) * CHUNK;
        const int tb1 = min(tb0 + CHUNK, n_tiles);
        const int valid_a = (int)max((int64_t)0, min((int64_t)NT, p.n - a0));
        if (tb1 <= ta || valid_a == 0) continue;              // chunk entirely below the diagonal / padding only
        const int tb_first = max(tb0, ta);

        // ---- which B tiles of the chunk can hold a pair within range? (uniform decision) ----
        const double4 abox = p.tile_box[ta];
        if (warp == 0) {
            uint64_t m = 0;
            for (int h = 0; h < CHUNK / 32; ++h) {
                const int tb = tb0 + h * 32 + lane;
                bool live = false;
                if (tb >= tb_first && tb < tb1) {
                    const double4 bbox = p.tile_box[tb];
                    if (ELLIPSE) {                      // whitened boxes + slack (see pair_ring_kernel)
                        const double gx = fmax(0.0, fmax(abox.x - bbox.z, bbox.x - abox.z) - p.box_slack);
                        const double gy = fmax(0.0, fmax(abox.y - bbox.w, bbox.y - abox.w) - p.box_slack);
                        live = (gx * gx + gy * gy) * (1.0 - 1e-12) <= p.max_d2;
                    } else live = box_min_d2(abox, bbox) <= p.max_d2;
                }
                m |= (uint64_t)__ballot_sync(0xffffffffu, live) << (32 * h);
            }
            if (lane == 0) s_live = m;
        }
        __syncthreads();
        uint64_t live = s_live;
        if (live == 0) { __syncthreads(); continue; }

        // ---- my A point ----------------------------------------------------------------------
        const int ig = (int)a0 + tid;
        const double2 pa = p.xy[ig];
        const double xi = pa.x, yi = pa.y, zi = p.z[ig];

        // prefetch the first live tile
        int tb = tb0 + __ffsll((long long)live) - 1;
        live &= live - 1;
        if (tid == 0) {
            const uint32_t b = seq & 1;
            mbar_expect_tx(&s_bar[b], TILE * (sizeof(double2) + sizeof(double)));
            tma_load_1d(s_txy[b], p.xy + (size_t)tb * TILE, TILE * sizeof(double2), &s_bar[b]);
            tma_load_1d(s_tz[b], p.z + (size_t)tb * TILE, TILE * sizeof(double), &s_bar[b]);
        }
        while (true) {
            const uint32_t b = seq & 1, phase = (seq >> 1) & 1;
            const int tb_next = live ? tb0 + __ffsll((long long)live) - 1 : -1;
            live &= live - 1;
            if (tid == 0 && tb_next >= 0) {                   // buffer b^1 was released by the barrier below
                const uint32_t nb = b ^ 1;
                mbar_expect_tx(&s_bar[nb], TILE * (sizeof(double2) + sizeof(double)));
                tma_load_1d(s_txy[nb], p.xy + (size_t)tb_next * TILE, TILE * sizeof(double2), &s_bar[nb]);
                tma_load_1d(s_tz[nb], p.z + (size_t)tb_next * TILE, TILE * sizeof(double), &s_bar[nb]);
            }
            mbar_wait(&s_bar[b], phase);
            ++seq;

            const double2 *bxy = s_txy[b];
            const double *bz = s_tz[b];
            const bool diagonal = (tb == ta);
            const int jbase = tb * TILE;
            if (tid == 0) {
                const long long vb = min((int64_t)TILE, p.n - (int64_t)tb * TILE), va = valid_a;
                // diagonal: pairs with j > i, i in [a0, a0+va), j in [jbase, jbase+vb)
                evaluated += diagonal ? (unsigned long long)(va * (jbase + vb - 1) - va * a0 - va * (va - 1) / 2)
                                      : (unsigned long long)(va * vb);
            }

            // slot and dz of U consecutive pairs, written stage by stage (all U distances, then all U
            // table lookups, ...) so the U independent dependency chains are interleaved in the schedule:
            // with one warp per scheduler the latency of each step is hidden by the other U-1 pairs
            auto compute = [&](int j0, int (&slot)[U], double (&dzv)[U]) {
                double d2[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const double2 pj = bxy[j0 + u];
                    dzv[u] = __dsub_rn(zi, bz[j0 + u]);
                    if (ELLIPSE) {
                        // vector_distance = other - centre (angular.py:448); W @ d as OpenBLAS dgemm
                        // evaluates it: fma(W[r][1], dy, W[r][0]*dx); einsum: n0*n0 + n1*n1 un-fused
                        const double dx = __dsub_rn(pj.x, xi), dy = __dsub_rn(pj.y, yi);
                        const double n0 = __fma_rn(w01, dy, __dmul_rn(w00, dx));
                        const double n1 = __fma_rn(w11, dy, __dmul_rn(w10, dx));
                        d2[u] = __dadd_rn(__dmul_rn(n0, n0), __dmul_rn(n1, n1));
                    } else {
                        // scipy cdist euclidean: sqrt(dx*dx + dy*dy), nothing fused (distance/point.py:56)
                        const double dx = __dsub_rn(xi, pj.x), dy = __dsub_rn(yi, pj.y);
                        d2[u] = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                    }
                }
                // first candidate slot from the exponent and top mantissa bits of the squared distance (integer ops
                // only: no conversion, no sqrt on the quarter-rate XU pipe) ...
                int sl[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    int idx = (__double2hiint(d2[u]) - base_hi) >> LUT2_SHIFT;
                    idx = min(max(idx, 0), lut_last);
                    sl[u] = s_lut[idx];
                }
                // ... then the exact fix-up on the bit pattern of the fp64 squared distance
                if (FIX > 0) {
#pragma unroll
                    for (int t = 0; t < FIX; ++t) {
                        long long thr[U];
#pragma unroll
                        for (int u = 0; u < U; ++u) thr[u] = s_upper[sl[u]];
#pragma unroll
                        for (int u = 0; u < U; ++u) sl[u] += (__double_as_longlong(d2[u]) > thr[u]) ? 1 : 0;
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        while (__double_as_longlong(d2[u]) > s_upper[sl[u]]) ++sl[u];
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (diagonal && !(jbase + j0 + u > ig)) sl[u] = 0;     // keep i < j only
                    slot[u] = sl[u] * NT;
                }
            };
            auto rmw = [&](const int (&slot)[U], const double (&dzv)[U]) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    double2 a = acc[slot[u]];
                    a.x += dzv[u];
                    a.y = __fma_rn(dzv[u], dzv[u], a.y);
                    acc[slot[u]] = a;
                    cnt[slot[u]] += 1u;
                }
            };
            int sA[U], sB[U];
            double dA[U], dB[U];
            // software pipeline: the slots of batch k+1 are computed while batch k is accumulated
            compute(0, sA, dA);
#pragma unroll 1
            for (int j0 = 0; j0 < TILE - 2 * U; j0 += 2 * U) {
                compute(j0 + U, sB, dB);
                rmw(sA, dA);
                compute(j0 + 2 * U, sA, dA);
                rmw(sB, dB);
            }
            compute(TILE - U, sB, dB);
            rmw(sA, dA);
            rmw(sB, dB);
            __syncthreads();                                  // everyone is done with buffer b
            if (tb_next < 0) break;
            tb = tb_next;
        }

        // ---- fold this item's private histograms (my z_i is fixed within the item) ------------
        //   sum z_i z_j   = sum z_i (z_i - dz) = c z_i^2 - z_i S1
        //   sum (z_i+z_j) = 2 c z_i - S1
        for (int s = 1; s < n_slots - 1; ++s) {
            const uint32_t c = cnt[s * NT];
            // any lane with something to add?  (most slots are empty for a given item)
            if (__ballot_sync(0xffffffffu, c != 0u) == 0u) continue;
            const double2 a = acc[s * NT];
            const double cd = (double)c;
            double v_sq = a.y;
            double v_prod = __fma_rn(-zi, a.x, cd * zi * zi);
            double v_val = __fma_rn(2.0 * cd, zi, -a.x);
            unsigned long long v_cnt = c;
            for (int o = 16; o; o >>= 1) {
                v_sq += __shfl_xor_sync(0xffffffffu, v_sq, o);
                v_prod += __shfl_xor_sync(0xffffffffu, v_prod, o);
                v_val += __shfl_xor_sync(0xffffffffu, v_val, o);
                v_cnt += __shfl_xor_sync(0xffffffffu, v_cnt, o);
            }
            if (lane == 0) {
                double *dst = my_partial + (size_t)s * 4;
                dst[0] += v_sq; dst[1] += v_prod; dst[2] += v_val;
                reinterpret_cast<unsigned long long *>(dst)[3] += v_cnt;
            }
            acc[s * NT] = make_double2(0.0, 0.0);
            cnt[s * NT] = 0u;
        }
        // trash slots are never read; reset them so their counters cannot overflow
        cnt[0] = 0u; cnt[(n_slots - 1) * NT] = 0u;
        acc[0] = make_double2(0.0, 0.0); acc[(n_slots - 1) * NT] = make_double2(0.0, 0.0);
    }
    if (tid == 0 && evaluated) atomicAdd(p.pairs_evaluated, evaluated);
}

// ---------------------------------------------------------------------------------------------
// Ring-window kernel: the fast path when tiles are small against the lag width.
//
// The full-histogram kernel above is limited to 128 threads per SM (one private histogram of
// n_lags + 2 slots per thread fills the shared memory), and it pays one 40-byte shared-memory
// read-modify-write per pair: ncu shows the shared-memory pipe, not the FP64 pipe, as its limit.
// This kernel removes both limits:
//  * RING.  A tile pair only touches the few lags between the smallest and largest distance of its
//    two bounding boxes, so a thread keeps a ring of RING slots (slot = lag & (RING-1)) plus two trash
//    slots instead of a full histogram: 680 B per thread -> 256 threads per CTA.  The CTA tracks the
//    range of lags touched since the last fold; before a tile pair that would stretch the range to
//    RING or more lags, every warp folds its rings into the per-warp global partials (lane l sums
//    ring slot l over the 32 threads of the warp).  A tile pair that alone spans >= RING lags is swept
//    once per window of RING lags.  No atomics: the result is deterministic.
//  * TWO-LAG REGISTERS.  Points are Hilbert sorted, so 8 consecutive points of a B tile are a few
//    hundred metres apart and their distances to a thread's own point fall into at most two adjacent
//    lags (lo, lo + 1) almost always.  A thread accumulates each group of 8 pairs in registers
//    (sum over all 8, sum over those in lo + 1) and touches shared memory twice per GROUP instead of
//    once per PAIR.  The rare pair outside {lo, lo + 1} takes a per-pair read-modify-write under a
//    warp vote.
//  * The first slot comes from a table indexed by the exponent and top 7 mantissa bits of the fp64
//    squared distance (integer ops only, no conversion / sqrt on the quarter-rate XU pipe); one or two
//    64-bit integer compares against the exact thresholds finish the assignment.
// ---------------------------------------------------------------------------------------------

template <int RING, int NT, int FIX, bool ELLIPSE>
__global__ void __launch_bounds__(NT, 1) pair_ring_kernel(const PairParams p)
{
    constexpr int U = 8;
    constexpr int A_TILES = NT / TILE;             // tiles per A sub-tile
    constexpr int TRASH_LO = RING, TRASH_HI = RING + 1;
    static_assert(NT % TILE == 0 && RING <= 32, "A sub-tile is whole tiles; one lane folds one ring slot");
    __shared__ __align__(16) double2 s_txy[2][TILE];
    __shared__ __align__(16) double s_tz[2][TILE];
    __shared__ __align__(16) double2 s_gs[2][TILE / 8];
    __shared__ long long s_upper[MAX_SLOTS];
    __shared__ uint16_t s_lut[LUT2_MAX];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ uint64_t s_live;
    __shared__ short2 s_range[CHUNK];              // lag-slot range [lo, hi] of each B tile of the chunk
    __shared__ double s_zi[NT];
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_slots = p.n_slots, n_lags = n_slots - 2;

    double2 *acc_all = reinterpret_cast<double2 *>(smem_raw);                         // [RING + 2][NT]
    uint32_t *cnt_all = reinterpret_cast<uint32_t *>(smem_raw + (size_t)(RING + 2) * NT * sizeof(double2));
    double2 *acc = acc_all + tid;
    uint32_t *cnt = cnt_all + tid;

    for (int s = 0; s < RING + 2; ++s) {
        acc[s * NT] = make_double2(0.0, 0.0);
        cnt[s * NT] = 0u;
    }
    for (int s = tid; s < n_slots; s += NT) s_upper[s] = __double_as_longlong(p.upper[s]);
    for (int s = tid; s < p.lut2_n; s += NT) s_lut[s] = p.lut2[s];
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int n_tiles = p.n_tiles;
    const int n_atiles = (n_tiles + A_TILES - 1) / A_TILES;
    const int n_chunks = (n_tiles + CHUNK - 1) / CHUNK;
    const int64_t n_items = (int64_t)n_atiles * n_chunks;
    uint32_t seq = 0;
    unsigned long long evaluated = 0;
    double *my_partial = p.partial + ((size_t)blockIdx.x * (NT / 32) + warp) * n_slots * 4;
    const int base_hi = p.lut2_base, lut_last = p.lut2_n - 1;
    const double w00 = p.w00, w01 = p.w01, w10 = p.w10, w11 = p.w11;
    // ELLIPSE: p.tile_box holds the boxes of the WHITENED coordinates W p; box_slack covers the rounding
    // difference between W (p_j - p_i) as the pairs compute it and W p_j - W p_i
    const double slack = p.box_slack;

    // Static, deterministic load balance: a pre-pass orders the items by descending cost (live B tiles) and drops
    // the empty ones; dealing that list round-robin over CTAs and ranks gives every CTA one item of each cost stratum.
    const int64_t n_work = p.item_list ? (int64_t)*p.n_live_items : n_items;
    for (int64_t k = ((int64_t)blockIdx.x) * p.shard_count + p.shard_index; k < n_work;
         k += (int64_t)gridDim.x * p.shard_count) {
        const int64_t item = p.item_list ? (int64_t)p.item_list[k] : k;
        const int64_t a0 = (item / n_chunks) * NT;            // first point of my A sub-tile
        const int ta = (int)(a0 / TILE);                      // first of its tiles
        const int ta_last = min(ta + A_TILES, n_tiles) - 1;
        const int tb0 = (int)(item % n_chunks) * CHUNK;
        const int tb1 = min(tb0 + CHUNK, n_tiles);
        const int valid_a = (int)max((int64_t)0, min((int64_t)NT, p.n - a0));
        if (tb1 <= ta || valid_a == 0) continue;
        const int tb_first = max(tb0, ta);

        // bounding box of the A sub-tile
        double4 abox = p.tile_box[ta];
        for (int t = ta + 1; t <= ta_last; ++t) {
            const double4 o = p.tile_box[t];
            abox.x = fmin(abox.x, o.x); abox.y = fmin(abox.y, o.y);
            abox.z = fmax(abox.z, o.z); abox.w = fmax(abox.w, o.w);
        }
        // ---- per B tile: can it hold a pair within range, and which lags can it touch? ----------
        if (warp == 0) {
            uint64_t m = 0;
            for (int h = 0; h < CHUNK / 32; ++h) {
                const int tb = tb0 + h * 32 + lane;
                bool live = false;
                if (tb >= tb_first && tb < tb1) {
                    const double4 bbox = p.tile_box[tb];
                    double dmin;
                    if (ELLIPSE) {
                        const double gx = fmax(0.0, fmax(abox.x - bbox.z, bbox.x - abox.z) - slack);
                        const double gy = fmax(0.0, fmax(abox.y - bbox.w, bbox.y - abox.w) - slack);
                        dmin = (gx * gx + gy * gy) * (1.0 - 1e-12);
                    } else dmin = box_min_d2(abox, bbox);
                    live = dmin <= p.max_d2;
                    if (live) {
                        // every pair of the two boxes has dmin <= d2 <= dmax in the kernel's own arithmetic
                        // (fp subtraction, multiplication and addition are monotone; the ellipse bounds carry slack)
                        const double dxm = fmax(abox.z - bbox.x, bbox.z - abox.x) + (ELLIPSE ? slack : 0.0);
                        const double dym = fmax(abox.w - bbox.y, bbox.w - abox.y) + (ELLIPSE ? slack : 0.0);
                        const double dmax = ELLIPSE ? (dxm * dxm + dym * dym) * (1.0 + 1e-12)
                                                    : __dadd_rn(__dmul_rn(dxm, dxm), __dmul_rn(dym, dym));
                        int a = 0, b = n_slots - 1;           // first slot whose threshold is >= dmin
                        const long long v = __double_as_longlong(dmin);
                        while (a < b) { const int mid = (a + b) >> 1; if (v <= s_upper[mid]) b = mid; else a = mid + 1; }
                        const int lo = max(a, 1);
                        a = 0; b = n_slots - 1;
                        const long long w = __double_as_longlong(dmax);
                        while (a < b) { const int mid = (a + b) >> 1; if (w <= s_upper[mid]) b = mid; else a = mid + 1; }
                        const int hi = min(a, n_lags);
                        s_range[h * 32 + lane] = make_short2((short)lo, (short)hi);
                    }
                }
                m |= (uint64_t)__ballot_sync(0xffffffffu, live) << (32 * h);
            }
            if (lane == 0) s_live = m;
        }
        __syncthreads();
        uint64_t live = s_live;
        if (live == 0) { __syncthreads(); continue; }

        const int ig = (int)a0 + tid;
        const int il = min(ig, p.n_tiles * TILE - 1);         // the last A sub-tile may reach past the padded arrays (rows >= n never count)
        PIB_DBG_ASSERT(il >= 0);
        const double2 pa = p.xy[il];
        const double xi = pa.x, yi = pa.y, zi = p.z[il];
        s_zi[tid] = zi;
        const double z8 = zi * 8.0, zm2 = zi * -2.0, zq8 = (zi * zi) * 8.0;
        int win_lo = 0x7fff, win_hi = 0;                      // lag slots touched since the last fold (empty)

        // fold the rings of my warp into its global partials (lane l owns ring slot l)
        auto fold = [&]() {
            __syncwarp();
            if (win_hi >= win_lo && lane < RING) {
                const int abs_slot = win_lo + ((lane - win_lo) & (RING - 1));       // the window's slot that maps to ring `lane`
                double v_sq = 0.0, v_prod = 0.0, v_val = 0.0;
                unsigned long long v_cnt = 0;
                const int row = lane * NT + warp * 32;
                for (int t = 0; t < 32; ++t) {
                    const int tt = (t + lane) & 31;           // rotate so the lanes of a quarter-warp hit different banks
                    const uint32_t c = cnt_all[row + tt];
                    if (c) {
                        const double2 a = acc_all[row + tt];
                        const double z = s_zi[warp * 32 + tt], cd = (double)c;
                        //   sum z_i z_j = c z_i^2 - z_i S1 ;  sum (z_i + z_j) = 2 c z_i - S1   (z_j = z_i - dz)
                        v_sq += a.y;
                        v_prod += __fma_rn(-z, a.x, cd * z * z);
                        v_val += __fma_rn(2.0 * cd, z, -a.x);
                        v_cnt += c;
                        acc_all[row + tt] = make_double2(0.0, 0.0);
                        cnt_all[row + tt] = 0u;
                    }
                }
                if (v_cnt && abs_slot <= win_hi) {
                    double *dst = my_partial + (size_t)abs_slot * 4;
                    dst[0] += v_sq; dst[1] += v_prod; dst[2] += v_val;
                    reinterpret_cast<unsigned long long *>(dst)[3] += v_cnt;
                }
            }
            __syncwarp();
            win_lo = 0x7fff; win_hi = 0;
        };

        int tb = tb0 + __ffsll((long long)live) - 1;
        live &= live - 1;
        if (tid == 0) {
            const uint32_t b = seq & 1;
            mbar_expect_tx(&s_bar[b], TILE * (sizeof(double2) + sizeof(double)) + (TILE / 8) * sizeof(double2));
            tma_load_1d(s_txy[b], p.xy + (size_t)tb * TILE, TILE * sizeof(double2), &s_bar[b]);
            tma_load_1d(s_tz[b], p.z + (size_t)tb * TILE, TILE * sizeof(double), &s_bar[b]);
            tma_load_1d(s_gs[b], p.gsum + (size_t)tb * (TILE / 8), (TILE / 8) * sizeof(double2), &s_bar[b]);
        }
        while (true) {
            const uint32_t b = seq & 1, phase = (seq >> 1) & 1;
            const int tb_next = live ? tb0 + __ffsll((long long)live) - 1 : -1;
            live &= live - 1;
            if (tid == 0 && tb_next >= 0) {
                const uint32_t nb = b ^ 1;
                mbar_expect_tx(&s_bar[nb], TILE * (sizeof(double2) + sizeof(double)) + (TILE / 8) * sizeof(double2));
                tma_load_1d(s_txy[nb], p.xy + (size_t)tb_next * TILE, TILE * sizeof(double2), &s_bar[nb]);
                tma_load_1d(s_tz[nb], p.z + (size_t)tb_next * TILE, TILE * sizeof(double), &s_bar[nb]);
                tma_load_1d(s_gs[nb], p.gsum + (size_t)tb_next * (TILE / 8), (TILE / 8) * sizeof(double2), &s_bar[nb]);
            }
            // ---- ring window bookkeeping (uniform over the CTA) ---------------------------------
            const short2 rng = s_range[tb - tb0];
            const bool wide = (rng.y - rng.x >= RING);
            if (!wide && rng.y >= rng.x) {
                const int nlo = min(win_lo, (int)rng.x), nhi = max(win_hi, (int)rng.y);
                if (nhi - nlo >= RING) { fold(); win_lo = rng.x; win_hi = rng.y; }
                else { win_lo = nlo; win_hi = nhi; }
            }
            mbar_wait(&s_bar[b], phase);
            ++seq;

            const double2 *bxy = s_txy[b];
            const double *bz = s_tz[b];
            const double2 *bgs = s_gs[b];
            const bool diagonal = (tb <= ta_last);
            const int jbase = tb * TILE;
            if (tid == 0) {
                const long long vb = min((int64_t)TILE, p.n - (int64_t)tb * TILE), va = valid_a;
                unsigned long long np = 0;                    // pairs with j > i, i in [a0, a0+va), j in [jbase, jbase+vb)
                if (!diagonal) np = (unsigned long long)(va * vb);
                else
                    for (long long i = a0; i < a0 + va; ++i) {
                        const long long first = max(i + 1, (long long)jbase);
                        if (jbase + vb > first) np += (unsigned long long)(jbase + vb - first);
                    }
                evaluated += np;
            }

            // (the lambdas below are forced inline: left to its own budget the compiler kept some of them out of line in
            // some instantiations, which puts every captured variable into a 1 KB stack frame and costs a factor of 4)
            int wlo = 1;                                      // lag slots accepted by the ring: [wlo, wlo + wlen)
            unsigned wlen = (unsigned)n_lags;
            auto ring_of = [&](int s, int trash) __attribute__((always_inline)) -> int {
                return (((unsigned)(s - wlo) < wlen) ? (s & (RING - 1)) : trash) * NT;
            };
            // ---- stage 1: exact lag slot and dz of U consecutive pairs (independent chains) -------
            auto compute = [&](int j0, int (&sl)[U], double (&dzv)[U]) __attribute__((always_inline)) {
                double d2[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const double2 pj = bxy[j0 + u];
                    dzv[u] = __dsub_rn(zi, bz[j0 + u]);
                    if (ELLIPSE) {
                        // vector_distance = other - centre (angular.py:448); W @ d as OpenBLAS dgemm evaluates it:
                        // fma(W[r][1], dy, W[r][0]*dx); einsum: n0*n0 + n1*n1 un-fused (angular.py:666-668)
                        const double dx = __dsub_rn(pj.x, xi), dy = __dsub_rn(pj.y, yi);
                        const double n0 = __fma_rn(w01, dy, __dmul_rn(w00, dx));
                        const double n1 = __fma_rn(w11, dy, __dmul_rn(w10, dx));
                        d2[u] = __dadd_rn(__dmul_rn(n0, n0), __dmul_rn(n1, n1));
                    } else {
                        // scipy cdist euclidean: sqrt(dx*dx + dy*dy), nothing fused (distance/point.py:56)
                        const double dx = __dsub_rn(xi, pj.x), dy = __dsub_rn(yi, pj.y);
                        d2[u] = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    int idx = (__double2hiint(d2[u]) - base_hi) >> LUT2_SHIFT;
                    idx = min(max(idx, 0), lut_last);
                    sl[u] = s_lut[idx];
                }
                if (FIX > 0) {
#pragma unroll
                    for (int t = 0; t < FIX; ++t) {
                        long long thr[U];
#pragma unroll
                        for (int u = 0; u < U; ++u) thr[u] = s_upper[sl[u]];
#pragma unroll
                        for (int u = 0; u < U; ++u) sl[u] += (__double_as_longlong(d2[u]) > thr[u]) ? 1 : 0;
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        while (__double_as_longlong(d2[u]) > s_upper[sl[u]]) ++sl[u];
                }
                if (diagonal) {
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (!(jbase + j0 + u > ig)) sl[u] = 0;                         // keep i < j only
                }
            };
            // ---- stage 2: accumulate the group in registers, touch the ring twice ----------------
            auto consume = [&](int (&sl)[U], double (&dzv)[U]) __attribute__((always_inline)) {
                int lo = sl[0], mx = sl[0];
#pragma unroll
                for (int u = 1; u < U; ++u) { lo = min(lo, sl[u]); mx = max(mx, sl[u]); }
                int n_out = 0;
                if (__any_sync(0xffffffffu, mx > lo + 1)) {                            // rare: a pair outside {lo, lo+1}
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (sl[u] > lo + 1) {
                            const int r = ring_of(sl[u], TRASH_HI);
                            double2 a = acc[r];
                            a.x += dzv[u];
                            a.y = __fma_rn(dzv[u], dzv[u], a.y);
                            acc[r] = a;
                            cnt[r] += 1u;
                            dzv[u] = 0.0; sl[u] = lo; ++n_out;
                        }
                    }
                }
                double dh[U];
                int ch = 0;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const bool is_hi = (sl[u] != lo);
                    dh[u] = is_hi ? dzv[u] : 0.0;
                    ch += is_hi ? 1 : 0;
                }
                // all-8 and lo+1-only sums, pairwise trees to keep the dependency chains short
                const double a1 = ((dzv[0] + dzv[1]) + (dzv[2] + dzv[3])) + ((dzv[4] + dzv[5]) + (dzv[6] + dzv[7]));
                const double h1 = ((dh[0] + dh[1]) + (dh[2] + dh[3])) + ((dh[4] + dh[5]) + (dh[6] + dh[7]));
                double a2e = dzv[0] * dzv[0], a2o = dzv[1] * dzv[1], h2e = dh[0] * dh[0], h2o = dh[1] * dh[1];
#pragma unroll
                for (int u = 2; u < U; u += 2) {
                    a2e = __fma_rn(dzv[u], dzv[u], a2e); a2o = __fma_rn(dzv[u + 1], dzv[u + 1], a2o);
                    h2e = __fma_rn(dh[u], dh[u], h2e);   h2o = __fma_rn(dh[u + 1], dh[u + 1], h2o);
                }
                const double a2 = a2e + a2o, h2 = h2e + h2o;
                const int r_lo = ring_of(lo, TRASH_LO), r_hi = ring_of(lo + 1, TRASH_HI);   // never equal
                double2 v_lo = acc[r_lo], v_hi = acc[r_hi];
                const uint32_t c_lo = cnt[r_lo], c_hi = cnt[r_hi];
                v_lo.x += a1 - h1; v_lo.y += a2 - h2;
                v_hi.x += h1;      v_hi.y += h2;
                acc[r_lo] = v_lo; acc[r_hi] = v_hi;
                cnt[r_lo] = c_lo + (uint32_t)(U - ch - n_out);
                cnt[r_hi] = c_hi + (uint32_t)ch;
            };
            // ---- group path (tile pairs off the diagonal): ONE table lookup per group of U pairs ----------
            // The smallest high word of the U squared distances, with a zero low word, is a lower bound of all
            // of them; its slot `lo` (same table + threshold steps as above) is therefore <= every pair's slot and
            // equals the smallest one unless a threshold falls into the 2^-20 relative gap.  Each pair then needs a
            // single fp64 compare against the threshold between lo and lo + 1.  A group whose largest high word
            // could exceed the NEXT threshold takes the per-pair slow path under a warp vote (exact slot search).
            struct Grp { double a1, h1, a2, h2; int lo, ch, nout; };
            auto compute2 = [&](int j0, Grp &g) __attribute__((always_inline)) {
                double d2[U], dzv[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const double2 pj = bxy[j0 + u];
                    dzv[u] = __dsub_rn(zi, bz[j0 + u]);
                    if (ELLIPSE) {
                        const double dx = __dsub_rn(pj.x, xi), dy = __dsub_rn(pj.y, yi);
                        const double n0 = __fma_rn(w01, dy, __dmul_rn(w00, dx));
                        const double n1 = __fma_rn(w11, dy, __dmul_rn(w10, dx));
                        d2[u] = __dadd_rn(__dmul_rn(n0, n0), __dmul_rn(n1, n1));
                    } else {
                        const double dx = __dsub_rn(xi, pj.x), dy = __dsub_rn(yi, pj.y);
                        d2[u] = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                    }
                }
                int mn = __double2hiint(d2[0]), mx = mn;
#pragma unroll
                for (int u = 1; u < U; ++u) {
                    const int h = __double2hiint(d2[u]);
                    mn = min(mn, h); mx = max(mx, h);
                }
                int idx = (mn - base_hi) >> LUT2_SHIFT;
                idx = min(max(idx, 0), lut_last);
                int lo = s_lut[idx];
                const int *upper_hi = reinterpret_cast<const int *>(s_upper) + 1;      // high words of the thresholds
                if (FIX > 0) {
#pragma unroll
                    for (int t = 0; t < FIX; ++t) lo += (mn > upper_hi[2 * lo]) ? 1 : 0;   // bound = (mn << 32)
                } else {
                    while (mn > upper_hi[2 * lo]) ++lo;
                }
                const int lo1 = min(lo + 1, n_slots - 1);
                const double t_lo = __longlong_as_double(s_upper[lo]);
                const long long t_hi_bits = s_upper[lo1];
                int nout = 0;
                // sum of dz over the whole group from the pre-summed z of the group: 8 z_i - sum z_j (one fp64 op
                // instead of seven; this sum only feeds the covariance terms, where it meets c z_i^2 - z_i * sum)
                // ... and the sum of dz^2 from the pre-summed z and z^2: 8 z_i^2 - 2 z_i sum z_j + sum z_j^2.  The z values
                // are centred by the host side, so the cancellation costs eps * var(z) / (2 gamma(h)) relative — far
                // below the 1e-9 bar for any field with gamma(h) > 1e-7 var(z)
                const double2 gs = bgs[j0 >> 3];
                double a1 = __dsub_rn(z8, gs.x);
                double a2 = __fma_rn(zm2, gs.x, zq8) + gs.y;
                if (__any_sync(0xffffffffu, mx + 1 > (int)(t_hi_bits >> 32))) {            // rare: maybe beyond lo + 1
                    const double t_hi = __longlong_as_double(t_hi_bits);
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (d2[u] > t_hi) {
                            int sx = lo1 + 1;
                            while (__double_as_longlong(d2[u]) > s_upper[sx]) ++sx;
                            const int r = ring_of(sx, TRASH_HI);
                            double2 a = acc[r];
                            a.x += dzv[u];
                            a.y = __fma_rn(dzv[u], dzv[u], a.y);
                            acc[r] = a;
                            cnt[r] += 1u;
                            a1 -= dzv[u];
                            a2 = __fma_rn(-dzv[u], dzv[u], a2);
                            dzv[u] = 0.0; d2[u] = -1.0; ++nout;
                        }
                    }
                }
                double dh[U];
                unsigned hi_mask = 0;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const bool is_hi = d2[u] > t_lo;
                    // one select instead of two: a "lo" pair keeps the low word of dz under a zero high word, a
                    // subnormal below 1e-308 whose contribution to the fp64 sums is far below one ulp
                    dh[u] = __hiloint2double(is_hi ? __double2hiint(dzv[u]) : 0, __double2loint(dzv[u]));
                    if (is_hi) hi_mask |= 1u << u;                                        // one predicated OR per pair
                }
                const int ch = __popc(hi_mask);
                g.a1 = a1;
                g.h1 = ((dh[0] + dh[1]) + (dh[2] + dh[3])) + ((dh[4] + dh[5]) + (dh[6] + dh[7]));
                double h2e = dh[0] * dh[0], h2o = dh[1] * dh[1];
#pragma unroll
                for (int u = 2; u < U; u += 2) {
                    h2e = __fma_rn(dh[u], dh[u], h2e);   h2o = __fma_rn(dh[u + 1], dh[u + 1], h2o);
                }
                g.a2 = a2; g.h2 = h2e + h2o;
                g.lo = lo; g.ch = ch; g.nout = nout;
            };
            auto consume2 = [&](const Grp &g) __attribute__((always_inline)) {
                const int r_lo = ring_of(g.lo, TRASH_LO), r_hi = ring_of(g.lo + 1, TRASH_HI);   // never equal
                double2 v_lo = acc[r_lo], v_hi = acc[r_hi];
                const uint32_t c_lo = cnt[r_lo], c_hi = cnt[r_hi];
                v_lo.x += g.a1 - g.h1; v_lo.y += g.a2 - g.h2;
                v_hi.x += g.h1;        v_hi.y += g.h2;
                acc[r_lo] = v_lo; acc[r_hi] = v_hi;
                cnt[r_lo] = c_lo + (uint32_t)(U - g.ch - g.nout);
                cnt[r_hi] = c_hi + (uint32_t)g.ch;
            };
            int sA[U], sB[U];
            double dA[U], dB[U];
            Grp gA, gB;
            auto sweep = [&]() __attribute__((always_inline)) {                              // all 128 points of the B tile, software pipelined
                if (diagonal || !p.group_path) {
                    compute(0, sA, dA);
#pragma unroll 1
                    for (int j0 = 0; j0 < TILE - 2 * U; j0 += 2 * U) {
                        compute(j0 + U, sB, dB);
                        consume(sA, dA);
                        compute(j0 + 2 * U, sA, dA);
                        consume(sB, dB);
                    }
                    compute(TILE - U, sB, dB);
                    consume(sA, dA);
                    consume(sB, dB);
                } else {
                    compute2(0, gA);
#pragma unroll 1
                    for (int j0 = 0; j0 < TILE - 2 * U; j0 += 2 * U) {
                        compute2(j0 + U, gB);
                        consume2(gA);
                        compute2(j0 + 2 * U, gA);
                        consume2(gB);
                    }
                    compute2(TILE - U, gB);
                    consume2(gA);
                    consume2(gB);
                }
            };
            if (!wide) {
                sweep();
            } else {
                // this tile pair alone spans >= RING lags: sweep it once per window of RING lags, folding in
                // between (no atomics, so the result stays deterministic)
                fold();
                for (int w0 = rng.x; w0 <= rng.y; w0 += RING) {
                    wlo = w0; wlen = (unsigned)(min(w0 + RING - 1, (int)rng.y) - w0 + 1);
                    sweep();
                    win_lo = w0; win_hi = w0 + (int)wlen - 1;
                    fold();
                }
            }
            __syncthreads();
            if (tb_next < 0) break;
            tb = tb_next;
        }
        fold();                                               // my z_i changes with the next item
        cnt[TRASH_LO * NT] = 0u; cnt[TRASH_HI * NT] = 0u;
        acc[TRASH_LO * NT] = make_double2(0.0, 0.0); acc[TRASH_HI * NT] = make_double2(0.0, 0.0);
    }
    if (tid == 0 && evaluated) atomicAdd(p.pairs_evaluated, evaluated);
}

}  // namespace pib
#include "pair_ring6.cuh"
namespace pib {

// Cost of every ring-kernel work item = number of B tiles of its chunk that can hold a pair within range
// (the same box test as the kernel's).  One warp per item; key = CHUNK - cost so that an ascending radix sort
// puts the most expensive items first.
constexpr int ITEM_COST_BITS = 10, ITEM_COST_MAX = (1 << ITEM_COST_BITS) - 1;
template <bool ELLIPSE>
__global__ void item_cost_kernel(const PairParams p, int a_tiles, int64_t n_items, uint32_t *__restrict__ keys,
                                 int32_t *__restrict__ vals, unsigned *__restrict__ n_live)
{
    const int64_t item = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (item >= n_items) return;
    const int n_tiles = p.n_tiles;
    const int chunk = p.chunk;
    const int n_chunks = (n_tiles + chunk - 1) / chunk;
    const int ta = (int)(item / n_chunks) * a_tiles;
    const int ta_last = min(ta + a_tiles, n_tiles) - 1;
    const int tb0 = (int)(item % n_chunks) * chunk;
    const int tb1 = min(tb0 + chunk, n_tiles);
    // cost of the item ~ sum over its live B tiles of (1 + number of the item's A tiles within range of the B tile): the
    // kernels cull per warp, so an item at the rim of the range costs a fraction of one in the middle.  Tiles on the
    // diagonal take the per-pair path: three times the weight.  (<= CHUNK * (1 + 3 * 8) ... clamped to 10 bits.)
    int cost = 0;
    if (tb1 > ta && (int64_t)ta * TILE < p.n) {
        double4 abox = p.tile_box[ta];
        for (int t = ta + 1; t <= ta_last; ++t) {
            const double4 o = p.tile_box[t];
            abox.x = fmin(abox.x, o.x); abox.y = fmin(abox.y, o.y);
            abox.z = fmax(abox.z, o.z); abox.w = fmax(abox.w, o.w);
        }
        const int tb_first = max(tb0, ta);
        auto min_d2 = [&](const double4 &A, const double4 &B) {
            if (ELLIPSE) {
                const double gx = fmax(0.0, fmax(A.x - B.z, B.x - A.z) - p.box_slack);
                const double gy = fmax(0.0, fmax(A.y - B.w, B.y - A.w) - p.box_slack);
                return (gx * gx + gy * gy) * (1.0 - 1e-12);
            }
            return box_min_d2(A, B);
        };
        for (int h = 0; h < (chunk + 31) / 32; ++h) {
            const int tb = tb0 + h * 32 + lane;
            int c = 0;
            if (tb >= tb_first && tb < tb1) {
                const double4 bbox = p.tile_box[tb];
                if (min_d2(abox, bbox) <= p.max_d2) {
                    c = 1;
                    for (int t = ta; t <= ta_last && t <= tb; ++t)
                        if (min_d2(p.tile_box[t], bbox) <= p.max_d2) c += (tb <= ta_last) ? 3 : 1;
                }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            cost += c;
        }
        cost = min(cost, ITEM_COST_MAX);
    }
    if (lane == 0) {
        keys[item] = (uint32_t)(ITEM_COST_MAX - cost);      // ascending sort = most expensive first, dead items last
        vals[item] = (int32_t)item;
        if (cost) atomicAdd(n_live, 1u);
    }
}

// mean of z over the real points in a fixed summation order (256 blocks of partial sums, then one block):
// the ring kernel works on centred values
__global__ void __launch_bounds__(256) mean_partial_kernel(const double *__restrict__ z, int64_t n, double *__restrict__ partial)
{
    __shared__ double sh[256];
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) s += z[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

__global__ void __launch_bounds__(256) mean_final_kernel(const double *__restrict__ partial, int n_partial, int64_t n,
                                                         double *__restrict__ mean)
{
    __shared__ double sh[256];
    sh[threadIdx.x] = ((int)threadIdx.x < n_partial) ? partial[threadIdx.x] : 0.0;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *mean = sh[0] / (double)n;
}

// centred copy of z and, per group of G consecutive sorted points (pads included, as the pair kernel sees them),
// the sums of the centred values and of their squares
template <int G>
__global__ void group_sum_kernel(const double *__restrict__ z, const double *__restrict__ mean, int64_t n,
                                 int64_t n_groups, double *__restrict__ zc, double2 *__restrict__ gsum)
{
    const int64_t gidx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (gidx >= n_groups) return;
    const double m = *mean;
    double v[G];
#pragma unroll
    for (int u = 0; u < G; ++u) {
        // padding rows sit at the mean (centred 0): their pairs only ever reach the trash slots, but the two-lag
        // sums form "all minus second lag", and a huge padding value would cancel there (data with a large offset)
        v[u] = (gidx * G + u < n) ? z[gidx * G + u] - m : 0.0;
        zc[gidx * G + u] = v[u];
    }
    double s1 = 0.0, qe = 0.0, qo = 0.0;
#pragma unroll
    for (int u = 0; u < G; u += 8)
        s1 += ((v[u] + v[u + 1]) + (v[u + 2] + v[u + 3])) + ((v[u + 4] + v[u + 5]) + (v[u + 6] + v[u + 7]));
#pragma unroll
    for (int u = 0; u < G; u += 2) { qe = __fma_rn(v[u], v[u], qe); qo = __fma_rn(v[u + 1], v[u + 1], qo); }
    gsum[gidx] = make_double2(s1, qe + qo);
}

// fp32 screening tiles of pair_ring6_kernel: coordinates relative to the centre of the tile's box, rounded to single
// precision, two points per 16-byte entry {x0, x1, y0, y1}; the origin follows the TILE / 2 entries (one thread per entry)
__global__ void tile_f32_kernel(const double2 *__restrict__ xy, const double4 *__restrict__ tile_box, int64_t n_tiles,
                                unsigned char *__restrict__ out)
{
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= n_tiles * (TILE / 2)) return;
    const int64_t t = e / (TILE / 2);
    const int m = (int)(e % (TILE / 2));
    const double4 box = tile_box[t];
    // (an empty tile has an infinite box: its entries are never used, the kernel only screens full tiles)
    const double ox = __dmul_rn(0.5, __dadd_rn(box.x, box.z)), oy = __dmul_rn(0.5, __dadd_rn(box.y, box.w));
    const double2 a = xy[t * TILE + 2 * m], b = xy[t * TILE + 2 * m + 1];
    unsigned char *rec = out + (size_t)t * TF32_TILE_BYTES;
    reinterpret_cast<float4 *>(rec)[m] = make_float4(__double2float_rn(__dsub_rn(a.x, ox)), __double2float_rn(__dsub_rn(b.x, ox)),
                                                     __double2float_rn(__dsub_rn(a.y, oy)), __double2float_rn(__dsub_rn(b.y, oy)));
    if (m == 0) *reinterpret_cast<double2 *>(rec + TILE * 8) = make_double2(ox, oy);
}

// bounding boxes of the whitened coordinates (u, v) = W p of every tile (one warp per tile)
__global__ void tile_box_whitened_kernel(const double2 *__restrict__ xy, int64_t n, int64_t n_tiles,
                                         double w00, double w01, double w10, double w11, double4 *__restrict__ box)
{
    const int64_t t = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= n_tiles) return;
    double umin = INFINITY, vmin = INFINITY, umax = -INFINITY, vmax = -INFINITY;
    for (int k = lane; k < TILE; k += 32) {
        const int64_t i = t * TILE + k;
        if (i < n) {
            const double2 q = xy[i];
            const double u = w00 * q.x + w01 * q.y, v = w10 * q.x + w11 * q.y;
            umin = fmin(umin, u); umax = fmax(umax, u);
            vmin = fmin(vmin, v); vmax = fmax(vmax, v);
        }
    }
    for (int o = 16; o; o >>= 1) {
        umin = fmin(umin, __shfl_xor_sync(0xffffffffu, umin, o));
        vmin = fmin(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
        umax = fmax(umax, __shfl_xor_sync(0xffffffffu, umax, o));
        vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    }
    if (lane == 0) box[t] = make_double4(umin, vmin, umax, vmax);
}

// Fixed-order sum of the per-warp partials -> outputs.  One 128-thread block per lag: thread t sums parts
// t, t + 128, ... in order, then a fixed shared-memory tree combines the 128 sub-sums (deterministic).
__global__ void __launch_bounds__(128) reduce_partials_kernel(const double *__restrict__ partial, int n_parts, int n_slots,
                                                              int n_lags, const double *__restrict__ mean,
                                                              double *__restrict__ sum_sq,
                                                              double *__restrict__ sum_prod, double *__restrict__ sum_val,
                                                              int64_t *__restrict__ count, const int *__restrict__ run_flag = nullptr)
{
    if (run_flag && *run_flag == 0) return;
    __shared__ double sh[3][128];
    __shared__ unsigned long long shc[128];
    const int b = blockIdx.x, t = threadIdx.x;
    if (b >= n_lags) return;
    double s0 = 0, s1 = 0, s2 = 0;
    unsigned long long c = 0;
    for (int q = t; q < n_parts; q += 128) {
        const double *src = partial + ((size_t)q * n_slots + (b + 1)) * 4;
        s0 += src[0]; s1 += src[1]; s2 += src[2];
        c += reinterpret_cast<const unsigned long long *>(src)[3];
    }
    sh[0][t] = s0; sh[1][t] = s1; sh[2][t] = s2; shc[t] = c;
    __syncthreads();
    for (int o = 64; o; o >>= 1) {
        if (t < o) {
            sh[0][t] += sh[0][t + o]; sh[1][t] += sh[1][t + o]; sh[2][t] += sh[2][t + o]; shc[t] += shc[t + o];
        }
        __syncthreads();
    }
    if (t == 0) {
        double prod = sh[1][0], val = sh[2][0];
        if (mean) {
            // the ring kernel summed centred values z' = z - m:  sum (z_i + z_j) = sum' + 2 c m,
            // sum z_i z_j = sum' z'_i z'_j + m sum' (z'_i + z'_j) + c m^2
            const double m = *mean, c = (double)shc[0];
            prod = prod + m * val + c * m * m;
            val = val + 2.0 * c * m;
        }
        sum_sq[b] = sh[0][0]; sum_prod[b] = prod; sum_val[b] = val; count[b] = (int64_t)shc[0];
    }
}

// Cancellation guard of the z_j-form / group-sum kernels: they form sum dz^2 from sums of (centred) z and z^2, which
// costs eps * var(z) / (2 gamma(h)) relative.  flag = 1 when some lag has gamma(h) = sum_sq / (2 count) < 1e-6 var(z):
// the host side has then already queued a conditional re-run with the per-pair dz-form kernel (pair_kernel).
__global__ void __launch_bounds__(1024) cancellation_guard_kernel(const double2 *__restrict__ gsum, int64_t n_groups, int64_t n,
                                                                  const double *__restrict__ sum_sq,
                                                                  const int64_t *__restrict__ count, int n_lags,
                                                                  int *__restrict__ flag)
{
    __shared__ double sh[1024];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n_groups; i += 1024) s += gsum[i].y;      // sum of the squared centred values
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 512; o; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    const double var = sh[0] / (double)n;
    bool bad = false;
    for (int b = threadIdx.x; b < n_lags; b += 1024)
        if (count[b] > 0 && sum_sq[b] < 1e-6 * var * 2.0 * (double)count[b]) bad = true;
    if (__syncthreads_or(bad) && threadIdx.x == 0) *flag = 1;
}

// ---------------------------------------------------------------------------------------------
// Generic kernel: any number of lags, irregular ellipse limits (lower[b] != edges[b-1]).
// One thread per point i, all j > i, every lag tested like the reference's masks; global fp64
// REDs.  Slow and simple; only used outside the envelope of pair_kernel.
// ---------------------------------------------------------------------------------------------
__global__ void pair_generic_kernel(const double2 *__restrict__ xy, const double *__restrict__ z, int64_t n,
                                    const double *__restrict__ up_d2, const double *__restrict__ low_d2, int n_lags,
                                    bool ellipse, double w00, double w01, double w10, double w11,
                                    int shard_index, int shard_count,
                                    double *sum_sq, double *sum_prod, double *sum_val, unsigned long long *count,
                                    unsigned long long *pairs_evaluated)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * shard_count;
    unsigned long long evaluated = 0;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * shard_count + shard_index; i < n; i += stride) {
        const double2 pi = xy[i];
        const double zi = z[i];
        for (int64_t j = i + 1; j < n; ++j) {
            const double2 pj = xy[j];
            double d2;
            if (ellipse) {
                const double dx = __dsub_rn(pj.x, pi.x), dy = __dsub_rn(pj.y, pi.y);
                const double n0 = __fma_rn(w01, dy, __dmul_rn(w00, dx));
                const double n1 = __fma_rn(w11, dy, __dmul_rn(w10, dx));
                d2 = __dadd_rn(__dmul_rn(n0, n0), __dmul_rn(n1, n1));
            } else {
                const double dx = __dsub_rn(pi.x, pj.x), dy = __dsub_rn(pi.y, pj.y);
                d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
            }
            ++evaluated;
            if (!(d2 > 0.0)) continue;
            const double zj = z[j], dz = zi - zj;
            for (int b = 0; b < n_lags; ++b) {
                if (d2 <= up_d2[b] && d2 > low_d2[b]) {
                    atomicAdd(&sum_sq[b], dz * dz);
                    atomicAdd(&sum_prod[b], zi * zj);
                    atomicAdd(&sum_val[b], zi + zj);
                    atomicAdd(&count[b], 1ull);
                }
            }
        }
    }
    atomicAdd(pairs_evaluated, evaluated);
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
// largest double T with sqrt(T) <= e (correctly rounded sqrt), e >= 0
static double sq_threshold(double e)
{
    if (!(e > 0)) return 0.0;
    double t = e * e;
    while (std::sqrt(t) > e) t = std::nextafter(t, 0.0);
    while (std::sqrt(std::nextafter(t, INFINITY)) <= e) t = std::nextafter(t, INFINITY);
    return t;
}

template <int NT, bool ELLIPSE, int FIX>
static int launch_pair_kernel_fix(const PairParams &p, size_t smem, int grid, cudaStream_t stream)
{
    PIB_CUDA(cudaFuncSetAttribute(pair_kernel<NT, ELLIPSE, FIX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pair_kernel<NT, ELLIPSE, FIX><<<grid, NT, smem, stream>>>(p);
    PIB_CUDA(cudaGetLastError());
    return PIB_OK;
}

template <int NT, bool ELLIPSE>
static int launch_pair_kernel(const PairParams &p, int fix_steps, int grid, cudaStream_t stream)
{
    const size_t smem = pair_smem_bytes<NT>(p.n_slots);
    if (fix_steps <= 1) return launch_pair_kernel_fix<NT, ELLIPSE, 1>(p, smem, grid, stream);
    if (fix_steps <= 2) return launch_pair_kernel_fix<NT, ELLIPSE, 2>(p, smem, grid, stream);
    return launch_pair_kernel_fix<NT, ELLIPSE, 0>(p, smem, grid, stream);
}

// ALL = every (multi-step threshold fix-up, ellipse) combination; experiment variants are only built for the
// plain omnidirectional case
template <int RING, int NT, int APT, int U, int NL, bool ALL>
static int launch_pair_ring6_t(const PairParams &p, bool multi_fix, bool ellipse, int grid, cudaStream_t stream)
{
    const size_t smem = ring6_smem_bytes<RING, NT, APT>(p.n_slots, p.lut2_n);
    // + static: 4 B-tile buffers (and the fp32 screening tiles of the two-A-point build)
    PIB_CHECK(smem + 14336 + (RING == 10 && APT == 2 ? 4 * TF32_TILE_BYTES : 0) <= (size_t)MAX_SMEM, PIB_ERR_UNSUPPORTED,
              "pair_ring6_kernel: tables do not fit in shared memory");
#define PIB_L6(M, E)                                                                                                              \
    do {                                                                                                                          \
        PIB_CUDA(cudaFuncSetAttribute(pair_ring6_kernel<RING, NT, APT, U, NL, M, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      (int)smem));                                                                                \
        pair_ring6_kernel<RING, NT, APT, U, NL, M, E><<<grid, NT, smem, stream>>>(p);                                             \
    } while (0)
    if constexpr (ALL) {
        if (multi_fix) { if (ellipse) PIB_L6(true, true); else PIB_L6(true, false); }
        else { if (ellipse) PIB_L6(false, true); else PIB_L6(false, false); }
    } else {
        PIB_CHECK(!multi_fix && !ellipse, PIB_ERR_UNSUPPORTED, "pair_ring6_kernel: experiment variant, plain omnidirectional only");
        PIB_L6(false, false);
    }
#undef PIB_L6
    PIB_CUDA(cudaGetLastError());
    return PIB_OK;
}

static int launch_pair_ring6(const PairParams &p, int ring, int nt, int apt, int u, int nl, bool multi_fix, bool ellipse, int grid,
                             cudaStream_t stream)
{
#ifdef PIB_V6_VARIANTS
#define PIB_V6_TRY(R, T, A, UU) \
    if (ring == R && nt == T && apt == A && u == UU && nl == 2) return launch_pair_ring6_t<R, T, A, UU, 2, false>(p, multi_fix, ellipse, grid, stream)
    PIB_V6_TRY(16, 512, 1, 8);
    PIB_V6_TRY(16, 512, 1, 16);
    PIB_V6_TRY(16, 256, 2, 8);
    PIB_V6_TRY(12, 384, 2, 8);
    PIB_V6_TRY(12, 640, 1, 8);
    PIB_V6_TRY(10, 768, 1, 8);
    PIB_V6_TRY(9, 512, 2, 8);
    PIB_V6_TRY(10, 512, 2, 8);
    PIB_V6_TRY(11, 512, 2, 8);
#undef PIB_V6_TRY
#else
    if (ring == 10 && nt == 512 && apt == 2 && u == 8 && nl == 2) return launch_pair_ring6_t<10, 512, 2, 8, 2, true>(p, multi_fix, ellipse, grid, stream);
    if (ring == 16 && nt == 512 && apt == 1 && u == 8 && nl == 2) return launch_pair_ring6_t<16, 512, 1, 8, 2, true>(p, multi_fix, ellipse, grid, stream);
    if (ring == 32 && nt == 256 && apt == 1 && u == 8 && nl == 2) return launch_pair_ring6_t<32, 256, 1, 8, 2, true>(p, multi_fix, ellipse, grid, stream);
    if (ring == 32 && nt == 256 && apt == 1 && u == 8 && nl == 4) return launch_pair_ring6_t<32, 256, 1, 8, 4, true>(p, multi_fix, ellipse, grid, stream);
    if (ring == 16 && nt == 512 && apt == 1 && u == 8 && nl == 4) return launch_pair_ring6_t<16, 512, 1, 8, 4, true>(p, multi_fix, ellipse, grid, stream);
#endif
    set_error("pair_ring6_kernel: variant not built");
    return PIB_ERR_UNSUPPORTED;
}

int variogram_device(const double *d_xyz, int64_t n, const double *edges, const double *lower, int n_lags,
                     const double *W, int n_dirs, int shard_index, int shard_count,
                     double *d_sum_sq, double *d_sum_prod, double *d_sum_val, int64_t *d_count,
                     int64_t *pairs_evaluated, cudaStream_t stream)
{
    PIB_CHECK(n >= 0 && n_lags >= 1 && edges != nullptr, PIB_ERR_INVALID, "variogram: need n >= 0 and at least one lag");
    PIB_CHECK(n_dirs >= 0 && (n_dirs == 0 || W != nullptr), PIB_ERR_INVALID, "variogram: n_dirs > 0 needs W");
    PIB_CHECK(shard_count >= 1 && shard_index >= 0 && shard_index < shard_count, PIB_ERR_INVALID, "variogram: bad shard");
    PIB_CHECK(edges[0] > 0, PIB_ERR_INVALID, "variogram: lags must be positive");
    for (int b = 1; b < n_lags; ++b)
        PIB_CHECK(edges[b] > edges[b - 1], PIB_ERR_INVALID, "variogram: lags must be strictly ascending");
    PIB_CHECK(std::isfinite(edges[n_lags - 1]), PIB_ERR_INVALID, "variogram: lags must be finite");

    const int out_dirs = n_dirs > 0 ? n_dirs : 1;
    const size_t out_n = (size_t)out_dirs * n_lags;
    PIB_CUDA(cudaMemsetAsync(d_sum_sq, 0, out_n * sizeof(double), stream));
    PIB_CUDA(cudaMemsetAsync(d_sum_prod, 0, out_n * sizeof(double), stream));
    PIB_CUDA(cudaMemsetAsync(d_sum_val, 0, out_n * sizeof(double), stream));
    PIB_CUDA(cudaMemsetAsync(d_count, 0, out_n * sizeof(int64_t), stream));
    if (pairs_evaluated) *pairs_evaluated = 0;
    timing_reset(PIB_T_PAIR);
    if (n < 2) return PIB_OK;

    Scratch ws(stream);
    SortedPoints sp;
    PIB_TRY(build_sorted_points(d_xyz, n, TILE, TILE, ws, stream, &sp));

    // thresholds on the squared distance
    const int n_slots = n_lags + 2;
    std::vector<double> up(n_slots), up_lag(n_lags), low_lag(n_lags);
    bool regular = true;
    for (int b = 0; b < n_lags; ++b) {
        up_lag[b] = sq_threshold(edges[b]);
        const double lo = (n_dirs > 0 && lower) ? lower[b] : (b ? edges[b - 1] : 0.0);
        low_lag[b] = sq_threshold(lo);
        if (lo != (b ? edges[b - 1] : 0.0)) regular = false;
    }
    up[0] = 0.0;
    for (int b = 0; b < n_lags; ++b) up[b + 1] = up_lag[b];
    up[n_slots - 1] = INFINITY;

    unsigned long long *d_evaluated;
    PIB_TRY(ws.alloc(&d_evaluated, 1));
    PIB_CUDA(cudaMemsetAsync(d_evaluated, 0, sizeof(unsigned long long), stream));

    int nt = 0;
    const size_t dyn_max = (size_t)MAX_SMEM - STATIC_SMEM - 256;
    if (n_slots <= MAX_SLOTS) {
        if (pair_smem_bytes<128>(n_slots) <= dyn_max) nt = 128;
        else if (pair_smem_bytes<64>(n_slots) <= dyn_max) nt = 64;
        else if (pair_smem_bytes<32>(n_slots) <= dyn_max) nt = 32;
    }
    const char *force = std::getenv("PIB_FORCE_GENERIC");
    const bool generic = !regular || nt == 0 || (force && force[0] == '1');

    if (generic) {
        set_kernel_name(PIB_T_PAIR, "pair_generic_kernel");
        double *d_up, *d_low;
        PIB_TRY(ws.alloc(&d_up, n_lags)); PIB_TRY(ws.alloc(&d_low, n_lags));
        PIB_CUDA(cudaMemcpyAsync(d_up, up_lag.data(), n_lags * sizeof(double), cudaMemcpyHostToDevice, stream));
        PIB_CUDA(cudaMemcpyAsync(d_low, low_lag.data(), n_lags * sizeof(double), cudaMemcpyHostToDevice, stream));
        const int grid = (int)std::min<int64_t>((n + 127) / 128, 8 * sm_count());
        for (int t = 0; t < out_dirs; ++t) {
            const double *w = n_dirs ? W + 4 * t : nullptr;
            pair_generic_kernel<<<grid, 128, 0, stream>>>(
                sp.xy, sp.z, n, d_up, d_low, n_lags, n_dirs > 0, w ? w[0] : 0, w ? w[1] : 0, w ? w[2] : 0, w ? w[3] : 0,
                shard_index, shard_count, d_sum_sq + (size_t)t * n_lags, d_sum_prod + (size_t)t * n_lags,
                d_sum_val + (size_t)t * n_lags, reinterpret_cast<unsigned long long *>(d_count) + (size_t)t * n_lags,
                d_evaluated);
            PIB_CUDA(cudaGetLastError());
        }
    } else {
        double *d_upper;
        PIB_TRY(ws.alloc(&d_upper, n_slots));
        PIB_CUDA(cudaMemcpyAsync(d_upper, up.data(), n_slots * sizeof(double), cudaMemcpyHostToDevice, stream));
        // slot table indexed by (high word of d2 - base) >> 13, 128 cells per octave; fix_steps = most thresholds any
        // cell straddles (the kernels make that many exact threshold steps after the lookup)
        std::vector<uint16_t> lut2;
        int lut2_base = 0, fix_steps = 0;
        uint16_t *d_lut2 = nullptr;
        {
            auto hiword = [](double v) { long long b; std::memcpy(&b, &v, 8); return (int)(b >> 32); };
            auto from_hi = [](int hi) { long long b = (long long)hi << 32; double v; std::memcpy(&v, &b, 8); return v; };
            const int cell = 1 << LUT2_SHIFT;
            const int hi_max = hiword(up[n_slots - 2]);
            lut2_base = hiword(up[1] / 4.0) & ~(cell - 1);
            int n_cells = ((hi_max - lut2_base) >> LUT2_SHIFT) + 3;
            if (n_cells > LUT2_MAX) {                          // very wide lag range: coarser start, more fix-up steps
                n_cells = LUT2_MAX;
                lut2_base = (hi_max & ~(cell - 1)) - (LUT2_MAX - 3) * cell;
            }
            lut2.resize(n_cells);
            auto slot_of = [&](double v) { int q = 0; while (q < n_slots - 1 && v > up[q]) ++q; return q; };
            for (int e = 0; e < n_cells; ++e) {
                const double v_lo = e == 0 ? 0.0 : from_hi(lut2_base + e * cell);
                const int q_lo = slot_of(v_lo);
                const int q_hi = e == n_cells - 1 ? n_slots - 1
                                                  : slot_of(std::nextafter(from_hi(lut2_base + (e + 1) * cell), 0.0));
                lut2[e] = (uint16_t)q_lo;
                fix_steps = std::max(fix_steps, q_hi - q_lo);
            }
        }
        int fix2 = fix_steps;                      // (the table is uploaded further down: the fp32 screening path lowers some cells)

        const int n_tiles = (int)(sp.n_pad / TILE);
        int chunk = CHUNK;                         // B tiles per work item; pair_ring6_kernel takes fewer for many ranks (below)
        int n_chunks = (n_tiles + CHUNK - 1) / CHUNK;

        // ring-window kernel or full-histogram kernel?  The ring needs a tile pair to span < 32 lags:
        // estimate the span from the tile bounding boxes (A sub-tile = 2 tiles, B tile = 1 tile).
        bool use_ring = false, ring16_ok = false;
        const char *force_full = std::getenv("PIB_FORCE_FULL");
        double stretch = 1.0;                      // largest singular value of the whitening matrices
        for (int t = 0; t < n_dirs; ++t) {
            const double *w = W + 4 * t;
            const double frob = w[0] * w[0] + w[1] * w[1] + w[2] * w[2] + w[3] * w[3], det = w[0] * w[3] - w[1] * w[2];
            stretch = std::max(stretch, std::sqrt(0.5 * (frob + std::sqrt(std::max(0.0, frob * frob - 4.0 * det * det)))));
        }
        double d90_scaled = INFINITY, min_width = edges[0];
        for (int b = 1; b < n_lags; ++b) min_width = std::min(min_width, edges[b] - edges[b - 1]);
        if (n_lags <= MAX_SLOTS - 2 && !(force_full && force_full[0] == '1')) {
            // 90th percentile of the tile diagonals decides which kernel fits (a performance choice only: every kernel
            // is exact for any geometry).  It needs the tile boxes on the host — a blocking read-back in the middle of
            // the pre-pass — so the answer is remembered per problem shape (point count, extent, lag layout, stretch): repeated
            // calls on the same kind of data (benchmark steps, the directions / thresholds of one data set, the ranks of a
            // sharded run) skip the round trip.
            PIB_TRY(sp.wait_box());        // (recorded at the head of the stream: it has arrived by now)
            const double box_w = sp.box[2] - sp.box[0], box_h = sp.box[3] - sp.box[1];
            struct Shape { int64_t n; int n_lags; double e0, e1, stretch, w, h, d90; };
            static std::mutex cache_mutex;
            static std::vector<Shape> cache;
            double d90 = -1.0;
            {
                std::lock_guard<std::mutex> lock(cache_mutex);
                for (const Shape &c : cache)
                    if (c.n == n && c.n_lags == n_lags && c.e0 == edges[0] && c.e1 == edges[n_lags - 1] && c.stretch == stretch &&
                        std::fabs(c.w - box_w) <= 1e-3 * box_w && std::fabs(c.h - box_h) <= 1e-3 * box_h) d90 = c.d90;
            }
            const char *nocache = std::getenv("PIB_NO_SHAPE_CACHE");
            if (d90 < 0.0 || (nocache && nocache[0] == '1')) {
                std::vector<double4> boxes(n_tiles);
                PIB_CUDA(cudaMemcpyAsync(boxes.data(), sp.tile_box, sizeof(double4) * n_tiles, cudaMemcpyDeviceToHost, stream));
                PIB_CUDA(cudaStreamSynchronize(stream));
                std::vector<double> diag;
                diag.reserve(n_tiles);
                for (auto &b : boxes)
                    if (b.z >= b.x) diag.push_back(std::hypot(b.z - b.x, b.w - b.y));
                if (!diag.empty()) {
                    std::sort(diag.begin(), diag.end());
                    d90 = stretch * diag[(size_t)(0.9 * (diag.size() - 1))];
                    std::lock_guard<std::mutex> lock(cache_mutex);
                    if (cache.size() >= 64) cache.erase(cache.begin());
                    cache.push_back(Shape{n, n_lags, edges[0], edges[n_lags - 1], stretch, box_w, box_h, d90});
                }
            }
            if (d90 >= 0.0) {
                d90_scaled = d90;
                use_ring = (2.5 * d90 / min_width + 2.0) <= 24.0;
                ring16_ok = (3.0 * d90 / min_width + 2.0) <= 16.0;       // A sub-tile = 4 tiles (~2x the diagonal)
            }
            const char *force_ring = std::getenv("PIB_FORCE_RING");
            if (force_ring && force_ring[0] == '1') use_ring = true;
            if (force_ring && force_ring[0] == '1' && force_ring[1] == '6') ring16_ok = true;    // "16"
            const char *ring_env = std::getenv("PIB_RING");
            if (ring_env) ring16_ok = (ring_env[0] == '1' && ring_env[1] == '6');
        }
        // old formulation (pair_ring_kernel): 32-slot ring / 256 threads, or 16-slot ring / 512 threads (more warps to
        // hide latency) when tile pairs are narrow enough; CTA-wide ring window
        int ring = 32, ring_nt = 256;
        if (use_ring && ring16_ok) { ring = 16; ring_nt = 512; }
        // new formulation (pair_ring6_kernel, the default): per-warp ring windows, so the window only has to cover the
        // lags between ONE warp's 32 * APT points and a B tile
        const char *ver_env = std::getenv("PIB_PAIR_V");
        bool v6 = !(ver_env && ver_env[0] == '5');
        int v6_apt = 1, v6_u = 8, v6_nl = 2;
        if (v6 && !(force_full && force_full[0] == '1') && n_lags <= MAX_SLOTS - 2) {
            const double span1 = 1.5 * d90_scaled / min_width + 2.0;     // A points per warp = 32: box ~ half a tile's diagonal
            if (span1 <= 16.0) { ring = 16; ring_nt = 512; use_ring = true; }
            else if (span1 <= 40.0) { ring = 32; ring_nt = 256; use_ring = true; }      // (tile pairs wider than the ring are swept per window)
            else use_ring = false;
            const char *force_ring = std::getenv("PIB_FORCE_RING");
            if (force_ring && force_ring[0] == '1') {
                use_ring = true;
                if (force_ring[1] == '6') { ring = 16; ring_nt = 512; }
            }
            const char *ring_env = std::getenv("PIB_RING");
            if (ring_env && use_ring) {
                if (ring_env[0] == '1' && ring_env[1] == '6') { ring = 16; ring_nt = 512; } else { ring = 32; ring_nt = 256; }
            }
            // lags per group: a group of 8 consecutive points spans about a quarter of its tile's diagonal; when that is more
            // than ~1.3 lag widths (sparse points, fine lags, stretched ellipses) most groups straddle three lags or more
            // and the four-lag scheme (three cumulative masks per pair, four ring updates per group) pays
            v6_nl = (0.25 * d90_scaled <= 1.3 * min_width) ? 2 : 4;
            const char *nl_env = std::getenv("PIB_LAGS_PER_GROUP");
            if (nl_env && (nl_env[0] == '2' || nl_env[0] == '4')) v6_nl = nl_env[0] - '0';
            // two A points per thread (every B point read from shared memory serves two pairs): 64 A points per warp, so the
            // ring shrinks to 10 slots to keep 1024 A points per CTA — taken when a warp's box against a B tile stays inside
            // 10 lags (BASELINE config 2: 9 at the 99th percentile); the two-lag scheme only
            const double span2 = 1.75 * d90_scaled / min_width + 2.0;    // 64 consecutive points: ~0.75 of a tile's diagonal
            const char *apt_env = std::getenv("PIB_APT");
            const bool apt2_fits = ring6_smem_bytes<10, 512, 2>(n_slots, (int)lut2.size()) + 14336 + 4 * TF32_TILE_BYTES <= (size_t)MAX_SMEM;
            if (use_ring && v6_nl == 2 && !ring_env && !force_ring && span2 <= 10.0 && !(apt_env && apt_env[0] == '1') && apt2_fits) {
                v6_apt = 2; ring = 10; ring_nt = 512;
            }
            // PIB_V6 = "apt,u,threads,ring" picks another build of the kernel for A/B measurements, e.g. "2,8,384,12"
            const char *v6_env = std::getenv("PIB_V6");
            if (v6_env && use_ring) {
                int e_apt = 1, e_u = 8, e_nt = ring_nt, e_ring = ring;
                if (std::sscanf(v6_env, "%d,%d,%d,%d", &e_apt, &e_u, &e_nt, &e_ring) >= 2) {
                    v6_apt = e_apt; v6_u = e_u; ring_nt = e_nt; ring = e_ring;
                }
            }
        } else v6 = false;
        if (!use_ring) v6 = false;
        if (v6) {
            // B tiles per work item: CHUNK.  Per-rank kernel times of config 2 at 8 ranks (scripts/exp_shard_balance.py, ideal
            // 17.91 ms) with the snake-order deal of pair_ring6_kernel: 64-tile items 18.20 ms, 32-tile items 18.41 ms (the
            // per-item set-up outweighs the finer balance).  PIB_PAIR_CHUNK overrides.
            const char *chunk_env = std::getenv("PIB_PAIR_CHUNK");
            if (chunk_env && std::atoi(chunk_env) >= 1 && std::atoi(chunk_env) <= CHUNK) chunk = std::atoi(chunk_env);
            n_chunks = (n_tiles + chunk - 1) / chunk;
        }
        // fp32 screening path of the two-A-point kernel (pair_ring6.cuh, sweep_f32): per-slot bands from the error bound
        // u (6 T + 5.66 R sqrt(T)) x 1.05, u = 2^-24, R = the widest tile half-extent the path accepts; the slot table is lowered
        // to the first slot a cell's smallest value may still belong to (a valid lower bound for the exact stages as well)
        std::vector<float2> slotf;
        double f32_rcap = 0.0;
        {
            const char *f32_env = std::getenv("PIB_F32");
            const bool want = v6 && v6_apt == 2 && ring == 10 && ring_nt == 512 && v6_u == 8 && v6_nl == 2 && n_dirs == 0 &&
                              std::isfinite(d90_scaled) && d90_scaled > 0.0 && !(f32_env && f32_env[0] == '0') &&
                              up[1] >= 1e-30 && up[n_slots - 2] <= 1e30;
            if (want) {
                f32_rcap = d90_scaled;
                const double u24 = std::ldexp(1.0, -24);
                slotf.resize(n_slots + 4);
                for (int sl = 0; sl < n_slots + 4; ++sl) {
                    if (sl == 0) { slotf[sl] = make_float2(0.0f, 0.0f); continue; }
                    if (sl >= n_slots - 1) { slotf[sl] = make_float2(FLT_MAX, 0.0f); continue; }
                    const double T = up[sl];
                    const double band = 1.05 * u24 * (6.0 * T + 5.66 * f32_rcap * std::sqrt(T)) + 1e-37;
                    const float tmid = (float)T;
                    float hw = (float)(band + std::fabs(T - (double)tmid));
                    while ((double)hw < band + std::fabs(T - (double)tmid)) hw = std::nextafterf(hw, INFINITY);
                    hw = std::nextafterf(hw, INFINITY);
                    slotf[sl] = make_float2(tmid, hw);
                }
                auto slot_f = [&](double v) { int q = 0; while (q < n_slots - 1 && v - (double)slotf[q].x > (double)slotf[q].y) ++q; return q; };
                auto slot_x = [&](double v) { int q = 0; while (q < n_slots - 1 && v > up[q]) ++q; return q; };
                auto from_hi = [](int hi) { long long b = (long long)hi << 32; double v; std::memcpy(&v, &b, 8); return v; };
                const int cell = 1 << LUT2_SHIFT, n_cells = (int)lut2.size();
                int steps = 0;
                for (int e = 0; e < n_cells; ++e) {
                    const double v_lo = e == 0 ? 0.0 : from_hi(lut2_base + e * cell);
                    const int q_lo = std::min((int)lut2[e], slot_f(v_lo));
                    lut2[e] = (uint16_t)q_lo;
                    if (e < n_cells - 1) {
                        const double v_hi = from_hi(lut2_base + (e + 1) * cell);
                        steps = std::max(steps, std::max(slot_x(std::nextafter(v_hi, 0.0)), slot_f((double)std::nextafterf((float)v_hi, 0.0f))) - q_lo);
                    } else steps = std::max(steps, n_slots - 1 - q_lo);
                }
                fix2 = std::max(fix2, steps);
            }
        }
        PIB_TRY(ws.alloc(&d_lut2, lut2.size()));
        PIB_CUDA(cudaMemcpyAsync(d_lut2, lut2.data(), lut2.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, stream));
        const int a_per_cta = use_ring ? ring_nt * (v6 ? v6_apt : 1) : nt;
        if (use_ring) nt = ring_nt;
        const int64_t a_tiles = use_ring ? (n_tiles + a_per_cta / TILE - 1) / (a_per_cta / TILE) : (int64_t)n_tiles * (TILE / nt);
        const int64_t n_items = (a_tiles * n_chunks + shard_count - 1) / shard_count;
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(sm_count(), n_items));
        const int warps = nt / 32;
        double *d_partial;
        const size_t part_n = ((size_t)grid * warps + 1) * n_slots * 4;
        PIB_TRY(ws.alloc(&d_partial, part_n));

        PairParams p;
        p.xy = sp.xy; p.z = sp.z; p.tile_box = sp.tile_box; p.n = n; p.n_tiles = n_tiles; p.n_slots = n_slots;
        p.upper = d_upper; p.max_d2 = up_lag[n_lags - 1];
        p.hiw_ok = 1;
        for (int sl = 0; sl < n_slots; ++sl) {
            uint64_t bits;
            std::memcpy(&bits, &up[sl], sizeof(bits));
            if ((uint32_t)bits == 0xffffffffu) p.hiw_ok = 0;
        }
        { const char *hiw_env = std::getenv("PIB_HIW"); if (hiw_env && hiw_env[0] == '0') p.hiw_ok = 0; }
        p.shard_index = shard_index; p.shard_count = shard_count; p.partial = d_partial;
        p.pairs_evaluated = d_evaluated;
        p.lut2 = d_lut2; p.lut2_base = lut2_base; p.lut2_n = (int)lut2.size();
        p.lut2_base_f = (int)((uint32_t)(lut2_base - 0x38000000) << 3);
        p.box_slack = 0.0;
        p.gsum = nullptr;
        double *d_mean = nullptr;                  // ring kernel: mean of z (its sums come out centred)
        int64_t guard_groups = 0;                  // ring kernel: number of group sums (cancellation guard)
        p.item_list = nullptr; p.n_live_items = nullptr; p.run_flag = nullptr; p.chunk = chunk;
        p.tf32 = nullptr; p.slotf = nullptr; p.f32_rcap = f32_rcap;
        if (!slotf.empty()) {
            unsigned char *d_tf32;
            float2 *d_slotf;
            PIB_TRY(ws.alloc(&d_tf32, (size_t)n_tiles * TF32_TILE_BYTES));
            PIB_TRY(ws.alloc(&d_slotf, slotf.size()));
            PIB_CUDA(cudaMemcpyAsync(d_slotf, slotf.data(), slotf.size() * sizeof(float2), cudaMemcpyHostToDevice, stream));
            tile_f32_kernel<<<(int)(((int64_t)n_tiles * (TILE / 2) + 255) / 256), 256, 0, stream>>>(sp.xy, sp.tile_box, n_tiles, d_tf32);
            PIB_CUDA(cudaGetLastError());
            p.tf32 = d_tf32; p.slotf = d_slotf;
        }
        const int64_t n_items_all = a_tiles * n_chunks;            // over all ranks
        const char *bal_env = std::getenv("PIB_BALANCE");
        const bool balance = use_ring && n_items_all < ((int64_t)1 << 31) && !(bal_env && bal_env[0] == '0');
        uint32_t *d_keys = nullptr, *d_keys2 = nullptr;
        int32_t *d_items = nullptr, *d_items2 = nullptr;
        unsigned *d_nlive = nullptr;
        if (balance) {
            PIB_TRY(ws.alloc(&d_keys, (size_t)n_items_all)); PIB_TRY(ws.alloc(&d_keys2, (size_t)n_items_all));
            PIB_TRY(ws.alloc(&d_items, (size_t)n_items_all)); PIB_TRY(ws.alloc(&d_items2, (size_t)n_items_all));
            PIB_TRY(ws.alloc(&d_nlive, 1));
        }
        if (use_ring) {
            const int gsz = v6 ? v6_u : 8;                     // points per group (the group sums ride along with every B tile)
            const int64_t n_groups = (int64_t)n_tiles * (TILE / gsz);
            double2 *d_gsum;
            double *d_zc;
            PIB_TRY(ws.alloc(&d_gsum, (size_t)n_groups));
            PIB_TRY(ws.alloc(&d_zc, (size_t)n_tiles * TILE + 1024));
            PIB_TRY(ws.alloc(&d_mean, 1));
            PIB_CUDA(cudaMemsetAsync(d_zc, 0, ((size_t)n_tiles * TILE + 1024) * sizeof(double), stream));
            double *d_mpart;
            PIB_TRY(ws.alloc(&d_mpart, 256));
            mean_partial_kernel<<<256, 256, 0, stream>>>(sp.z, n, d_mpart);
            mean_final_kernel<<<1, 256, 0, stream>>>(d_mpart, 256, n, d_mean);
            if (gsz == 16)
                group_sum_kernel<16><<<(int)((n_groups + 255) / 256), 256, 0, stream>>>(sp.z, d_mean, n, n_groups, d_zc, d_gsum);
            else
                group_sum_kernel<8><<<(int)((n_groups + 255) / 256), 256, 0, stream>>>(sp.z, d_mean, n, n_groups, d_zc, d_gsum);
            p.z = d_zc;
            PIB_CUDA(cudaGetLastError());
            p.gsum = d_gsum;
            guard_groups = n_groups;
        }
        {
            // the two-lag group scheme stays on even when groups of 8 points are wider than a lag (sparse points, stretched
            // ellipses): the pairs beyond the second lag take their own ring update, the rest still share two (measured on
            // BASELINE config 3: 70 ms against 86 ms with per-pair updates only)
            const char *gp = std::getenv("PIB_GROUP_PATH");
            p.group_path = (gp && gp[0] == '0') ? 0 : 1;
        }
        double4 *d_boxw = nullptr;                 // ellipse: bounding boxes of the whitened coordinates, per direction
        int *d_guard_flag = nullptr;               // cancellation guard: flag, fallback partials and a scratch pair counter
        double *d_partial_fb = nullptr;
        size_t fb_part_n = 0;
        int fb_grid = 0;
        unsigned long long *d_evaluated_fb = nullptr;
        PIB_TRY(ws.alloc(&d_evaluated_fb, 1));
        if (n_dirs > 0) PIB_TRY(ws.alloc(&d_boxw, n_tiles));
        for (int t = 0; t < out_dirs; ++t) {
            PIB_CUDA(cudaMemsetAsync(d_partial, 0, part_n * sizeof(double), stream));
            int rc;
            if (n_dirs > 0) {
                p.w00 = W[4 * t]; p.w01 = W[4 * t + 1]; p.w10 = W[4 * t + 2]; p.w11 = W[4 * t + 3];
                tile_box_whitened_kernel<<<(int)(((int64_t)n_tiles * 32 + 255) / 256), 256, 0, stream>>>(
                    sp.xy, n, n_tiles, p.w00, p.w01, p.w10, p.w11, d_boxw);
                PIB_CUDA(cudaGetLastError());
                p.tile_box = d_boxw;
                // |W (p_j - p_i) as computed  -  (W p_j - W p_i) as boxed| <= a few ulps of |W| |p|; 1e-12 relative is ample
                PIB_TRY(sp.wait_box());
                const double pmax = std::fabs(sp.box[0]) + std::fabs(sp.box[1]) + std::fabs(sp.box[2]) + std::fabs(sp.box[3]);
                p.box_slack = 1e-12 * pmax * (std::fabs(p.w00) + std::fabs(p.w01) + std::fabs(p.w10) + std::fabs(p.w11));
            } else p.w00 = p.w01 = p.w10 = p.w11 = 0;
            if (balance) {
                PIB_CUDA(cudaMemsetAsync(d_nlive, 0, sizeof(unsigned), stream));
                const int cgrid = (int)((n_items_all * 32 + 255) / 256);
                if (n_dirs > 0)
                    item_cost_kernel<true><<<cgrid, 256, 0, stream>>>(p, a_per_cta / TILE, n_items_all, d_keys, d_items, d_nlive);
                else
                    item_cost_kernel<false><<<cgrid, 256, 0, stream>>>(p, a_per_cta / TILE, n_items_all, d_keys, d_items, d_nlive);
                PIB_CUDA(cudaGetLastError());
                PIB_TRY(sort_pairs_u32(d_keys, d_keys2, d_items, d_items2, n_items_all, ITEM_COST_BITS, ws, stream));
                p.item_list = d_items2; p.n_live_items = d_nlive;
            }
            {
                char name[128];
                if (use_ring && v6) std::snprintf(name, sizeof(name), "pair_ring6_kernel<ring %d, %d threads, %d A/thread, groups of %d over %d lags%s%s>", ring, ring_nt, v6_apt, v6_u, v6_nl, n_dirs > 0 ? ", ellipse" : "", p.tf32 ? ", fp32 screening" : "");
                else if (use_ring) std::snprintf(name, sizeof(name), "pair_ring_kernel<ring %d, %d threads%s>", ring, ring_nt, n_dirs > 0 ? ", ellipse" : "");
                else std::snprintf(name, sizeof(name), "pair_kernel<%d threads%s>", nt, n_dirs > 0 ? ", ellipse" : "");
                set_kernel_name(PIB_T_PAIR, name);
            }
            timing_begin(PIB_T_PAIR, stream);
            if (use_ring && v6) {
                rc = launch_pair_ring6(p, ring, ring_nt, v6_apt, v6_u, v6_nl, fix2 > 1, n_dirs > 0, grid, stream);
            } else if (use_ring) {
                const size_t smem = (size_t)(ring + 2) * ring_nt * (sizeof(double2) + sizeof(uint32_t));
#define PIB_LAUNCH_RING(R, T, F, E)                                                                                     \
    do {                                                                                                                \
        PIB_CUDA(cudaFuncSetAttribute(pair_ring_kernel<R, T, F, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        pair_ring_kernel<R, T, F, E><<<grid, T, smem, stream>>>(p);                                                    \
    } while (0)
#define PIB_LAUNCH_RING_E(R, T, F)                        \
    do {                                                  \
        if (n_dirs > 0) PIB_LAUNCH_RING(R, T, F, true);   \
        else PIB_LAUNCH_RING(R, T, F, false);             \
    } while (0)
                if (ring == 16) {
                    if (fix2 <= 1) PIB_LAUNCH_RING_E(16, 512, 1);
                    else if (fix2 <= 2) PIB_LAUNCH_RING_E(16, 512, 2);
                    else PIB_LAUNCH_RING_E(16, 512, 0);
                } else {
                    if (fix2 <= 1) PIB_LAUNCH_RING_E(32, 256, 1);
                    else if (fix2 <= 2) PIB_LAUNCH_RING_E(32, 256, 2);
                    else PIB_LAUNCH_RING_E(32, 256, 0);
                }
#undef PIB_LAUNCH_RING_E
#undef PIB_LAUNCH_RING
                rc = cudaGetLastError() == cudaSuccess ? PIB_OK : PIB_ERR_CUDA;
                if (rc != PIB_OK) set_error("pair_ring_kernel launch failed");
            } else if (n_dirs > 0) {
                rc = nt == 128 ? launch_pair_kernel<128, true>(p, fix_steps, grid, stream)
                   : nt == 64 ? launch_pair_kernel<64, true>(p, fix_steps, grid, stream)
                              : launch_pair_kernel<32, true>(p, fix_steps, grid, stream);
            } else {
                p.w00 = p.w01 = p.w10 = p.w11 = 0;
                rc = nt == 128 ? launch_pair_kernel<128, false>(p, fix_steps, grid, stream)
                   : nt == 64 ? launch_pair_kernel<64, false>(p, fix_steps, grid, stream)
                              : launch_pair_kernel<32, false>(p, fix_steps, grid, stream);
            }
            timing_end(PIB_T_PAIR, stream);
            PIB_TRY(rc);
            reduce_partials_kernel<<<n_lags, 128, 0, stream>>>(
                d_partial, grid * warps + 1, n_slots, n_lags, d_mean, d_sum_sq + (size_t)t * n_lags, d_sum_prod + (size_t)t * n_lags,
                d_sum_val + (size_t)t * n_lags, d_count + (size_t)t * n_lags);
            PIB_CUDA(cudaGetLastError());
            if (use_ring) {
                // Cancellation guard (no host round trip): a flag kernel looks at the finished lags; the per-pair dz-form
                // kernel and a second reduction are queued behind it and return at once unless the flag is set.
                // (A sharded call guards its own slice of every lag: the bound holds for each rank's partial sums.)
                int nt_fb = 0;
                if (pair_smem_bytes<128>(n_slots) <= dyn_max) nt_fb = 128;
                else if (pair_smem_bytes<64>(n_slots) <= dyn_max) nt_fb = 64;
                else if (pair_smem_bytes<32>(n_slots) <= dyn_max) nt_fb = 32;
                if (nt_fb) {
                    if (!d_guard_flag) {
                        PIB_TRY(ws.alloc(&d_guard_flag, 1));
                        const int64_t fb_items = ((int64_t)n_tiles * (TILE / nt_fb)) * n_chunks;
                        fb_grid = (int)std::max<int64_t>(1, std::min<int64_t>(sm_count(), fb_items));
                        fb_part_n = ((size_t)fb_grid * (nt_fb / 32) + 1) * n_slots * 4;
                        PIB_TRY(ws.alloc(&d_partial_fb, fb_part_n));
                    }
                    PIB_CUDA(cudaMemsetAsync(d_guard_flag, 0, sizeof(int), stream));
                    PIB_CUDA(cudaMemsetAsync(d_partial_fb, 0, fb_part_n * sizeof(double), stream));
                    cancellation_guard_kernel<<<1, 1024, 0, stream>>>(p.gsum, guard_groups, n, d_sum_sq + (size_t)t * n_lags,
                                                                      d_count + (size_t)t * n_lags, n_lags, d_guard_flag);
                    PIB_CUDA(cudaGetLastError());
                    PairParams pf = p;
                    pf.z = sp.z; pf.gsum = nullptr; pf.item_list = nullptr; pf.n_live_items = nullptr;
                    pf.partial = d_partial_fb; pf.run_flag = d_guard_flag; pf.pairs_evaluated = d_evaluated_fb;
                    if (n_dirs > 0)
                        rc = nt_fb == 128 ? launch_pair_kernel<128, true>(pf, fix_steps, fb_grid, stream)
                           : nt_fb == 64 ? launch_pair_kernel<64, true>(pf, fix_steps, fb_grid, stream)
                                         : launch_pair_kernel<32, true>(pf, fix_steps, fb_grid, stream);
                    else
                        rc = nt_fb == 128 ? launch_pair_kernel<128, false>(pf, fix_steps, fb_grid, stream)
                           : nt_fb == 64 ? launch_pair_kernel<64, false>(pf, fix_steps, fb_grid, stream)
                                         : launch_pair_kernel<32, false>(pf, fix_steps, fb_grid, stream);
                    PIB_TRY(rc);
                    reduce_partials_kernel<<<n_lags, 128, 0, stream>>>(
                        d_partial_fb, fb_grid * (nt_fb / 32) + 1, n_slots, n_lags, nullptr, d_sum_sq + (size_t)t * n_lags,
                        d_sum_prod + (size_t)t * n_lags, d_sum_val + (size_t)t * n_lags, d_count + (size_t)t * n_lags, d_guard_flag);
                    PIB_CUDA(cudaGetLastError());
                }
            }
        }
    }
    if (pairs_evaluated) {
        unsigned long long h = 0;
        PIB_CUDA(cudaMemcpyAsync(&h, d_evaluated, sizeof(h), cudaMemcpyDeviceToHost, stream));
        PIB_CUDA(cudaStreamSynchronize(stream));
        *pairs_evaluated = (int64_t)h;
    }
    return sp.wait_box();              // non-finite coordinates are reported here (the copy arrived long ago: no stall)
}


// ---------------------------------------------------------------------------------------------
// Variogram point cloud: per lag, the squared differences (z_i - z_j)^2 of the ORDERED
// pairs in the reference's order — rows i ascending, within a row j ascending, i.e. np.where on the
// N x N mask (semivariogram/experimental/functions/general.py:13-75, :150-233; ellipse:
// functions/directional.py:83-134, :468-530).  Output is O(pairs), so this is for small N; it works on the
// points in the caller's order with one thread per row i (no tiling, no culling).
//   pass 1: counts[b][i] = ordered pairs of row i in lag b;  exclusive scan along i per lag
//   pass 2: values[lag_offset[b] + scan[b][i] + k] = (z_i - z_j)^2 for the k-th j of row i in lag b
// ---------------------------------------------------------------------------------------------
template <bool WRITE>
__global__ void cloud_kernel(const double *__restrict__ xyz, int64_t n, const double *__restrict__ up_d2,
                             const double *__restrict__ low_d2, int n_lags, bool ellipse, double w00, double w01,
                             double w10, double w11, unsigned long long *__restrict__ cursor /* [n_lags][n] */,
                             double *__restrict__ values)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double xi = xyz[3 * i], yi = xyz[3 * i + 1], zi = xyz[3 * i + 2];
    for (int64_t j = 0; j < n; ++j) {
        if (j == i) continue;
        double d2;
        if (ellipse) {
            const double dx = __dsub_rn(xyz[3 * j], xi), dy = __dsub_rn(xyz[3 * j + 1], yi);
            const double n0 = __fma_rn(w01, dy, __dmul_rn(w00, dx));
            const double n1 = __fma_rn(w11, dy, __dmul_rn(w10, dx));
            d2 = __dadd_rn(__dmul_rn(n0, n0), __dmul_rn(n1, n1));
        } else {
            const double dx = __dsub_rn(xi, xyz[3 * j]), dy = __dsub_rn(yi, xyz[3 * j + 1]);
            d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        }
        if (!(d2 > 0.0)) continue;
        for (int b = 0; b < n_lags; ++b) {
            if (d2 <= up_d2[b] && d2 > low_d2[b]) {
                unsigned long long *c = cursor + (size_t)b * n + i;
                if (WRITE) {
                    const double dz = zi - xyz[3 * j + 2];
                    values[(*c)++] = dz * dz;
                } else ++(*c);
            }
        }
    }
}

// exclusive scan of each lag's row counts, shifted by the lag's global offset (one CTA per lag)
__global__ void cloud_scan_kernel(unsigned long long *cursor, int64_t n, const int64_t *__restrict__ lag_offsets,
                                  unsigned long long *__restrict__ lag_end /* NULL, or [n_lags]: offset + total of each lag */)
{
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    unsigned long long *row = cursor + (size_t)blockIdx.x * n;
    if (threadIdx.x == 0) s_carry = (unsigned long long)lag_offsets[blockIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = 0; base < n; base += blockDim.x) {
        const int64_t i = base + threadIdx.x;
        const unsigned long long v = i < n ? row[i] : 0ull;
        unsigned long long incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0ull, wi = w;
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            s_warp[lane] = wi - w;                               // exclusive prefix of the warp totals
        }
        __syncthreads();
        const unsigned long long carry = s_carry;
        if (i < n) row[i] = carry + s_warp[warp] + incl - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = carry + s_warp[warp] + incl;
        __syncthreads();
    }
    if (lag_end && threadIdx.x == 0) lag_end[blockIdx.x] = s_carry;
}

int variogram_cloud_device(const double *d_xyz, int64_t n, const double *edges, const double *lower, int n_lags,
                           const double *W, int n_dirs, const int64_t *lag_offsets, double *d_values,
                           cudaStream_t stream)
{
    PIB_CHECK(n >= 0 && n_lags >= 1 && edges && lag_offsets, PIB_ERR_INVALID, "cloud: bad arguments");
    PIB_CHECK(n_dirs == 0 || (n_dirs == 1 && W), PIB_ERR_INVALID, "cloud: one direction at a time");
    if (n < 2) return PIB_OK;
    Scratch ws(stream);
    std::vector<double> up(n_lags), low(n_lags);
    for (int b = 0; b < n_lags; ++b) {
        up[b] = sq_threshold(edges[b]);
        low[b] = sq_threshold((n_dirs > 0 && lower) ? lower[b] : (b ? edges[b - 1] : 0.0));
    }
    double *d_up, *d_low;
    int64_t *d_off;
    unsigned long long *d_cursor;
    PIB_TRY(ws.alloc(&d_up, n_lags)); PIB_TRY(ws.alloc(&d_low, n_lags)); PIB_TRY(ws.alloc(&d_off, n_lags));
    PIB_TRY(ws.alloc(&d_cursor, (size_t)n_lags * n));
    PIB_CUDA(cudaMemcpyAsync(d_up, up.data(), n_lags * sizeof(double), cudaMemcpyHostToDevice, stream));
    PIB_CUDA(cudaMemcpyAsync(d_low, low.data(), n_lags * sizeof(double), cudaMemcpyHostToDevice, stream));
    PIB_CUDA(cudaMemcpyAsync(d_off, lag_offsets, n_lags * sizeof(int64_t), cudaMemcpyHostToDevice, stream));
    PIB_CUDA(cudaMemsetAsync(d_cursor, 0, sizeof(unsigned long long) * n_lags * n, stream));
    const double w00 = n_dirs ? W[0] : 0, w01 = n_dirs ? W[1] : 0, w10 = n_dirs ? W[2] : 0, w11 = n_dirs ? W[3] : 0;
    const int grid = (int)((n + 127) / 128);
    cloud_kernel<false><<<grid, 128, 0, stream>>>(d_xyz, n, d_up, d_low, n_lags, n_dirs > 0, w00, w01, w10, w11, d_cursor, nullptr);
    PIB_CUDA(cudaGetLastError());
    cloud_scan_kernel<<<n_lags, 1024, 0, stream>>>(d_cursor, n, d_off, nullptr);
    PIB_CUDA(cudaGetLastError());
    cloud_kernel<true><<<grid, 128, 0, stream>>>(d_xyz, n, d_up, d_low, n_lags, n_dirs > 0, w00, w01, w10, w11, d_cursor, d_values);
    PIB_CUDA(cudaGetLastError());
    return PIB_OK;
}


// ---------------------------------------------------------------------------------------------
// Weighted semivariance (custom_weights):
//   semivariogram/weights/experimental/weighting.py:4-54        omnidirectional (u = z / w)
//   semivariogram/experimental/functions/directional.py:276-392 directional, always ellipse selection
// Per lag, over unordered pairs (the reference's ordered sums are twice these):
//   s1 = sum w_p q       w_p = w_i w_j / (w_i + w_j);  q = (u_i - u_j)^2 (omni) or (z_i - z_j)^2 (directional)
//   s2 = sum w_p
//   s3 = sum (f_i + f_j) f = u (omni) or w z (directional)
//   s4 = sum (w_i + w_j)
// One thread per point i over all j > i, global fp64 REDs (this path serves small areal data sets).
// ---------------------------------------------------------------------------------------------
__global__ void gather_weights_kernel(const double *__restrict__ w, const int32_t *__restrict__ orig, int64_t n,
                                      double *__restrict__ out)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = w[orig[i]];
}

// points are Hilbert sorted, so for a fixed i consecutive j fall into the same lag for long runs: a thread
// accumulates a run in registers and issues its REDs only when the lag changes
__global__ void __launch_bounds__(128) weighted_pair_kernel(const double2 *__restrict__ xy, const double *__restrict__ z,
                                                            const double *__restrict__ wts, int64_t n,
                                                            const double4 *__restrict__ tbox /* per tile; whitened for ellipse */,
                                                            int n_tiles, double box_slack,
                                                            const double *__restrict__ up_g, const double *__restrict__ low_g,
                                                            int n_lags, bool ellipse, double w00, double w01, double w10,
                                                            double w11, double *sums /* [4][n_lags] */,
                                                            unsigned long long *count, int use_smem)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    // CTA-private accumulators and thresholds in shared memory; one global RED per lag per CTA at the end
    extern __shared__ __align__(16) unsigned char wt_smem[];
    double *const g_sums = sums;
    unsigned long long *const g_count = count;
    const double *up_d2 = up_g, *low_d2 = low_g;
    if (use_smem) {
        double *sh = reinterpret_cast<double *>(wt_smem);
        for (int k = threadIdx.x; k < n_lags; k += blockDim.x) { sh[k] = up_g[k]; sh[n_lags + k] = low_g[k]; }
        for (int k = threadIdx.x; k < 5 * n_lags; k += blockDim.x) sh[2 * n_lags + k] = 0.0;
        up_d2 = sh; low_d2 = sh + n_lags;
        sums = sh + 2 * n_lags;
        count = reinterpret_cast<unsigned long long *>(sh + 6 * n_lags);
        __syncthreads();
    }
    const double max_d2 = up_d2[n_lags - 1];
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < n; base += stride) {
        const int64_t i = base + threadIdx.x;
        const bool valid = i < n;
        const int64_t ic = valid ? i : n - 1;
        const double2 pi = xy[ic];
        const double zi = z[ic], wi = wts[ic];
        const double ui = zi / wi;
        // bounding box of my warp's 32 points (whitened coordinates for ellipse runs, like the tile boxes)
        double bx0 = ellipse ? w00 * pi.x + w01 * pi.y : pi.x, by0 = ellipse ? w10 * pi.x + w11 * pi.y : pi.y;
        double bx1 = bx0, by1 = by0;
        for (int o = 16; o; o >>= 1) {
            bx0 = fmin(bx0, __shfl_xor_sync(0xffffffffu, bx0, o)); bx1 = fmax(bx1, __shfl_xor_sync(0xffffffffu, bx1, o));
            by0 = fmin(by0, __shfl_xor_sync(0xffffffffu, by0, o)); by1 = fmax(by1, __shfl_xor_sync(0xffffffffu, by1, o));
        }
        const int64_t warp_first = base + (threadIdx.x & ~31);          // smallest i of my warp
        int run = -1;
        double r1 = 0, r2 = 0, r3 = 0, r4 = 0;
        unsigned long long rc = 0;
        auto flush = [&]() {
            if (run >= 0 && rc) {
                atomicAdd(&sums[run], r1); atomicAdd(&sums[n_lags + run], r2);
                atomicAdd(&sums[2 * n_lags + run], r3); atomicAdd(&sums[3 * n_lags + run], r4);
                atomicAdd(&count[run], rc);
            }
            r1 = r2 = r3 = r4 = 0; rc = 0;
        };
        // only j > i: tiles before my warp's first point hold nothing; tiles whose box is farther than the last lag
        // from the warp's box are skipped
        for (int t0 = (int)(warp_first / TILE) & ~31; t0 < n_tiles; t0 += 32) {
            bool live = false;
            const int t = t0 + lane;
            if (t < n_tiles && (int64_t)(t + 1) * TILE > warp_first + 1) {
                const double4 b = tbox[t];
                const double gx = fmax(0.0, fmax(bx0 - b.z, b.x - bx1) - box_slack);
                const double gy = fmax(0.0, fmax(by0 - b.w, b.y - by1) - box_slack);
                live = (gx * gx + gy * gy) * (1.0 - 1e-12) <= max_d2;
            }
            unsigned todo = __ballot_sync(0xffffffffu, live);
            while (todo) {
                const int tl = __ffs(todo) - 1;
                todo &= todo - 1;
                if (!valid) continue;
                const int64_t j0 = max((int64_t)(t0 + tl) * TILE, i + 1), j1 = min((int64_t)(t0 + tl + 1) * TILE, n);
                for (int64_t j = j0; j < j1; ++j) {
                    const double2 pj = xy[j];
                    double d2;
                    if (ellipse) {
                        const double dx = __dsub_rn(pj.x, pi.x), dy = __dsub_rn(pj.y, pi.y);
                        const double n0 = __fma_rn(w01, dy, __dmul_rn(w00, dx));
                        const double n1 = __fma_rn(w11, dy, __dmul_rn(w10, dx));
                        d2 = __dadd_rn(__dmul_rn(n0, n0), __dmul_rn(n1, n1));
                    } else {
                        const double dx = __dsub_rn(pi.x, pj.x), dy = __dsub_rn(pi.y, pj.y);
                        d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                    }
                    if (!(d2 > 0.0) || d2 > max_d2) continue;
                    // fast check of the current run's lag, else search (lags are disjoint here: regular limits are required)
                    int b = run;
                    if (b < 0 || !(d2 <= up_d2[b] && d2 > low_d2[b])) {
                        int lo = 0, hi = n_lags - 1;
                        while (lo < hi) { const int mid = (lo + hi) >> 1; if (d2 <= up_d2[mid]) hi = mid; else lo = mid + 1; }
                        b = (d2 > low_d2[lo]) ? lo : -1;
                        if (b < 0) continue;
                        if (b != run) { flush(); run = b; }
                    }
                    const double zj = z[j], wj = wts[j];
                    const double wp = (wi * wj) / (wi + wj);
                    if (ellipse) { const double dz = zj - zi; r1 += wp * (dz * dz); r3 += wi * zi + wj * zj; }
                    else { const double uj = zj / wj, du = ui - uj; r1 += wp * (du * du); r3 += ui + uj; }
                    r2 += wp; r4 += wi + wj; ++rc;
                }
            }
        }
        flush();
    }
    if (use_smem) {
        __syncthreads();
        for (int k = threadIdx.x; k < n_lags; k += blockDim.x) {
            if (count[k]) {
                for (int q = 0; q < 4; ++q) atomicAdd(&g_sums[q * n_lags + k], sums[q * n_lags + k]);
                atomicAdd(&g_count[k], count[k]);
            }
        }
    }
}

int variogram_weighted_device(const double *d_xyz, const double *d_weights, int64_t n, const double *edges,
                              const double *lower, int n_lags, const double *W, int n_dirs, double *d_sums,
                              int64_t *d_count, cudaStream_t stream)
{
    PIB_CHECK(n >= 0 && n_lags >= 1 && edges && d_sums && d_count, PIB_ERR_INVALID, "weighted variogram: bad arguments");
    PIB_CHECK(n_dirs == 0 || (n_dirs == 1 && W), PIB_ERR_INVALID, "weighted variogram: one direction at a time");
    PIB_CUDA(cudaMemsetAsync(d_sums, 0, sizeof(double) * 4 * n_lags, stream));
    PIB_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int64_t) * n_lags, stream));
    if (n < 2) return PIB_OK;
    Scratch ws(stream);
    std::vector<double> up(n_lags), low(n_lags);
    for (int b = 0; b < n_lags; ++b) {
        up[b] = sq_threshold(edges[b]);
        low[b] = sq_threshold((n_dirs > 0 && lower) ? lower[b] : (b ? edges[b - 1] : 0.0));
    }
    double *d_up, *d_low;
    PIB_TRY(ws.alloc(&d_up, n_lags)); PIB_TRY(ws.alloc(&d_low, n_lags));
    PIB_CUDA(cudaMemcpyAsync(d_up, up.data(), n_lags * sizeof(double), cudaMemcpyHostToDevice, stream));
    PIB_CUDA(cudaMemcpyAsync(d_low, low.data(), n_lags * sizeof(double), cudaMemcpyHostToDevice, stream));
    for (int b = 0; b < n_lags; ++b)       // the run-length kernel assumes disjoint lags
        PIB_CHECK(low[b] == (b ? up[b - 1] : 0.0), PIB_ERR_UNSUPPORTED,
                  "weighted variogram: irregular ellipse limits (lag - (lag - prev) != prev) are not supported");
    SortedPoints sp;
    PIB_TRY(build_sorted_points(d_xyz, n, TILE, TILE, ws, stream, &sp));
    double *d_ws;
    PIB_TRY(ws.alloc(&d_ws, (size_t)n));
    gather_weights_kernel<<<(int)((n + 255) / 256), 256, 0, stream>>>(d_weights, sp.orig, n, d_ws);
    PIB_CUDA(cudaGetLastError());
    const int n_tiles = (int)((n + TILE - 1) / TILE);
    const double4 *d_tbox = sp.tile_box;
    double box_slack = 0.0;
    if (n_dirs > 0) {                                          // whitened boxes, as in the ring kernel
        double4 *d_boxw;
        PIB_TRY(ws.alloc(&d_boxw, n_tiles));
        tile_box_whitened_kernel<<<(int)(((int64_t)n_tiles * 32 + 255) / 256), 256, 0, stream>>>(sp.xy, n, n_tiles, W[0], W[1],
                                                                                                 W[2], W[3], d_boxw);
        PIB_CUDA(cudaGetLastError());
        d_tbox = d_boxw;
        PIB_TRY(sp.wait_box());
        const double pmax = std::fabs(sp.box[0]) + std::fabs(sp.box[1]) + std::fabs(sp.box[2]) + std::fabs(sp.box[3]);
        box_slack = 1e-12 * pmax * (std::fabs(W[0]) + std::fabs(W[1]) + std::fabs(W[2]) + std::fabs(W[3]));
    }
    PIB_TRY(sp.wait_box());
    const int use_smem = n_lags <= 2048;
    const size_t wt_smem = use_smem ? (size_t)n_lags * 7 * sizeof(double) : 0;
    if (wt_smem > 48 * 1024)
        PIB_CUDA(cudaFuncSetAttribute(weighted_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wt_smem));
    const int grid = (int)std::min<int64_t>((n + 127) / 128, 8 * sm_count());
    weighted_pair_kernel<<<grid, 128, wt_smem, stream>>>(sp.xy, sp.z, d_ws, n, d_tbox, n_tiles, box_slack, d_up, d_low, n_lags,
                                                         n_dirs > 0, n_dirs ? W[0] : 0, n_dirs ? W[1] : 0, n_dirs ? W[2] : 0,
                                                         n_dirs ? W[3] : 0, d_sums,
                                                         reinterpret_cast<unsigned long long *>(d_count), use_smem);
    PIB_CUDA(cudaGetLastError());
    return PIB_OK;
}


// ---------------------------------------------------------------------------------------------
// Triangle ("t") directional selection — the reference's DEFAULT method:
//   semivariogram/experimental/functions/directional.py:137-210 (from_triangle), :536-565
//   distance/angular.py:283-339 generate_triangles / get_triangles_vertices, :375-409 select_points_within_triangle,
//   :459-494 triangle_mask, :9-64 build_mask_indices, :188-218 clean_mask_indices
// For a centre p and lag h the reference tests every q against two triangles sharing the base through p
// (half width h*tol, perpendicular to the direction) with apexes at p +- h*(cos, sin): together a rhombus
//   R_h(p) = { q : |u| + |v| / tol <= h },  (u, v) = (q - p) in the rotated frame.
// The raw mask of lag k is "q in R_{h_k}(p)"; the cleaned mask is raw_k XOR raw_{k-1} (raw_0 for the first lag).
// Exact arithmetic would put q into the single lag k* with h_{k*-1} < rho <= h_{k*}; the reference decides with
// floating-point sign tests, so here
//   * rho = |u| + |v|/tol is computed per ordered pair; when rho is farther than `band` from every lag edge the
//     raw masks are certain (0 below k*, 1 from k* on) and the pair counts once, in lag k*;
//   * otherwise (and always for q == p, which sits on the shared base edge of every triangle) the raw masks of
//     the neighbouring lags are evaluated with the reference's own operation sequence (every +, -, * rounded
//     separately, vertices p + offset with the offsets h*cos(..) computed on the host by numpy like the
//     reference), so the cleaned masks are bit-identical.
// One thread per centre point p (caller order), all q; per-lag sums through global fp64 REDs.
// ---------------------------------------------------------------------------------------------
struct TriLag {                   // per lag, host-computed with the reference's numpy calls
    double h;
    double apex_x, apex_y;        // h * cos(a), h * sin(a)
    double inv_x, inv_y;          // (-h) * cos(a), (-h) * sin(a)
    double a_x, a_y;              // (h tol) * cos(a + pi/2), (h tol) * sin(a + pi/2)
    double c_x, c_y;              // (h tol) * cos(a - pi/2), (h tol) * sin(a - pi/2)
};

static __device__ __forceinline__ bool in_triangle_ref(double ax, double ay, double bx, double by, double cx, double cy,
                                                       double qx, double qy)
{
    // distance/angular.py:393-407, operation for operation
    const double s1 = __dsub_rn(__dmul_rn(__dsub_rn(qx, bx), __dsub_rn(ay, by)), __dmul_rn(__dsub_rn(ax, bx), __dsub_rn(qy, by)));
    const double s2 = __dsub_rn(__dmul_rn(__dsub_rn(qx, cx), __dsub_rn(by, cy)), __dmul_rn(__dsub_rn(bx, cx), __dsub_rn(qy, cy)));
    const double s3 = __dsub_rn(__dmul_rn(__dsub_rn(qx, ax), __dsub_rn(cy, ay)), __dmul_rn(__dsub_rn(cx, ax), __dsub_rn(qy, ay)));
    const bool t1 = s1 < 0, t2 = s2 < 0, t3 = s3 < 0;
    return (t1 && t2 && t3) || (!t1 && !t2 && !t3);
}

static __device__ __forceinline__ bool raw_mask_ref(const TriLag &L, double px, double py, double qx, double qy)
{
    const double ax = __dadd_rn(px, L.a_x), ay = __dadd_rn(py, L.a_y);          // base_a
    const double cx = __dadd_rn(px, L.c_x), cy = __dadd_rn(py, L.c_y);          // base_b
    const double bx = __dadd_rn(px, L.apex_x), by = __dadd_rn(py, L.apex_y);    // apex
    const double ix = __dadd_rn(px, L.inv_x), iy = __dadd_rn(py, L.inv_y);      // inverted apex
    return in_triangle_ref(ax, ay, bx, by, cx, cy, qx, qy) || in_triangle_ref(ax, ay, ix, iy, cx, cy, qx, qy);
}

// The lags of the ORDERED pair (centre p, neighbour q) under the triangle selection, reported through emit(lag, certain).
// A pair far from every rhombus edge belongs to exactly one lag — the first with rho = |u| + |v| / tol <= h — and is
// reported as certain; lags whose edge lies within the band of rho (and every lag for q == p, which sits on an edge of
// every triangle) are decided by the reference's own operation sequence and its XOR with the previous lag's mask
// (clean_mask_indices, distance/angular.py:188-218).  H(k) = lag distance k.
template <class LagH, class Emit>
static __device__ __forceinline__ void triangle_pair_lags(const TriLag *__restrict__ lags, int n_lags, LagH H, double h_max,
                                                          double cos_t, double sin_t, double inv_tol, double band_rel,
                                                          double px, double py, double qx, double qy, bool same_point,
                                                          Emit emit)
{
    const double dx = qx - px, dy = qy - py;
    const double u = dx * cos_t + dy * sin_t, v = dy * cos_t - dx * sin_t;
    const double rho = fabs(u) + fabs(v) * inv_tol;
    const double band = band_rel * (fabs(px) + fabs(py) + fabs(qx) + fabs(qy) + h_max) * (1.0 + inv_tol);
    if (rho > h_max + band) return;                             // outside every rhombus for certain
    // k* = first lag with rho <= h_k
    int lo = 0, hi = n_lags;                                    // hi == n_lags: beyond the last lag
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (rho <= H(mid)) hi = mid; else lo = mid + 1; }
    const int ks = lo;
    // lags whose edge is within the band of rho need the reference's own test; q == p needs it everywhere
    int k_first = ks, k_last = ks - 1;                          // empty range = everything certain
    if (same_point) { k_first = 0; k_last = n_lags - 1; }
    else {
        if (ks < n_lags && H(ks) - rho <= band) k_last = ks;
        if (ks > 0 && rho - H(ks - 1) <= band) { k_first = ks - 1; if (k_last < k_first) k_last = ks - 1; }
        // (edges farther away than the neighbours of k* are farther than the band too: lags ascend)
        while (k_first > 0 && rho - H(k_first - 1) <= band) --k_first;
        while (k_last + 1 < n_lags && H(k_last + 1) - rho <= band) ++k_last;
    }
    if (k_last < k_first) {
        if (ks < n_lags) emit(ks, true);
        return;
    }
    // uncertain range [k_first, k_last]: raw masks from the reference's arithmetic, certain outside it
    bool prev = (k_first > 0) ? (k_first - 1 >= ks) : false;   // raw_{k_first-1}: certain
    for (int k = k_first; k <= min(k_last + 1, n_lags - 1); ++k) {
        const bool raw = (k <= k_last) ? raw_mask_ref(lags[k], px, py, qx, qy) : (k >= ks);
        const bool in_lag = (k == 0) ? raw : (raw != prev);    // clean_mask_indices: XOR with the previous lag
        if (in_lag) emit(k, false);
        prev = raw;
    }
    // lags beyond k_last + 1 are certain and equal to their predecessor (both 0 or both 1): nothing to add,
    // except when k* lies beyond the uncertain range
    if (ks > k_last + 1 && ks < n_lags) emit(ks, false);
}

__global__ void __launch_bounds__(128) triangle_kernel(const double2 *__restrict__ xy, const double *__restrict__ z,
                                                       const int32_t *__restrict__ orig, int64_t n,
                                                       const double4 *__restrict__ rbox /* per tile, rotated (u, v) */,
                                                       int n_tiles, double cull_rho,
                                                       const TriLag *__restrict__ lags, int n_lags,
                                                       double cos_t, double sin_t, double inv_tol, double band_rel,
                                                       double *sums /* [3][n_lags]: sq, prod, neighbour values */,
                                                       unsigned long long *count, int use_smem)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const double h_max = lags[n_lags - 1].h;
    const int lane = threadIdx.x & 31;
    // CTA-private accumulators and the lag distances in shared memory (the per-lag global addresses are shared
    // by every SM; with global REDs they were the hottest lines of the L2); g_sums/g_count get one RED per lag per CTA
    extern __shared__ __align__(16) unsigned char tri_smem[];
    double *const g_sums = sums;
    unsigned long long *const g_count = count;
    const double *lag_h = nullptr;
    if (use_smem) {
        double *sh = reinterpret_cast<double *>(tri_smem);
        for (int k = threadIdx.x; k < n_lags; k += blockDim.x) sh[k] = lags[k].h;
        for (int k = threadIdx.x; k < 4 * n_lags; k += blockDim.x) sh[n_lags + k] = 0.0;     // 3 sums + count (bits of 0)
        lag_h = sh;
        sums = sh + n_lags;
        count = reinterpret_cast<unsigned long long *>(sh + 4 * n_lags);
        __syncthreads();
    }
    auto H = [&](int k) -> double { return lag_h ? lag_h[k] : lags[k].h; };
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < n; base += stride) {
        const int64_t i = base + threadIdx.x;
        const bool valid = i < n;
        const int64_t ic = valid ? i : n - 1;
        const double px = xy[ic].x, py = xy[ic].y, pz = z[ic];
        // box of my warp's 32 centre points in the rotated frame (u along the direction, v across)
        double u0 = px * cos_t + py * sin_t, v0 = py * cos_t - px * sin_t, u1 = u0, v1 = v0;
        for (int o = 16; o; o >>= 1) {
            u0 = fmin(u0, __shfl_xor_sync(0xffffffffu, u0, o)); u1 = fmax(u1, __shfl_xor_sync(0xffffffffu, u1, o));
            v0 = fmin(v0, __shfl_xor_sync(0xffffffffu, v0, o)); v1 = fmax(v1, __shfl_xor_sync(0xffffffffu, v1, o));
        }
        // certain pairs of the same lag come in runs (Hilbert order): accumulate a run in registers
        int run = -1;
        double r1 = 0, r2 = 0, r3 = 0;
        unsigned long long rc = 0;
        auto flush = [&]() {
            if (run >= 0 && rc) {
                atomicAdd(&sums[run], r1); atomicAdd(&sums[n_lags + run], r2); atomicAdd(&sums[2 * n_lags + run], r3);
                atomicAdd(&count[run], rc);
            }
            r1 = r2 = r3 = 0; rc = 0;
        };
        // tiles whose rotated box lies outside every rhombus of every point of the warp are skipped: the union of the
        // lags' double triangles is the rhombus |u| + |v| / tol <= h_max (cull_rho = h_max + the certainty band + slack)
        for (int t0 = 0; t0 < n_tiles; t0 += 32) {
            bool live = false;
            if (t0 + lane < n_tiles) {
                const double4 b = rbox[t0 + lane];
                const double gu = fmax(0.0, fmax(u0 - b.z, b.x - u1)), gv = fmax(0.0, fmax(v0 - b.w, b.y - v1));
                live = gu + gv * inv_tol <= cull_rho;
            }
            unsigned todo = __ballot_sync(0xffffffffu, live);
            while (todo) {
                const int tl = __ffs(todo) - 1;
                todo &= todo - 1;
                const int64_t j0 = (int64_t)(t0 + tl) * TILE, j1 = min(j0 + (int64_t)TILE, n);
                if (!valid) continue;
        for (int64_t j = j0; j < j1; ++j) {
            const double qx = xy[j].x, qy = xy[j].y, qz = z[j];
            const double dz = pz - qz;
            triangle_pair_lags(lags, n_lags, H, h_max, cos_t, sin_t, inv_tol, band_rel, px, py, qx, qy, orig[j] == orig[i],
                               [&](int k, bool certain) {
                if (certain) {                                      // counted once, in lag k: runs accumulate in registers
                    if (k != run) { flush(); run = k; }
                    r1 += dz * dz; r2 += pz * qz; r3 += qz; ++rc;
                } else {
                    atomicAdd(&sums[k], dz * dz);
                    atomicAdd(&sums[n_lags + k], pz * qz);
                    atomicAdd(&sums[2 * n_lags + k], qz);
                    atomicAdd(&count[k], 1ull);
                }
            });
        }
            }
        }
        flush();
    }
    if (use_smem) {
        __syncthreads();
        for (int k = threadIdx.x; k < n_lags; k += blockDim.x) {
            if (count[k]) {
                atomicAdd(&g_sums[k], sums[k]); atomicAdd(&g_sums[n_lags + k], sums[n_lags + k]);
                atomicAdd(&g_sums[2 * n_lags + k], sums[2 * n_lags + k]);
                atomicAdd(&g_count[k], count[k]);
            }
        }
    }
}

int variogram_triangle_device(const double *d_xyz, int64_t n, const double *tri_lags /* host [n_lags][9] */, int n_lags,
                              double cos_t, double sin_t, double tolerance, double *d_sums, int64_t *d_count,
                              cudaStream_t stream)
{
    PIB_CHECK(n >= 0 && n_lags >= 1 && tri_lags && d_sums && d_count, PIB_ERR_INVALID, "triangle variogram: bad arguments");
    PIB_CHECK(tolerance > 0, PIB_ERR_INVALID, "triangle variogram: tolerance must be positive");
    static_assert(sizeof(TriLag) == 9 * sizeof(double), "TriLag is nine doubles");
    PIB_CUDA(cudaMemsetAsync(d_sums, 0, sizeof(double) * 3 * n_lags, stream));
    PIB_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int64_t) * n_lags, stream));
    if (n < 1) return PIB_OK;
    Scratch ws(stream);
    TriLag *d_lags;
    PIB_TRY(ws.alloc(&d_lags, n_lags));
    PIB_CUDA(cudaMemcpyAsync(d_lags, tri_lags, sizeof(TriLag) * n_lags, cudaMemcpyHostToDevice, stream));
    SortedPoints sp;
    PIB_TRY(build_sorted_points(d_xyz, n, TILE, TILE, ws, stream, &sp));
    const int n_tiles = (int)((n + TILE - 1) / TILE);
    double4 *d_rbox;
    PIB_TRY(ws.alloc(&d_rbox, n_tiles));
    tile_box_whitened_kernel<<<(int)(((int64_t)n_tiles * 32 + 255) / 256), 256, 0, stream>>>(sp.xy, n, n_tiles, cos_t, sin_t,
                                                                                             -sin_t, cos_t, d_rbox);
    PIB_CUDA(cudaGetLastError());
    // every pair the kernel keeps has rho <= h_max + band; the boxes are of u = x cos + y sin, the pairs use
    // (x_j - x_i) cos + ...: a few ulps of |coordinates| apart, covered by the same 1e-9-relative band once more
    const double h_max = tri_lags[9 * (n_lags - 1)];
    PIB_TRY(sp.wait_box());
    const double cmax = std::max(std::fabs(sp.box[0]), std::fabs(sp.box[2])) + std::max(std::fabs(sp.box[1]), std::fabs(sp.box[3]));
    const double band_max = 1e-9 * (2.0 * cmax + h_max) * (1.0 + 1.0 / tolerance);
    const double cull_rho = h_max + 2.0 * band_max;
    const int grid = (int)std::min<int64_t>((n + 127) / 128, 8 * sm_count());
    const int use_smem = n_lags <= 2048;
    const size_t tri_smem = use_smem ? (size_t)n_lags * 5 * sizeof(double) : 0;
    if (tri_smem > 48 * 1024)
        PIB_CUDA(cudaFuncSetAttribute(triangle_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tri_smem));
    triangle_kernel<<<grid, 128, tri_smem, stream>>>(sp.xy, sp.z, sp.orig, n, d_rbox, n_tiles, cull_rho, d_lags, n_lags, cos_t,
                                                     sin_t, 1.0 / tolerance, 1e-9, d_sums,
                                                     reinterpret_cast<unsigned long long *>(d_count), use_smem);
    PIB_CUDA(cudaGetLastError());
    return PIB_OK;
}

// ---------------------------------------------------------------------------------------------
// Triangle-method point cloud (SURVEY section 8f-2 with the reference's DEFAULT directional method):
//   semivariogram/experimental/functions/directional.py:213-273 (from_triangle_cloud), :583-620
// Per lag the values (z_i - z_j)^2 of the cleaned masks, row i ascending and j ascending within a row (points[mask]),
// the pair q == p of the first lag included.  Same two passes as cloud_kernel (count, scan, write) on the points in
// the caller's order; membership comes from triangle_pair_lags, i.e. exactly what triangle_kernel counts.
// ---------------------------------------------------------------------------------------------
template <bool WRITE>
__global__ void triangle_cloud_kernel(const double *__restrict__ xyz, int64_t n, const TriLag *__restrict__ lags, int n_lags,
                                      double cos_t, double sin_t, double inv_tol, double band_rel,
                                      unsigned long long *__restrict__ cursor /* [n_lags][n] */, double *__restrict__ values,
                                      unsigned long long n_values)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2];
    const double h_max = lags[n_lags - 1].h;
    auto H = [&](int k) -> double { return lags[k].h; };
    for (int64_t j = 0; j < n; ++j) {
        const double qx = xyz[3 * j], qy = xyz[3 * j + 1];
        triangle_pair_lags(lags, n_lags, H, h_max, cos_t, sin_t, inv_tol, band_rel, px, py, qx, qy, j == i, [&](int k, bool) {
            unsigned long long *c = cursor + (size_t)k * n + i;
            if (WRITE) {
                const double dz = pz - xyz[3 * j + 2];
                PIB_DBG_ASSERT(*c < n_values);
                values[(*c)++] = dz * dz;
            } else ++(*c);
        });
    }
}

int variogram_triangle_cloud_device(const double *d_xyz, int64_t n, const double *tri_lags /* host [n_lags][9] */, int n_lags,
                                    double cos_t, double sin_t, double tolerance, const int64_t *lag_offsets /* host [n_lags + 1] */,
                                    double *d_values, cudaStream_t stream)
{
    PIB_CHECK(n >= 0 && n_lags >= 1 && tri_lags && lag_offsets, PIB_ERR_INVALID, "triangle cloud: bad arguments");
    PIB_CHECK(tolerance > 0, PIB_ERR_INVALID, "triangle cloud: tolerance must be positive");
    if (n < 1) return PIB_OK;
    Scratch ws(stream);
    TriLag *d_lags;
    int64_t *d_off;
    unsigned long long *d_cursor, *d_end;
    PIB_TRY(ws.alloc(&d_lags, n_lags)); PIB_TRY(ws.alloc(&d_off, n_lags)); PIB_TRY(ws.alloc(&d_end, n_lags));
    PIB_TRY(ws.alloc(&d_cursor, (size_t)n_lags * n));
    PIB_CUDA(cudaMemcpyAsync(d_lags, tri_lags, sizeof(TriLag) * n_lags, cudaMemcpyHostToDevice, stream));
    PIB_CUDA(cudaMemcpyAsync(d_off, lag_offsets, n_lags * sizeof(int64_t), cudaMemcpyHostToDevice, stream));
    PIB_CUDA(cudaMemsetAsync(d_cursor, 0, sizeof(unsigned long long) * n_lags * n, stream));
    const int grid = (int)((n + 127) / 128);
    triangle_cloud_kernel<false><<<grid, 128, 0, stream>>>(d_xyz, n, d_lags, n_lags, cos_t, sin_t, 1.0 / tolerance, 1e-9, d_cursor, nullptr, 0);
    PIB_CUDA(cudaGetLastError());
    cloud_scan_kernel<<<n_lags, 1024, 0, stream>>>(d_cursor, n, d_off, d_end);
    PIB_CUDA(cudaGetLastError());
    // the caller sized `values` from the counts of pib_variogram_triangle_host: refuse to write if this pass disagrees
    std::vector<unsigned long long> end(n_lags);
    PIB_CUDA(cudaMemcpyAsync(end.data(), d_end, sizeof(unsigned long long) * n_lags, cudaMemcpyDeviceToHost, stream));
    PIB_CUDA(cudaStreamSynchronize(stream));
    for (int b = 0; b < n_lags; ++b)
        PIB_CHECK((int64_t)end[b] == lag_offsets[b + 1], PIB_ERR_INVALID, "triangle cloud: lag_offsets do not match the pair counts");
    triangle_cloud_kernel<true><<<grid, 128, 0, stream>>>(d_xyz, n, d_lags, n_lags, cos_t, sin_t, 1.0 / tolerance, 1e-9, d_cursor, d_values,
                                                           (unsigned long long)lag_offsets[n_lags]);
    PIB_CUDA(cudaGetLastError());
    return PIB_OK;
}

}  // namespace pib
