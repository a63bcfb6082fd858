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
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace pib {

constexpr int TILE = 128;        // points per tile
constexpr int CHUNK = 64;        // B tiles per work item
constexpr int LUT_N = 1024;      // distance -> first candidate slot table
constexpr int MAX_SMEM = 232448; // opt-in dynamic shared memory per CTA on sm_100

struct PairParams {
    const double2 *xy;           // [n_pad]
    const double *z;             // [n_pad]
    const double4 *tile_box;     // [n_tiles]
    int64_t n;                   // real points
    int n_tiles;
    int n_slots;                 // n_lags + 2: slot 0 = "d2 <= 0 / masked", slot s = lag s-1, last = beyond range
    const double *upper;         // [n_slots] slot upper thresholds on d2: U[0] = 0, U[s] = T(edge[s-1]), U[last] = +inf
    const uint16_t *lut;         // [LUT_N]
    float lut_scale;
    double max_d2;               // T(edges[n_lags-1])
    double w00, w01, w10, w11;   // whitening matrix (ellipse)
    int shard_index, shard_count;
    double *partial;             // [grid][warps][n_slots][4]  (sum_sq, sum_prod, sum_val, count as double-exact int64 bits)
    unsigned long long *pairs_evaluated;
};

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
static __device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
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

// Shared memory layout (dynamic):
//   double2  acc[n_slots][NT]     private (sum dz, sum dz^2)
//   uint32_t cnt[n_slots][NT]     private pair counts
//   double   upper[n_slots]
//   uint16_t lut[LUT_N]
//   double2  txy[2][TILE]; double tz[2][TILE]    B tile double buffer
//   uint64_t bar[2]; uint64_t live_mask
template <int NT>
static size_t pair_smem_bytes(int n_slots)
{
    size_t b = (size_t)n_slots * NT * (sizeof(double2) + sizeof(uint32_t));
    b += ((size_t)n_slots * sizeof(double) + 15) / 16 * 16;
    b += LUT_N * sizeof(uint16_t);
    b += 2 * TILE * (sizeof(double2) + sizeof(double));
    b += 64;
    return b;
}

template <int NT, bool ELLIPSE>
__global__ void __launch_bounds__(NT, 1) pair_kernel(const PairParams p)
{
    static_assert(TILE % NT == 0, "an A sub-tile is NT consecutive points of one tile");
    constexpr int A_PER_TILE = TILE / NT;          // thread t owns point t of the A sub-tile
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_slots = p.n_slots;

    double2 *acc = reinterpret_cast<double2 *>(smem_raw);
    uint32_t *cnt = reinterpret_cast<uint32_t *>(acc + (size_t)n_slots * NT);
    double *upper = reinterpret_cast<double *>(cnt + (size_t)n_slots * NT);
    uint16_t *lut = reinterpret_cast<uint16_t *>(upper + (n_slots + 1) / 2 * 2);
    double2 *txy = reinterpret_cast<double2 *>(lut + LUT_N);
    double *tz = reinterpret_cast<double *>(txy + 2 * TILE);
    uint64_t *bar = reinterpret_cast<uint64_t *>(tz + 2 * TILE);
    uint64_t *live_mask = bar + 2;

    for (int s = 0; s < n_slots; ++s) {
        acc[s * NT + tid] = make_double2(0.0, 0.0);
        cnt[s * NT + tid] = 0u;
    }
    for (int s = tid; s < n_slots; s += NT) upper[s] = p.upper[s];
    for (int s = tid; s < LUT_N; s += NT) lut[s] = p.lut[s];
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int n_tiles = p.n_tiles;
    const int n_chunks = (n_tiles + CHUNK - 1) / CHUNK;
    const int64_t n_items = (int64_t)n_tiles * A_PER_TILE * n_chunks;
    uint32_t seq = 0;                       // tiles staged so far by this CTA (buffer = seq & 1, phase = (seq >> 1) & 1)
    unsigned long long evaluated = 0;
    double *my_partial = p.partial + ((size_t)blockIdx.x * (NT / 32) + warp) * n_slots * 4;
    const float lut_scale = p.lut_scale;

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
                    // margin: ellipse norms are >= the plain distance up to rounding of the rotation
                    live = box_min_d2(abox, bbox) * (1.0 - 1e-12) <= p.max_d2;
                }
                m |= (uint64_t)__ballot_sync(0xffffffffu, live) << (32 * h);
            }
            if (lane == 0) *live_mask = m;
        }
        __syncthreads();
        uint64_t live = *live_mask;
        if (live == 0) { __syncthreads(); continue; }

        // ---- my A points -------------------------------------------------------------------
        const int ig = (int)a0 + tid;
        const double2 pa = p.xy[ig];
        const double xi = pa.x, yi = pa.y, zi = p.z[ig];

        // prefetch the first live tile
        int tb = tb0 + __ffsll((long long)live) - 1;
        live &= live - 1;
        if (tid == 0) {
            const uint32_t b = seq & 1;
            mbar_expect_tx(&bar[b], TILE * (sizeof(double2) + sizeof(double)));
            tma_load_1d(txy + b * TILE, p.xy + (size_t)tb * TILE, TILE * sizeof(double2), &bar[b]);
            tma_load_1d(tz + b * TILE, p.z + (size_t)tb * TILE, TILE * sizeof(double), &bar[b]);
        }
        while (true) {
            const uint32_t b = seq & 1, phase = (seq >> 1) & 1;
            const int tb_next = live ? tb0 + __ffsll((long long)live) - 1 : -1;
            live &= live - 1;
            if (tid == 0 && tb_next >= 0) {                   // buffer b^1 was released by the barrier below
                const uint32_t nb = b ^ 1;
                mbar_expect_tx(&bar[nb], TILE * (sizeof(double2) + sizeof(double)));
                tma_load_1d(txy + nb * TILE, p.xy + (size_t)tb_next * TILE, TILE * sizeof(double2), &bar[nb]);
                tma_load_1d(tz + nb * TILE, p.z + (size_t)tb_next * TILE, TILE * sizeof(double), &bar[nb]);
            }
            mbar_wait(&bar[b], phase);
            ++seq;

            const double2 *bxy = txy + b * TILE;
            const double *bz = tz + b * TILE;
            const bool diagonal = (tb == ta);
            const int jbase = tb * TILE;
            if (tid == 0) {
                const long long vb = min((int64_t)TILE, p.n - (int64_t)tb * TILE), va = valid_a;
                // diagonal: pairs with j > i, i in [a0, a0+va), j in [jbase, jbase+vb)
                evaluated += diagonal ? (unsigned long long)(va * (jbase + vb - 1) - va * a0 - va * (va - 1) / 2)
                                      : (unsigned long long)(va * vb);
            }
#pragma unroll 2
            for (int j = 0; j < TILE; ++j) {
                const double2 pj = bxy[j];
                const double zj = bz[j];
                {
                    double d2;
                    if (ELLIPSE) {
                        // vector_distance = other - centre (angular.py:448); W @ d as OpenBLAS dgemm
                        // evaluates it: fma(W[r][1], dy, W[r][0]*dx); einsum: n0*n0 + n1*n1 un-fused
                        const double dx = __dsub_rn(pj.x, xi), dy = __dsub_rn(pj.y, yi);
                        const double n0 = __fma_rn(p.w01, dy, __dmul_rn(p.w00, dx));
                        const double n1 = __fma_rn(p.w11, dy, __dmul_rn(p.w10, dx));
                        d2 = __dadd_rn(__dmul_rn(n0, n0), __dmul_rn(n1, n1));
                    } else {
                        // scipy cdist euclidean: sqrt(dx*dx + dy*dy), nothing fused (distance/point.py:56)
                        const double dx = __dsub_rn(xi, pj.x), dy = __dsub_rn(yi, pj.y);
                        d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                    }
                    // first candidate slot from a float estimate of the distance, then exact fix-up
                    float r;
                    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__double2float_rn(d2)));
                    int e = __float2int_rz(r * lut_scale);
                    e = min(e, LUT_N - 1);
                    int s = lut[e];
                    while (d2 > upper[s]) ++s;
                    if (diagonal && !(jbase + j > ig)) s = 0;     // keep i < j only
                    const double dz = __dsub_rn(zi, zj);
                    double2 a = acc[s * NT + tid];
                    a.x += dz;
                    a.y = __fma_rn(dz, dz, a.y);
                    acc[s * NT + tid] = a;
                    cnt[s * NT + tid] += 1u;
                }
            }
            __syncthreads();                                  // everyone is done with buffer b
            if (tb_next < 0) break;
            tb = tb_next;
        }

        // ---- fold this item's private histograms (my z_i is fixed within the item) ------------
        //   sum z_i z_j   = sum z_i (z_i - dz) = c z_i^2 - z_i S1
        //   sum (z_i+z_j) = 2 c z_i - S1
        for (int s = 1; s < n_slots - 1; ++s) {
            const double2 a = acc[s * NT + tid];
            const uint32_t c = cnt[s * NT + tid];
            const double cd = (double)c;
            double v_sq = a.y;
            double v_prod = __fma_rn(-zi, a.x, cd * zi * zi);
            double v_val = __fma_rn(2.0 * cd, zi, -a.x);
            unsigned long long v_cnt = c;
            // any lane with something to add?  (most slots are empty for a given item)
            if (__ballot_sync(0xffffffffu, c != 0u) == 0u) continue;
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
            acc[s * NT + tid] = make_double2(0.0, 0.0);
            cnt[s * NT + tid] = 0u;
        }
        // trash slots are never read; reset their counters so they cannot overflow
        cnt[tid] = 0u; cnt[(n_slots - 1) * NT + tid] = 0u;
        acc[tid] = make_double2(0.0, 0.0); acc[(n_slots - 1) * NT + tid] = make_double2(0.0, 0.0);
    }
    if (tid == 0 && evaluated) atomicAdd(p.pairs_evaluated, evaluated);
}

// Fixed-order sum of the per-warp partials -> outputs (one thread per lag)
__global__ void reduce_partials_kernel(const double *__restrict__ partial, int n_parts, int n_slots, int n_lags,
                                       double *__restrict__ sum_sq, double *__restrict__ sum_prod,
                                       double *__restrict__ sum_val, int64_t *__restrict__ count)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_lags) return;
    double s0 = 0, s1 = 0, s2 = 0;
    unsigned long long c = 0;
    for (int q = 0; q < n_parts; ++q) {
        const double *src = partial + ((size_t)q * n_slots + (b + 1)) * 4;
        s0 += src[0]; s1 += src[1]; s2 += src[2];
        c += reinterpret_cast<const unsigned long long *>(src)[3];
    }
    sum_sq[b] = s0; sum_prod[b] = s1; sum_val[b] = s2; count[b] = (int64_t)c;
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

template <int NT, bool ELLIPSE>
static int launch_pair_kernel(const PairParams &p, size_t smem, int grid, cudaStream_t stream)
{
    PIB_CUDA(cudaFuncSetAttribute(pair_kernel<NT, ELLIPSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pair_kernel<NT, ELLIPSE><<<grid, NT, smem, stream>>>(p);
    PIB_CUDA(cudaGetLastError());
    return PIB_OK;
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
    if (pair_smem_bytes<128>(n_slots) <= (size_t)MAX_SMEM) nt = 128;
    else if (pair_smem_bytes<64>(n_slots) <= (size_t)MAX_SMEM) nt = 64;
    else if (pair_smem_bytes<32>(n_slots) <= (size_t)MAX_SMEM) nt = 32;
    const char *force = std::getenv("PIB_FORCE_GENERIC");
    const bool generic = !regular || nt == 0 || (force && force[0] == '1');

    if (generic) {
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
        // distance -> first candidate slot: cell e covers r in [e, e+1) / scale; start from the slot of
        // (e - 0.5) / scale so float rounding can never start above the true slot
        const double rmax = edges[n_lags - 1];
        const double scale = (LUT_N - 2) / rmax;
        std::vector<uint16_t> lut(LUT_N);
        for (int e = 0; e < LUT_N; ++e) {
            const double r_lo = std::max(0.0, (e - 0.5) / scale);
            const double t = r_lo * r_lo * (1.0 - 1e-6);
            int s = 0;
            while (s < n_slots - 1 && t > up[s]) ++s;
            lut[e] = (uint16_t)s;
        }
        double *d_upper;
        uint16_t *d_lut;
        PIB_TRY(ws.alloc(&d_upper, n_slots)); PIB_TRY(ws.alloc(&d_lut, LUT_N));
        PIB_CUDA(cudaMemcpyAsync(d_upper, up.data(), n_slots * sizeof(double), cudaMemcpyHostToDevice, stream));
        PIB_CUDA(cudaMemcpyAsync(d_lut, lut.data(), LUT_N * sizeof(uint16_t), cudaMemcpyHostToDevice, stream));

        const int n_tiles = (int)(sp.n_pad / TILE);
        const int n_chunks = (n_tiles + CHUNK - 1) / CHUNK;
        const int64_t n_items = ((int64_t)n_tiles * (TILE / nt) * n_chunks + shard_count - 1) / shard_count;
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(sm_count(), n_items));
        const int warps = nt / 32;
        double *d_partial;
        const size_t part_n = (size_t)grid * warps * n_slots * 4;
        PIB_TRY(ws.alloc(&d_partial, part_n));

        PairParams p;
        p.xy = sp.xy; p.z = sp.z; p.tile_box = sp.tile_box; p.n = n; p.n_tiles = n_tiles; p.n_slots = n_slots;
        p.upper = d_upper; p.lut = d_lut; p.lut_scale = (float)scale; p.max_d2 = up_lag[n_lags - 1];
        p.shard_index = shard_index; p.shard_count = shard_count; p.partial = d_partial;
        p.pairs_evaluated = d_evaluated;
        for (int t = 0; t < out_dirs; ++t) {
            PIB_CUDA(cudaMemsetAsync(d_partial, 0, part_n * sizeof(double), stream));
            int rc;
            if (n_dirs > 0) {
                p.w00 = W[4 * t]; p.w01 = W[4 * t + 1]; p.w10 = W[4 * t + 2]; p.w11 = W[4 * t + 3];
                rc = nt == 128 ? launch_pair_kernel<128, true>(p, pair_smem_bytes<128>(n_slots), grid, stream)
                   : nt == 64 ? launch_pair_kernel<64, true>(p, pair_smem_bytes<64>(n_slots), grid, stream)
                              : launch_pair_kernel<32, true>(p, pair_smem_bytes<32>(n_slots), grid, stream);
            } else {
                p.w00 = p.w01 = p.w10 = p.w11 = 0;
                rc = nt == 128 ? launch_pair_kernel<128, false>(p, pair_smem_bytes<128>(n_slots), grid, stream)
                   : nt == 64 ? launch_pair_kernel<64, false>(p, pair_smem_bytes<64>(n_slots), grid, stream)
                              : launch_pair_kernel<32, false>(p, pair_smem_bytes<32>(n_slots), grid, stream);
            }
            PIB_TRY(rc);
            reduce_partials_kernel<<<(n_lags + 127) / 128, 128, 0, stream>>>(
                d_partial, grid * warps, n_slots, n_lags, d_sum_sq + (size_t)t * n_lags, d_sum_prod + (size_t)t * n_lags,
                d_sum_val + (size_t)t * n_lags, d_count + (size_t)t * n_lags);
            PIB_CUDA(cudaGetLastError());
        }
    }
    if (pairs_evaluated) {
        unsigned long long h = 0;
        PIB_CUDA(cudaMemcpyAsync(&h, d_evaluated, sizeof(h), cudaMemcpyDeviceToHost, stream));
        PIB_CUDA(cudaStreamSynchronize(stream));
        *pairs_evaluated = (int64_t)h;
    }
    return PIB_OK;
}

}  // namespace pib
