// Spatial preprocessing shared by the variogram and kriging paths:
//   * Morton-ordered SoA copy of the points (compact tiles => tile-pair culling and few lags per tile pair)
//   * uniform grid (cell-sorted SoA + cell offsets) for exact k-nearest-neighbour search
// Sorting uses CUB's radix sort from the CUDA toolkit (stable => deterministic layouts); everything
// else is hand-written.  Nothing here is on the per-pair / per-location hot loop.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"

namespace pib {

// ---------------------------------------------------------------------------------------------
// bounding box of the x / y columns of [n][3] (stride 3) or [m][2] (stride 2) rows
// ---------------------------------------------------------------------------------------------
__global__ void bbox_partial_kernel(const double *__restrict__ pts, int64_t n, int stride,
                                    double4 *__restrict__ partial)
{
    double xmin = INFINITY, ymin = INFINITY, xmax = -INFINITY, ymax = -INFINITY;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = pts[i * stride], y = pts[i * stride + 1];
        xmin = fmin(xmin, x); xmax = fmax(xmax, x);
        ymin = fmin(ymin, y); ymax = fmax(ymax, y);
    }
    for (int o = 16; o; o >>= 1) {
        xmin = fmin(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
        ymin = fmin(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
        xmax = fmax(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        ymax = fmax(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
    }
    __shared__ double4 s[32];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) s[w] = make_double4(xmin, ymin, xmax, ymax);
    __syncthreads();
    if (threadIdx.x == 0) {
        double4 r = s[0];
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) {
            r.x = fmin(r.x, s[i].x); r.y = fmin(r.y, s[i].y);
            r.z = fmax(r.z, s[i].z); r.w = fmax(r.w, s[i].w);
        }
        partial[blockIdx.x] = r;
    }
}

__global__ void __launch_bounds__(32) bbox_final_kernel(const double4 *__restrict__ partial, int n_partial, double4 *__restrict__ box)
{
    double xmin = INFINITY, ymin = INFINITY, xmax = -INFINITY, ymax = -INFINITY;
    for (int i = threadIdx.x; i < n_partial; i += 32) {
        const double4 p = partial[i];
        xmin = fmin(xmin, p.x); ymin = fmin(ymin, p.y); xmax = fmax(xmax, p.z); ymax = fmax(ymax, p.w);
    }
    for (int o = 16; o; o >>= 1) {
        xmin = fmin(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
        ymin = fmin(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
        xmax = fmax(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        ymax = fmax(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
    }
    if (threadIdx.x == 0) *box = make_double4(xmin, ymin, xmax, ymax);
}

// bounding box left on the device (no host round trip): the variogram path only needs it to scale the curve keys
static int bounding_box_device(const double *d_pts, int64_t n, int stride, Scratch &ws, cudaStream_t stream, double4 **d_box)
{
    const int blocks = (int)std::min<int64_t>(296, (n + 255) / 256);
    double4 *d_partial;
    PIB_TRY(ws.alloc(&d_partial, blocks));
    PIB_TRY(ws.alloc(d_box, 1));
    bbox_partial_kernel<<<blocks, 256, 0, stream>>>(d_pts, n, stride, d_partial);
    bbox_final_kernel<<<1, 32, 0, stream>>>(d_partial, blocks, *d_box);
    PIB_CUDA(cudaGetLastError());
    return PIB_OK;
}

static int bounding_box(const double *d_pts, int64_t n, int stride, Scratch &ws, cudaStream_t stream,
                        double box[4])
{
    const int blocks = (int)std::min<int64_t>(296, (n + 255) / 256);
    double4 *d_partial;
    PIB_TRY(ws.alloc(&d_partial, blocks));
    bbox_partial_kernel<<<blocks, 256, 0, stream>>>(d_pts, n, stride, d_partial);
    PIB_CUDA(cudaGetLastError());
    std::vector<double4> h(blocks);
    PIB_CUDA(cudaMemcpyAsync(h.data(), d_partial, sizeof(double4) * blocks, cudaMemcpyDeviceToHost, stream));
    PIB_CUDA(cudaStreamSynchronize(stream));
    box[0] = box[1] = INFINITY; box[2] = box[3] = -INFINITY;
    for (auto &p : h) {
        box[0] = std::min(box[0], p.x); box[1] = std::min(box[1], p.y);
        box[2] = std::max(box[2], p.z); box[3] = std::max(box[3], p.w);
    }
    PIB_CHECK(std::isfinite(box[0]) && std::isfinite(box[1]) && std::isfinite(box[2]) && std::isfinite(box[3]),
              PIB_ERR_INVALID, "coordinates must be finite");
    return PIB_OK;
}

// ---------------------------------------------------------------------------------------------
// Hilbert order (16 bits per axis).  Unlike a Morton curve the Hilbert curve never jumps: any run of
// consecutive points is spatially compact, so every 128-point tile has a tight bounding box.
// ---------------------------------------------------------------------------------------------
static __device__ __forceinline__ uint32_t hilbert16(uint32_t x, uint32_t y)
{
    uint32_t d = 0;
    for (uint32_t s = 32768u; s > 0; s >>= 1) {
        const uint32_t rx = (x & s) ? 1u : 0u, ry = (y & s) ? 1u : 0u;
        d += s * s * ((3u * rx) ^ ry);
        if (ry == 0) {
            if (rx == 1) { x = 65535u - x; y = 65535u - y; }
            const uint32_t t = x; x = y; y = t;
        }
    }
    return d;
}

__global__ void curve_key_kernel(const double *__restrict__ xyz, int64_t n, const double4 *__restrict__ box,
                                  uint32_t *__restrict__ keys, int32_t *__restrict__ vals)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double4 b = *box;
    const double x0 = b.x, y0 = b.y, extent = fmax(b.z - b.x, b.w - b.y);
    const double scale = extent > 0 ? 65535.999 / extent : 0.0;
    const int ix = clampi((int)((xyz[3 * i] - x0) * scale), 0, 65535);
    const int iy = clampi((int)((xyz[3 * i + 1] - y0) * scale), 0, 65535);
    keys[i] = hilbert16((uint32_t)ix, (uint32_t)iy);
    vals[i] = (int32_t)i;
}

// pads: far away (but finite, so pad-pad distances are exactly 0 and pad-real distances are huge)
#define PIB_PAD_COORD 1.0e150

__global__ void gather_soa_kernel(const double *__restrict__ xyz, const int32_t *__restrict__ order,
                                  int64_t n, int64_t n_pad, double2 *__restrict__ xy,
                                  double *__restrict__ z, int32_t *__restrict__ orig)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    if (i < n) {
        const int32_t o = order[i];
        xy[i] = make_double2(xyz[3 * (int64_t)o], xyz[3 * (int64_t)o + 1]); z[i] = xyz[3 * (int64_t)o + 2];
        orig[i] = o;
    } else {
        xy[i] = make_double2(PIB_PAD_COORD, PIB_PAD_COORD); z[i] = 0.0; orig[i] = -1;
    }
}

// one warp per tile
__global__ void tile_box_kernel(const double2 *__restrict__ xy, int64_t n,
                                int tile, int64_t n_tiles, double4 *__restrict__ box)
{
    const int64_t t = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= n_tiles) return;
    double xmin = INFINITY, ymin = INFINITY, xmax = -INFINITY, ymax = -INFINITY;
    for (int k = lane; k < tile; k += 32) {
        const int64_t i = t * tile + k;
        if (i < n) {
            const double2 p = xy[i];
            xmin = fmin(xmin, p.x); xmax = fmax(xmax, p.x);
            ymin = fmin(ymin, p.y); ymax = fmax(ymax, p.y);
        }
    }
    for (int o = 16; o; o >>= 1) {
        xmin = fmin(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
        ymin = fmin(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
        xmax = fmax(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        ymax = fmax(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
    }
    if (lane == 0) box[t] = make_double4(xmin, ymin, xmax, ymax);   // empty tile: (+inf,+inf,-inf,-inf)
}

template <typename K>
static int radix_sort_pairs(const K *keys_in, K *keys_out, const int32_t *vals_in, int32_t *vals_out,
                            int64_t n, int end_bit, Scratch &ws, cudaStream_t stream)
{
    PIB_CHECK(n < (int64_t)1 << 31, PIB_ERR_UNSUPPORTED, "more than 2^31-1 points");
    size_t tmp_bytes = 0;
    PIB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in, keys_out, vals_in, vals_out, (int)n, 0,
                                             end_bit, stream));
    char *tmp;
    PIB_TRY(ws.alloc(&tmp, tmp_bytes));
    PIB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out, vals_in, vals_out, (int)n, 0,
                                             end_bit, stream));
    return PIB_OK;
}

// stable radix sort of (key, value) pairs on the low `end_bit` bits (used for the pair kernel's work-item order)
int sort_pairs_u32(const uint32_t *keys_in, uint32_t *keys_out, const int32_t *vals_in, int32_t *vals_out,
                   int64_t n, int end_bit, Scratch &ws, cudaStream_t stream)
{
    return radix_sort_pairs<uint32_t>(keys_in, keys_out, vals_in, vals_out, n, end_bit, ws, stream);
}

int build_sorted_points(const double *d_xyz, int64_t n, int tile, int pad_to, Scratch &ws,
                        cudaStream_t stream, SortedPoints *out)
{
    // the bounding box stays on the device (the keys are scaled there); a copy travels to a pinned host slot behind an
    // event, and SortedPoints::wait_box() is only called by the few callers that need the numbers on the host
    double4 *d_box;
    PIB_TRY(bounding_box_device(d_xyz, n, 3, ws, stream, &d_box));

    uint32_t *keys, *keys_s;
    int32_t *vals, *vals_s;
    PIB_TRY(ws.alloc(&keys, n)); PIB_TRY(ws.alloc(&keys_s, n));
    PIB_TRY(ws.alloc(&vals, n)); PIB_TRY(ws.alloc(&vals_s, n));
    const int blocks = (int)((n + 255) / 256);
    curve_key_kernel<<<blocks, 256, 0, stream>>>(d_xyz, n, d_box, keys, vals);
    PIB_CUDA(cudaGetLastError());
    PIB_TRY(radix_sort_pairs(keys, keys_s, vals, vals_s, n, 32, ws, stream));

    SortedPoints sp;
    sp.n = n; sp.tile = tile;
    sp.d_box = d_box;
    static thread_local double *h_box = nullptr;              // pinned, one slot per calling thread
    if (!h_box) PIB_CUDA(cudaMallocHost(reinterpret_cast<void **>(&h_box), 4 * sizeof(double)));
    PIB_CUDA(cudaMemcpyAsync(h_box, d_box, 4 * sizeof(double), cudaMemcpyDeviceToHost, stream));
    PIB_CUDA(cudaEventCreateWithFlags(&sp.box_ready, cudaEventDisableTiming));
    PIB_CUDA(cudaEventRecord(sp.box_ready, stream));
    sp.h_box = h_box;
    sp.n_pad = (n + pad_to - 1) / pad_to * pad_to;
    // the pair kernels read whole A sub-tiles (up to 1024 points) that may start in the last tile: keep that many
    // sentinel rows behind n_pad so the reads stay inside the arrays
    const int64_t n_alloc = sp.n_pad + 1024;
    PIB_TRY(ws.alloc(&sp.xy, n_alloc)); PIB_TRY(ws.alloc(&sp.z, n_alloc));
    PIB_TRY(ws.alloc(&sp.orig, n_alloc));
    gather_soa_kernel<<<(int)((n_alloc + 255) / 256), 256, 0, stream>>>(d_xyz, vals_s, n, n_alloc, sp.xy, sp.z,
                                                                        sp.orig);
    PIB_CUDA(cudaGetLastError());
    const int64_t n_tiles = sp.n_pad / tile;
    PIB_TRY(ws.alloc(&sp.tile_box, n_tiles));
    tile_box_kernel<<<(int)((n_tiles * 32 + 255) / 256), 256, 0, stream>>>(sp.xy, n, tile, n_tiles, sp.tile_box);
    PIB_CUDA(cudaGetLastError());
    *out = sp;
    return PIB_OK;
}

// ---------------------------------------------------------------------------------------------
// Uniform grid
// ---------------------------------------------------------------------------------------------
__global__ void cell_key_kernel(const double *__restrict__ pts, int64_t n, int stride, double x0, double y0,
                                double inv_h, int gx, int gy, uint32_t *__restrict__ keys,
                                int32_t *__restrict__ vals)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cx = clampi((int)floor((pts[i * stride] - x0) * inv_h), 0, gx - 1);
    const int cy = clampi((int)floor((pts[i * stride + 1] - y0) * inv_h), 0, gy - 1);
    keys[i] = (uint32_t)cy * (uint32_t)gx + (uint32_t)cx;
    vals[i] = (int32_t)i;
}

__global__ void gather_grid_kernel(const double *__restrict__ xyz, const int32_t *__restrict__ order, int64_t n,
                                   double *__restrict__ x, double *__restrict__ y, double *__restrict__ z,
                                   int32_t *__restrict__ orig)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t o = order[i];
    x[i] = xyz[3 * o]; y[i] = xyz[3 * o + 1]; z[i] = xyz[3 * o + 2];
    orig[i] = (int32_t)o;
}

// cell_start[c] = first sorted position whose cell key is >= c  (c in [0, n_cells])
__global__ void cell_start_kernel(const uint32_t *__restrict__ sorted_keys, int64_t n, int n_cells,
                                  int32_t *__restrict__ cell_start)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n_cells) return;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (sorted_keys[mid] < (uint32_t)c) lo = mid + 1; else hi = mid;
    }
    cell_start[c] = (int32_t)lo;
}

static int bits_for(uint32_t v)
{
    int b = 1;
    while (b < 32 && (v >> b)) ++b;
    return b;
}

int build_grid_index(const double *d_xyz, int64_t n, double target_per_cell, Scratch &ws,
                     cudaStream_t stream, GridIndex *out)
{
    double box[4];
    PIB_TRY(bounding_box(d_xyz, n, 3, ws, stream, box));
    GridIndex g;
    g.n = n; g.x0 = box[0]; g.y0 = box[1];
    const double w = box[2] - box[0], ht = box[3] - box[1];
    const double gmax = std::min(8192.0, 4.0 * std::sqrt((double)n / target_per_cell) + 1.0);
    double h = std::sqrt(std::max(w, 1e-300) * std::max(ht, 1e-300) * target_per_cell / (double)n);
    h = std::max(h, std::max(w, ht) / gmax);
    if (!(h > 0) || !std::isfinite(h)) h = 1.0;           // all points coincide
    g.h = h; g.inv_h = 1.0 / h;
    g.gx = std::max(1, (int)std::ceil(w / h)); g.gy = std::max(1, (int)std::ceil(ht / h));
    g.gx = std::min(g.gx, 16384); g.gy = std::min(g.gy, 16384);
    const int n_cells = g.gx * g.gy;

    uint32_t *keys, *keys_s;
    int32_t *vals, *vals_s;
    PIB_TRY(ws.alloc(&keys, n)); PIB_TRY(ws.alloc(&keys_s, n));
    PIB_TRY(ws.alloc(&vals, n)); PIB_TRY(ws.alloc(&vals_s, n));
    cell_key_kernel<<<(int)((n + 255) / 256), 256, 0, stream>>>(d_xyz, n, 3, g.x0, g.y0, g.inv_h, g.gx, g.gy, keys, vals);
    PIB_CUDA(cudaGetLastError());
    PIB_TRY(radix_sort_pairs(keys, keys_s, vals, vals_s, n, bits_for((uint32_t)n_cells), ws, stream));
    PIB_TRY(ws.alloc(&g.x, n)); PIB_TRY(ws.alloc(&g.y, n)); PIB_TRY(ws.alloc(&g.z, n)); PIB_TRY(ws.alloc(&g.orig, n));
    gather_grid_kernel<<<(int)((n + 255) / 256), 256, 0, stream>>>(d_xyz, vals_s, n, g.x, g.y, g.z, g.orig);
    PIB_CUDA(cudaGetLastError());
    PIB_TRY(ws.alloc(&g.cell_start, (size_t)n_cells + 1));
    cell_start_kernel<<<(n_cells + 1 + 255) / 256, 256, 0, stream>>>(keys_s, n, n_cells, g.cell_start);
    PIB_CUDA(cudaGetLastError());
    *out = g;
    return PIB_OK;
}

int sort_queries_by_cell(const double *d_unknown, int64_t m, const GridIndex &g, Scratch &ws,
                         cudaStream_t stream, int32_t **order_out)
{
    uint32_t *keys, *keys_s;
    int32_t *vals, *vals_s;
    PIB_TRY(ws.alloc(&keys, m)); PIB_TRY(ws.alloc(&keys_s, m));
    PIB_TRY(ws.alloc(&vals, m)); PIB_TRY(ws.alloc(&vals_s, m));
    cell_key_kernel<<<(int)((m + 255) / 256), 256, 0, stream>>>(d_unknown, m, 2, g.x0, g.y0, g.inv_h, g.gx, g.gy, keys,
                                                                vals);
    PIB_CUDA(cudaGetLastError());
    PIB_TRY(radix_sort_pairs(keys, keys_s, vals, vals_s, m, bits_for((uint32_t)(g.gx * g.gy)), ws, stream));
    *order_out = vals_s;
    return PIB_OK;
}

}  // namespace pib

namespace pib {
int SortedPoints::wait_box()
{
    if (box_ready) {
        PIB_CUDA(cudaEventSynchronize(box_ready));
        cudaEventDestroy(box_ready);
        box_ready = nullptr;
        for (int i = 0; i < 4; ++i) box[i] = h_box[i];
        PIB_CHECK(std::isfinite(box[0]) && std::isfinite(box[1]) && std::isfinite(box[2]) && std::isfinite(box[3]),
                  PIB_ERR_INVALID, "coordinates must be finite");
    }
    return PIB_OK;
}
}  // namespace pib
