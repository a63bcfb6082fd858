// extern "C" surface of libpyinterp_b200.so — see include/pyinterp_b200.h for the contract.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "common.cuh"

namespace pib {

static thread_local std::string g_error;
void set_error(const std::string &msg) { g_error = msg; }

// Keep the stream-ordered pool's memory across synchronisations (the default trims it to zero at every sync, which
// turns each call's scratch allocation into a driver allocation) — but only up to a bound, so that the peak scratch of one
// large call is handed back to the device instead of starving another allocator in the same process (PyTorch's);
// pib_trim() releases everything at once.
static std::mutex g_device_mutex;
static const unsigned long long POOL_KEEP_BYTES = 4ull << 30;

void init_device_pool()
{
    static bool done[64] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
    std::lock_guard<std::mutex> lock(g_device_mutex);
    if (done[dev]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = POOL_KEEP_BYTES;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    done[dev] = true;
}

int sm_count()
{
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    std::lock_guard<std::mutex> lock(g_device_mutex);
    if (!cached[dev]) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) return 148;
        cached[dev] = sms;
    }
    return cached[dev];
}

static std::string g_kernel_name[PIB_T_COUNT];
void set_kernel_name(int which, const char *name)
{
    std::lock_guard<std::mutex> lock(g_device_mutex);
    if (which >= 0 && which < PIB_T_COUNT) g_kernel_name[which] = name;
}

struct Segment { cudaEvent_t a, b; };
static std::vector<Segment> g_segments[PIB_T_COUNT];
static std::mutex g_timing_mutex;

static bool g_timing_keep = false;     // pib_timing_accumulate: records survive the per-call reset
static void timing_clear(int which)
{
    for (auto &s : g_segments[which]) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    g_segments[which].clear();
}
void timing_reset(int which)
{
    std::lock_guard<std::mutex> lock(g_timing_mutex);
    if (g_timing_keep) return;
    for (auto &s : g_segments[which]) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    g_segments[which].clear();
}
void timing_begin(int which, cudaStream_t stream)
{
    std::lock_guard<std::mutex> lock(g_timing_mutex);
    Segment s;
    if (cudaEventCreate(&s.a) != cudaSuccess || cudaEventCreate(&s.b) != cudaSuccess) return;
    cudaEventRecord(s.a, stream);
    g_segments[which].push_back(s);
}
void timing_end(int which, cudaStream_t stream)
{
    std::lock_guard<std::mutex> lock(g_timing_mutex);
    if (!g_segments[which].empty()) cudaEventRecord(g_segments[which].back().b, stream);
}

int variogram_device(const double *d_xyz, int64_t n, const double *edges, const double *lower, int n_lags,
                     const double *W, int n_dirs, int shard_index, int shard_count,
                     double *d_sum_sq, double *d_sum_prod, double *d_sum_val, int64_t *d_count,
                     int64_t *pairs_evaluated, cudaStream_t stream);
int variogram_cloud_device(const double *d_xyz, int64_t n, const double *edges, const double *lower, int n_lags,
                           const double *W, int n_dirs, const int64_t *lag_offsets, double *d_values,
                           cudaStream_t stream);
int variogram_weighted_device(const double *d_xyz, const double *d_weights, int64_t n, const double *edges,
                              const double *lower, int n_lags, const double *W, int n_dirs, double *d_sums,
                              int64_t *d_count, cudaStream_t stream);
int variogram_triangle_device(const double *d_xyz, int64_t n, const double *tri_lags, int n_lags, double cos_t,
                              double sin_t, double tolerance, double *d_sums, int64_t *d_count, cudaStream_t stream);
int variogram_triangle_cloud_device(const double *d_xyz, int64_t n, const double *tri_lags, int n_lags, double cos_t,
                                    double sin_t, double tolerance, const int64_t *lag_offsets, double *d_values,
                                    cudaStream_t stream);
int krige_device(const double *d_known, int64_t n, const double *d_values, int n_sets, const double *params,
                 const double *d_unknown, int64_t m, int model, int mode, double process_mean, int k,
                 double neighbors_range, double direction, double max_tick, int use_all, int allow_lsa,
                 const int32_t *d_exclude, double *d_out, int32_t *d_nbr_idx, uint8_t *d_status, cudaStream_t stream,
                 const GridIndex *prebuilt = nullptr);
int krige_build_index(const double *d_known, int64_t n, Scratch &ws, cudaStream_t stream, GridIndex *out);
int gamma_device(int model, double nugget, double sill, double rang, const double *d_h, int64_t n, double *d_out,
                 cudaStream_t stream);

// 8 independent DFMA chains per thread, 4096 DFMAs each
__global__ void dfma_peak_kernel(double *out, double a, double b, int iters)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = __fma_rn(x0, a, b); x1 = __fma_rn(x1, a, b); x2 = __fma_rn(x2, a, b); x3 = __fma_rn(x3, a, b);
            x4 = __fma_rn(x4, a, b); x5 = __fma_rn(x5, a, b); x6 = __fma_rn(x6, a, b); x7 = __fma_rn(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

}  // namespace pib

using namespace pib;

extern "C" {

const char *pib_last_error(void) { return g_error.c_str(); }
int pib_version(void) { return 100; }
int pib_device_sm_count(void) { return sm_count(); }

int pib_trim(void)
{
    int dev = 0;
    PIB_CUDA(cudaGetDevice(&dev));
    PIB_CUDA(cudaDeviceSynchronize());
    cudaMemPool_t pool;
    PIB_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
    PIB_CUDA(cudaMemPoolTrimTo(pool, 0));
    return PIB_OK;
}

int pib_variogram_device(const double *d_xyz, int64_t n, const double *edges, const double *lower, int n_lags,
                         const double *W, int n_dirs, int shard_index, int shard_count, double *d_sum_sq,
                         double *d_sum_prod, double *d_sum_val, int64_t *d_count, int64_t *pairs_evaluated, void *stream)
{
    PIB_CHECK(d_xyz || n == 0, PIB_ERR_INVALID, "variogram: xyz is NULL");
    PIB_CHECK(d_sum_sq && d_sum_prod && d_sum_val && d_count, PIB_ERR_INVALID, "variogram: output pointer is NULL");
    return variogram_device(d_xyz, n, edges, lower, n_lags, W, n_dirs, shard_index, shard_count, d_sum_sq, d_sum_prod,
                            d_sum_val, d_count, pairs_evaluated, static_cast<cudaStream_t>(stream));
}

int pib_variogram_host(const double *xyz, int64_t n, const double *edges, const double *lower, int n_lags,
                       const double *W, int n_dirs, int shard_index, int shard_count, double *sum_sq, double *sum_prod,
                       double *sum_val, int64_t *count, int64_t *pairs_evaluated)
{
    PIB_CHECK(xyz || n == 0, PIB_ERR_INVALID, "variogram: xyz is NULL");
    PIB_CHECK(sum_sq && sum_prod && sum_val && count, PIB_ERR_INVALID, "variogram: output pointer is NULL");
    PIB_CHECK(n >= 0 && n_lags >= 1, PIB_ERR_INVALID, "variogram: need n >= 0 and at least one lag");
    cudaStream_t stream;
    PIB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    int rc;
    {
        Scratch ws(stream);
        const size_t out_n = (size_t)(n_dirs > 0 ? n_dirs : 1) * n_lags;
        double *d_xyz, *d_sums;
        int64_t *d_count;
        rc = ws.alloc(&d_xyz, (size_t)n * 3);
        if (rc == PIB_OK) rc = ws.alloc(&d_sums, out_n * 3);
        if (rc == PIB_OK) rc = ws.alloc(&d_count, out_n);
        if (rc == PIB_OK && n > 0 &&
            cudaMemcpyAsync(d_xyz, xyz, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, stream) != cudaSuccess) {
            set_error("variogram: H2D copy failed"); rc = PIB_ERR_CUDA;
        }
        if (rc == PIB_OK)
            rc = variogram_device(d_xyz, n, edges, lower, n_lags, W, n_dirs, shard_index, shard_count, d_sums,
                                  d_sums + out_n, d_sums + 2 * out_n, d_count, pairs_evaluated, stream);
        if (rc == PIB_OK) {
            cudaError_t e = cudaMemcpyAsync(sum_sq, d_sums, sizeof(double) * out_n, cudaMemcpyDeviceToHost, stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(sum_prod, d_sums + out_n, sizeof(double) * out_n, cudaMemcpyDeviceToHost, stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(sum_val, d_sums + 2 * out_n, sizeof(double) * out_n, cudaMemcpyDeviceToHost, stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(count, d_count, sizeof(int64_t) * out_n, cudaMemcpyDeviceToHost, stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
            if (e != cudaSuccess) { set_error(std::string("variogram: ") + cudaGetErrorString(e)); rc = PIB_ERR_CUDA; }
        }
    }
    cudaStreamSynchronize(stream);
    cudaStreamDestroy(stream);
    return rc;
}

int pib_variogram_weighted_host(const double *xyz, const double *weights, int64_t n, const double *edges,
                                const double *lower, int n_lags, const double *W, int n_dirs, double *sums,
                                int64_t *count)
{
    PIB_CHECK((xyz && weights) || n == 0, PIB_ERR_INVALID, "weighted variogram: NULL pointer");
    PIB_CHECK(edges && sums && count && n >= 0 && n_lags >= 1, PIB_ERR_INVALID, "weighted variogram: bad arguments");
    cudaStream_t stream;
    PIB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    int rc;
    {
        Scratch ws(stream);
        double *d_xyz, *d_w, *d_sums;
        int64_t *d_count;
        rc = ws.alloc(&d_xyz, (size_t)n * 3);
        if (rc == PIB_OK) rc = ws.alloc(&d_w, (size_t)n);
        if (rc == PIB_OK) rc = ws.alloc(&d_sums, (size_t)4 * n_lags);
        if (rc == PIB_OK) rc = ws.alloc(&d_count, (size_t)n_lags);
        cudaError_t e = cudaSuccess;
        if (rc == PIB_OK && n > 0) {
            e = cudaMemcpyAsync(d_xyz, xyz, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(d_w, weights, sizeof(double) * n, cudaMemcpyHostToDevice, stream);
        }
        if (rc == PIB_OK && e == cudaSuccess)
            rc = variogram_weighted_device(d_xyz, d_w, n, edges, lower, n_lags, W, n_dirs, d_sums, d_count, stream);
        if (rc == PIB_OK && e == cudaSuccess) e = cudaMemcpyAsync(sums, d_sums, sizeof(double) * 4 * n_lags, cudaMemcpyDeviceToHost, stream);
        if (rc == PIB_OK && e == cudaSuccess) e = cudaMemcpyAsync(count, d_count, sizeof(int64_t) * n_lags, cudaMemcpyDeviceToHost, stream);
        if (rc == PIB_OK && e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) { set_error(std::string("weighted variogram: ") + cudaGetErrorString(e)); rc = PIB_ERR_CUDA; }
    }
    cudaStreamSynchronize(stream);
    cudaStreamDestroy(stream);
    return rc;
}

int pib_variogram_triangle_host(const double *xyz, int64_t n, const double *tri_lags, int n_lags, double cos_t,
                                double sin_t, double tolerance, double *sums, int64_t *count)
{
    PIB_CHECK(xyz || n == 0, PIB_ERR_INVALID, "triangle variogram: NULL pointer");
    PIB_CHECK(tri_lags && sums && count && n >= 0 && n_lags >= 1, PIB_ERR_INVALID, "triangle variogram: bad arguments");
    cudaStream_t stream;
    PIB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    int rc;
    {
        Scratch ws(stream);
        double *d_xyz, *d_sums;
        int64_t *d_count;
        rc = ws.alloc(&d_xyz, (size_t)n * 3);
        if (rc == PIB_OK) rc = ws.alloc(&d_sums, (size_t)3 * n_lags);
        if (rc == PIB_OK) rc = ws.alloc(&d_count, (size_t)n_lags);
        cudaError_t e = cudaSuccess;
        if (rc == PIB_OK && n > 0) e = cudaMemcpyAsync(d_xyz, xyz, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, stream);
        if (rc == PIB_OK && e == cudaSuccess)
            rc = variogram_triangle_device(d_xyz, n, tri_lags, n_lags, cos_t, sin_t, tolerance, d_sums, d_count, stream);
        if (rc == PIB_OK && e == cudaSuccess) e = cudaMemcpyAsync(sums, d_sums, sizeof(double) * 3 * n_lags, cudaMemcpyDeviceToHost, stream);
        if (rc == PIB_OK && e == cudaSuccess) e = cudaMemcpyAsync(count, d_count, sizeof(int64_t) * n_lags, cudaMemcpyDeviceToHost, stream);
        if (rc == PIB_OK && e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) { set_error(std::string("triangle variogram: ") + cudaGetErrorString(e)); rc = PIB_ERR_CUDA; }
    }
    cudaStreamSynchronize(stream);
    cudaStreamDestroy(stream);
    return rc;
}

int pib_variogram_cloud_host(const double *xyz, int64_t n, const double *edges, const double *lower, int n_lags,
                             const double *W, int n_dirs, const int64_t *lag_offsets, double *values)
{
    PIB_CHECK((xyz || n == 0) && edges && lag_offsets, PIB_ERR_INVALID, "cloud: NULL pointer");
    PIB_CHECK(n >= 0 && n_lags >= 1, PIB_ERR_INVALID, "cloud: need n >= 0 and at least one lag");
    const int64_t total = lag_offsets[n_lags];
    PIB_CHECK(total >= 0 && (values || total == 0), PIB_ERR_INVALID, "cloud: values is NULL");
    if (n < 2 || total == 0) return PIB_OK;
    cudaStream_t stream;
    PIB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    int rc;
    {
        Scratch ws(stream);
        double *d_xyz, *d_values;
        rc = ws.alloc(&d_xyz, (size_t)n * 3);
        if (rc == PIB_OK) rc = ws.alloc(&d_values, (size_t)total);
        cudaError_t e = cudaSuccess;
        if (rc == PIB_OK) e = cudaMemcpyAsync(d_xyz, xyz, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, stream);
        if (rc == PIB_OK && e == cudaSuccess)
            rc = variogram_cloud_device(d_xyz, n, edges, lower, n_lags, W, n_dirs, lag_offsets, d_values, stream);
        if (rc == PIB_OK && e == cudaSuccess) e = cudaMemcpyAsync(values, d_values, sizeof(double) * total, cudaMemcpyDeviceToHost, stream);
        if (rc == PIB_OK && e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) { set_error(std::string("cloud: ") + cudaGetErrorString(e)); rc = PIB_ERR_CUDA; }
    }
    cudaStreamSynchronize(stream);
    cudaStreamDestroy(stream);
    return rc;
}

int pib_variogram_triangle_cloud_host(const double *xyz, int64_t n, const double *tri_lags, int n_lags, double cos_t,
                                      double sin_t, double tolerance, const int64_t *lag_offsets, double *values)
{
    PIB_CHECK((xyz || n == 0) && tri_lags && lag_offsets, PIB_ERR_INVALID, "triangle cloud: NULL pointer");
    PIB_CHECK(n >= 0 && n_lags >= 1, PIB_ERR_INVALID, "triangle cloud: need n >= 0 and at least one lag");
    const int64_t total = lag_offsets[n_lags];
    PIB_CHECK(total >= 0 && (values || total == 0), PIB_ERR_INVALID, "triangle cloud: values is NULL");
    if (n < 1 || total == 0) return PIB_OK;
    cudaStream_t stream;
    PIB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    int rc;
    {
        Scratch ws(stream);
        double *d_xyz, *d_values;
        rc = ws.alloc(&d_xyz, (size_t)n * 3);
        if (rc == PIB_OK) rc = ws.alloc(&d_values, (size_t)total);
        cudaError_t e = cudaSuccess;
        if (rc == PIB_OK) e = cudaMemcpyAsync(d_xyz, xyz, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, stream);
        if (rc == PIB_OK && e == cudaSuccess)
            rc = variogram_triangle_cloud_device(d_xyz, n, tri_lags, n_lags, cos_t, sin_t, tolerance, lag_offsets, d_values, stream);
        if (rc == PIB_OK && e == cudaSuccess) e = cudaMemcpyAsync(values, d_values, sizeof(double) * total, cudaMemcpyDeviceToHost, stream);
        if (rc == PIB_OK && e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) { set_error(std::string("triangle cloud: ") + cudaGetErrorString(e)); rc = PIB_ERR_CUDA; }
    }
    cudaStreamSynchronize(stream);
    cudaStreamDestroy(stream);
    return rc;
}

int pib_krige_ex_device(const double *d_known, int64_t n, const double *d_values, int n_sets, const double *params,
                        const double *d_unknown, int64_t m, int model, int mode, double process_mean, int k,
                        double neighbors_range, double direction, double max_tick, int use_all, int allow_lsa,
                        double *d_out, int32_t *d_nbr_idx, uint8_t *d_status, void *stream)
{
    PIB_CHECK(d_known && (d_unknown || m == 0) && (d_out || m == 0), PIB_ERR_INVALID, "krige: NULL pointer");
    return krige_device(d_known, n, d_values, n_sets, params, d_unknown, m, model, mode, process_mean, k,
                        neighbors_range, direction, max_tick, use_all, allow_lsa, nullptr, d_out, d_nbr_idx, d_status,
                        static_cast<cudaStream_t>(stream));
}

int pib_krige_device(const double *d_known, int64_t n, const double *d_values, int n_sets, const double *params,
                     const double *d_unknown, int64_t m, int model, int mode, double process_mean, int k,
                     double *d_out, int32_t *d_nbr_idx, uint8_t *d_status, void *stream)
{
    return pib_krige_ex_device(d_known, n, d_values, n_sets, params, d_unknown, m, model, mode, process_mean, k, 0.0,
                               nan(""), 0.0, 0, 0, d_out, d_nbr_idx, d_status, stream);
}

int pib_krige_host(const double *known, int64_t n, const double *values, int n_sets, const double *params,
                   const double *unknown, int64_t m, int model, int mode, double process_mean, int k, double *out,
                   int32_t *nbr_idx, uint8_t *status)
{
    return pib_krige_ex_host(known, n, values, n_sets, params, unknown, m, model, mode, process_mean, k, 0.0, nan(""),
                             0.0, 0, 0, out, nbr_idx, status);
}

// Pinned staging buffers of the chunked host path, kept for the life of the process (one set per calling thread):
// allocating pinned memory costs milliseconds, far more than a chunk of work.
struct HostStage {
    void *ptr = nullptr;
    size_t bytes = 0;
    ~HostStage() { if (ptr) cudaFreeHost(ptr); }
    int reserve(size_t want)
    {
        if (want <= bytes) return PIB_OK;
        if (ptr) { cudaFreeHost(ptr); ptr = nullptr; bytes = 0; }
        PIB_CUDA(cudaMallocHost(&ptr, want));
        bytes = want;
        return PIB_OK;
    }
};

#ifndef PIB_KRIGE_LANES_DEFAULT
#define PIB_KRIGE_LANES_DEFAULT 2
#endif
// Host-buffer kriging: the known points go to the device once, the locations stream through in chunks on TWO lanes
// (stream + pinned staging each): while lane A computes chunk i, lane B's results of chunk i-1 travel back and the
// host refills B's input — host<->device copies overlap the kernels instead of bracketing them
// (reference loop: kriging/point/ordinary.py:205-219; SURVEY.md section 8e asks for per-chunk pipelining).
// one column of the [blocks][count][4] result rows -> [blocks][count]
__global__ void extract_column_kernel(const double *__restrict__ rows, int64_t n_rows, int col, double *__restrict__ out)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n_rows) out[i] = rows[4 * i + col];
}

// out_col < 0: `out` receives the full [zhat, sigma^2, x, y] rows ([blocks][m][4]); 0 / 1: only that column ([blocks][m]) —
// a quarter of the device-to-host traffic and of the host-side copy for callers that keep one column (indicator kriging)
static int krige_host_impl(const double *known, int64_t n, const double *values, int n_sets, const double *params,
                           const double *unknown, int64_t m, int model, int mode, double process_mean, int k,
                           double neighbors_range, double direction, double max_tick, int use_all, int allow_lsa,
                           int out_col, double *out, int32_t *nbr_idx, uint8_t *status)
{
    PIB_CHECK(known && (unknown || m == 0) && (out || m == 0), PIB_ERR_INVALID, "krige: NULL pointer");
    PIB_CHECK(out_col < 2, PIB_ERR_INVALID, "krige: the result column is 0 (prediction) or 1 (variance)");
    const size_t width = out_col < 0 ? 4 : 1;                     // doubles per location and block that travel back
    PIB_CHECK(n >= 1 && m >= 0 && k >= 1 && n_sets >= 1, PIB_ERR_INVALID, "krige: need n >= 1, m >= 0, k >= 1");
    PIB_CHECK(mode == PIB_KRIGE_ORDINARY || mode == PIB_KRIGE_SIMPLE || mode == PIB_KRIGE_BOTH, PIB_ERR_INVALID, "krige: unknown mode");
    if (m == 0) return PIB_OK;
    const int n_modes = mode == PIB_KRIGE_BOTH ? 2 : 1;
    const size_t blocks = (size_t)n_modes * n_sets;               // output blocks of m rows each
    // chunk: ~1M locations, fewer when the per-location output is wide (indicator kriging) or few locations are asked
    int64_t chunk = std::max<int64_t>(65536, ((int64_t)1 << 22) / (int64_t)(blocks * 4));
    // sixteen chunks when there is enough work for them, so that the copies of a call that would fit one chunk (a rank's
    // slice of a sharded run) still overlap its kernels and the un-overlapped head (first upload) and tail (last read-back) stay
    // short: one rank's 1.25M-location slice of BASELINE config 4 takes 117 ms in 4 chunks, 111 ms in 16 (97 ms device-resident)
    chunk = std::min(chunk, std::max<int64_t>(k > 32 ? 65536 : 131072, (m + 15) / 16));
    const char *chunk_env = std::getenv("PIB_KRIGE_CHUNK");
    if (chunk_env && std::atoll(chunk_env) > 0) chunk = std::atoll(chunk_env);
    chunk = std::min(chunk, m);
    // lanes of the pipeline: chunks in flight at once (stream + pinned staging each); PIB_KRIGE_LANES overrides (1 .. 4)
    constexpr int MAX_LANES = 4;
    int n_lanes = m > 2 * chunk ? PIB_KRIGE_LANES_DEFAULT : m > chunk ? 2 : 1;
    { const char *le = std::getenv("PIB_KRIGE_LANES"); if (le && std::atoi(le) >= 1) n_lanes = std::min({std::atoi(le), MAX_LANES, (int)((m + chunk - 1) / chunk)}); }

    static thread_local HostStage stage[MAX_LANES];
    const size_t in_bytes = (size_t)chunk * 2 * sizeof(double);
    const size_t out_bytes = blocks * chunk * width * sizeof(double);
    const size_t idx_bytes = nbr_idx ? (size_t)chunk * k * sizeof(int32_t) : 0;
    const size_t st_bytes = status ? blocks * chunk : 0;
    auto align256 = [](size_t b) { return (b + 255) / 256 * 256; };
    const size_t lane_bytes = align256(in_bytes) + align256(out_bytes) + align256(idx_bytes) + align256(st_bytes);
    for (int b = 0; b < n_lanes; ++b) PIB_TRY(stage[b].reserve(lane_bytes));

    cudaStream_t s[MAX_LANES] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t index_ready = nullptr;
    int rc = PIB_OK;
    for (int b = 0; b < n_lanes && rc == PIB_OK; ++b)
        if (cudaStreamCreateWithFlags(&s[b], cudaStreamNonBlocking) != cudaSuccess) { set_error("krige: stream creation failed"); rc = PIB_ERR_CUDA; }
    if (rc == PIB_OK && cudaEventCreateWithFlags(&index_ready, cudaEventDisableTiming) != cudaSuccess) { set_error("krige: event creation failed"); rc = PIB_ERR_CUDA; }
    if (rc == PIB_OK) {
        Scratch ws(s[0]);                                         // known points, value columns and the grid index: shared by both lanes
        double *d_known = nullptr, *d_values = nullptr;
        GridIndex grid;
        rc = ws.alloc(&d_known, (size_t)n * 3);
        if (rc == PIB_OK && values) rc = ws.alloc(&d_values, (size_t)n_sets * n);
        cudaError_t e = cudaSuccess;
        if (rc == PIB_OK) {
            e = cudaMemcpyAsync(d_known, known, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, s[0]);
            if (e == cudaSuccess && values) e = cudaMemcpyAsync(d_values, values, sizeof(double) * n_sets * n, cudaMemcpyHostToDevice, s[0]);
            if (e != cudaSuccess) { set_error(std::string("krige: ") + cudaGetErrorString(e)); rc = PIB_ERR_CUDA; }
        }
        if (rc == PIB_OK) rc = krige_build_index(d_known, n, ws, s[0], &grid);
        if (rc == PIB_OK && (cudaEventRecord(index_ready, s[0]) != cudaSuccess ||
                             [&] { for (int b = 1; b < n_lanes; ++b) if (cudaStreamWaitEvent(s[b], index_ready, 0) != cudaSuccess) return true; return false; }())) {
            set_error("krige: event failed"); rc = PIB_ERR_CUDA;
        }
        timing_reset(PIB_T_KNN);
        timing_reset(PIB_T_SOLVE);

        struct Lane { Scratch *ws = nullptr; double *d_unk = nullptr, *d_out = nullptr, *d_col = nullptr; int32_t *d_idx = nullptr; uint8_t *d_st = nullptr;
                      int64_t begin = -1, count = 0; };
        Lane lane[MAX_LANES];
        // copy a finished chunk out of the lane's pinned staging into the caller's arrays
        auto drain = [&](int b) -> int {
            Lane &L = lane[b];
            if (L.begin < 0) return PIB_OK;
            if (cudaStreamSynchronize(s[b]) != cudaSuccess) { set_error("krige: chunk failed"); return PIB_ERR_CUDA; }
            char *base = static_cast<char *>(stage[b].ptr);
            const double *h_out = reinterpret_cast<const double *>(base + align256(in_bytes));
            for (size_t blk = 0; blk < blocks; ++blk)
                std::memcpy(out + (blk * (size_t)m + L.begin) * width, h_out + blk * (size_t)L.count * width, sizeof(double) * width * L.count);
            if (nbr_idx)
                std::memcpy(nbr_idx + (size_t)L.begin * k, base + align256(in_bytes) + align256(out_bytes), sizeof(int32_t) * L.count * k);
            if (status) {
                const uint8_t *h_st = reinterpret_cast<const uint8_t *>(base + align256(in_bytes) + align256(out_bytes) + align256(idx_bytes));
                for (size_t blk = 0; blk < blocks; ++blk) std::memcpy(status + blk * (size_t)m + L.begin, h_st + blk * (size_t)L.count, L.count);
            }
            delete L.ws; L.ws = nullptr; L.begin = -1;             // the chunk's device buffers go back to the pool (stream ordered)
            return PIB_OK;
        };
        int64_t begin = 0;
        for (int i = 0; begin < m && rc == PIB_OK; ++i, begin += chunk) {
            const int b = i % n_lanes;
            rc = drain(b);
            if (rc != PIB_OK) break;
            Lane &L = lane[b];
            L.begin = begin; L.count = std::min(chunk, m - begin);
            L.ws = new Scratch(s[b]);
            rc = L.ws->alloc(&L.d_unk, (size_t)L.count * 2);
            if (rc == PIB_OK) rc = L.ws->alloc(&L.d_out, blocks * L.count * 4);
            if (rc == PIB_OK && out_col >= 0) rc = L.ws->alloc(&L.d_col, blocks * L.count);
            if (rc == PIB_OK && nbr_idx) rc = L.ws->alloc(&L.d_idx, (size_t)L.count * k);
            if (rc == PIB_OK && status) rc = L.ws->alloc(&L.d_st, blocks * L.count);
            if (rc != PIB_OK) break;
            char *base = static_cast<char *>(stage[b].ptr);
            std::memcpy(base, unknown + (size_t)begin * 2, sizeof(double) * 2 * L.count);
            e = cudaMemcpyAsync(L.d_unk, base, sizeof(double) * 2 * L.count, cudaMemcpyHostToDevice, s[b]);
            if (e != cudaSuccess) { set_error(std::string("krige: ") + cudaGetErrorString(e)); rc = PIB_ERR_CUDA; break; }
            rc = krige_device(d_known, n, d_values, n_sets, params, L.d_unk, L.count, model, mode, process_mean, k, neighbors_range,
                              direction, max_tick, use_all, allow_lsa, nullptr, L.d_out, L.d_idx, L.d_st, s[b], &grid);
            if (rc != PIB_OK) break;
            if (out_col >= 0) {
                const int64_t rows = (int64_t)blocks * L.count;
                extract_column_kernel<<<(int)((rows + 255) / 256), 256, 0, s[b]>>>(L.d_out, rows, out_col, L.d_col);
                e = cudaGetLastError();
                if (e == cudaSuccess)
                    e = cudaMemcpyAsync(base + align256(in_bytes), L.d_col, sizeof(double) * rows, cudaMemcpyDeviceToHost, s[b]);
            } else
            e = cudaMemcpyAsync(base + align256(in_bytes), L.d_out, sizeof(double) * 4 * blocks * L.count, cudaMemcpyDeviceToHost, s[b]);
            if (e == cudaSuccess && nbr_idx)
                e = cudaMemcpyAsync(base + align256(in_bytes) + align256(out_bytes), L.d_idx, sizeof(int32_t) * L.count * k, cudaMemcpyDeviceToHost, s[b]);
            if (e == cudaSuccess && status)
                e = cudaMemcpyAsync(base + align256(in_bytes) + align256(out_bytes) + align256(idx_bytes), L.d_st, blocks * L.count,
                                    cudaMemcpyDeviceToHost, s[b]);
            if (e != cudaSuccess) { set_error(std::string("krige: ") + cudaGetErrorString(e)); rc = PIB_ERR_CUDA; break; }
        }
        for (int b = 0; b < n_lanes; ++b) {
            const int r2 = drain(b);
            if (rc == PIB_OK) rc = r2;
            if (lane[b].ws) { cudaStreamSynchronize(s[b]); delete lane[b].ws; lane[b].ws = nullptr; }
        }
        for (int b = 0; b < n_lanes; ++b) if (s[b]) cudaStreamSynchronize(s[b]);
    }
    if (index_ready) cudaEventDestroy(index_ready);
    for (int b = 0; b < MAX_LANES; ++b)
        if (s[b]) { cudaStreamSynchronize(s[b]); cudaStreamDestroy(s[b]); }
    return rc;
}

int pib_krige_ex_host(const double *known, int64_t n, const double *values, int n_sets, const double *params,
                      const double *unknown, int64_t m, int model, int mode, double process_mean, int k,
                      double neighbors_range, double direction, double max_tick, int use_all, int allow_lsa,
                      double *out, int32_t *nbr_idx, uint8_t *status)
{
    return krige_host_impl(known, n, values, n_sets, params, unknown, m, model, mode, process_mean, k, neighbors_range, direction,
                           max_tick, use_all, allow_lsa, -1, out, nbr_idx, status);
}

int pib_krige_column_host(const double *known, int64_t n, const double *values, int n_sets, const double *params,
                          const double *unknown, int64_t m, int model, int mode, double process_mean, int k,
                          double neighbors_range, double direction, double max_tick, int use_all, int allow_lsa,
                          int column, double *out_column, uint8_t *status)
{
    PIB_CHECK(column == 0 || column == 1, PIB_ERR_INVALID, "krige: the result column is 0 (prediction) or 1 (variance)");
    return krige_host_impl(known, n, values, n_sets, params, unknown, m, model, mode, process_mean, k, neighbors_range, direction,
                           max_tick, use_all, allow_lsa, column, out_column, nullptr, status);
}

// leave-one-out: location i = known row i, kriged from the other rows
__global__ void loo_setup_kernel(const double *__restrict__ known, int64_t n, double *__restrict__ unknown,
                                 int32_t *__restrict__ exclude)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    unknown[2 * i] = known[3 * i]; unknown[2 * i + 1] = known[3 * i + 1];
    exclude[i] = (int32_t)i;
}

int pib_krige_loo_host(const double *known, int64_t n, const double *params, int model, int mode, double process_mean,
                       int k, double neighbors_range, double direction, double max_tick, int use_all, int allow_lsa,
                       double *out, uint8_t *status)
{
    PIB_CHECK(known && out && params, PIB_ERR_INVALID, "krige_loo: NULL pointer");
    PIB_CHECK(n >= 2 && k >= 1, PIB_ERR_INVALID, "krige_loo: need n >= 2, k >= 1");
    cudaStream_t stream;
    PIB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    int rc;
    {
        Scratch ws(stream);
        double *d_known, *d_unknown, *d_out;
        int32_t *d_exclude;
        uint8_t *d_status = nullptr;
        rc = ws.alloc(&d_known, (size_t)n * 3);
        if (rc == PIB_OK) rc = ws.alloc(&d_unknown, (size_t)n * 2);
        if (rc == PIB_OK) rc = ws.alloc(&d_exclude, (size_t)n);
        if (rc == PIB_OK) rc = ws.alloc(&d_out, (size_t)n * 4);
        if (rc == PIB_OK && status) rc = ws.alloc(&d_status, (size_t)n);
        cudaError_t e = cudaSuccess;
        if (rc == PIB_OK) {
            e = cudaMemcpyAsync(d_known, known, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, stream);
            if (e == cudaSuccess) {
                loo_setup_kernel<<<(int)((n + 255) / 256), 256, 0, stream>>>(d_known, n, d_unknown, d_exclude);
                e = cudaGetLastError();
            }
            if (e != cudaSuccess) { set_error(std::string("krige_loo: ") + cudaGetErrorString(e)); rc = PIB_ERR_CUDA; }
        }
        if (rc == PIB_OK)
            rc = krige_device(d_known, n, nullptr, 1, params, d_unknown, n, model, mode, process_mean, k, neighbors_range,
                              direction, max_tick, use_all, allow_lsa, d_exclude, d_out, nullptr, d_status, stream);
        if (rc == PIB_OK) {
            e = cudaMemcpyAsync(out, d_out, sizeof(double) * 4 * n, cudaMemcpyDeviceToHost, stream);
            if (e == cudaSuccess && status) e = cudaMemcpyAsync(status, d_status, (size_t)n, cudaMemcpyDeviceToHost, stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
            if (e != cudaSuccess) { set_error(std::string("krige_loo: ") + cudaGetErrorString(e)); rc = PIB_ERR_CUDA; }
        }
    }
    cudaStreamSynchronize(stream);
    cudaStreamDestroy(stream);
    return rc;
}

int pib_gamma_host(int model, double nugget, double sill, double rang, const double *h, int64_t n, double *out)
{
    PIB_CHECK(model >= 0 && model <= 6, PIB_ERR_INVALID, "gamma: unknown model id");
    if (n <= 0) return PIB_OK;
    double *d_h, *d_o;
    PIB_CUDA(cudaMalloc(&d_h, sizeof(double) * n));
    PIB_CUDA(cudaMalloc(&d_o, sizeof(double) * n));
    int rc = PIB_OK;
    if (cudaMemcpy(d_h, h, sizeof(double) * n, cudaMemcpyHostToDevice) != cudaSuccess) rc = PIB_ERR_CUDA;
    if (rc == PIB_OK) rc = gamma_device(model, nugget, sill, rang, d_h, n, d_o, 0);
    if (rc == PIB_OK && cudaMemcpy(out, d_o, sizeof(double) * n, cudaMemcpyDeviceToHost) != cudaSuccess) rc = PIB_ERR_CUDA;
    cudaFree(d_h); cudaFree(d_o);
    if (rc == PIB_ERR_CUDA) set_error("gamma: CUDA copy failed");
    return rc;
}

int pib_last_kernel_ms(int which, double *ms_out, int *launches_out)
{
    PIB_CHECK(which >= 0 && which < PIB_T_COUNT && ms_out, PIB_ERR_INVALID, "last_kernel_ms: bad argument");
    std::lock_guard<std::mutex> lock(g_timing_mutex);
    double total = 0.0;
    for (auto &s : g_segments[which]) {
        PIB_CUDA(cudaEventSynchronize(s.b));
        float ms = 0;
        PIB_CUDA(cudaEventElapsedTime(&ms, s.a, s.b));
        total += ms;
    }
    *ms_out = total;
    if (launches_out) *launches_out = (int)g_segments[which].size();
    return PIB_OK;
}

int pib_timing_accumulate(int on)
{
    std::lock_guard<std::mutex> lock(g_timing_mutex);
    for (int w = 0; w < PIB_T_COUNT; ++w) timing_clear(w);
    g_timing_keep = on != 0;
    return PIB_OK;
}

int pib_last_kernel_name(int which, char *buf, int len)
{
    PIB_CHECK(which >= 0 && which < PIB_T_COUNT && buf && len > 0, PIB_ERR_INVALID, "last_kernel_name: bad argument");
    std::lock_guard<std::mutex> lock(g_device_mutex);
    std::snprintf(buf, (size_t)len, "%s", g_kernel_name[which].c_str());
    return PIB_OK;
}

int pib_fp64_peak_tflops(double *tflops_out)
{
    PIB_CHECK(tflops_out, PIB_ERR_INVALID, "fp64 peak: NULL output");
    const int blocks = sm_count() * 8, threads = 256, iters = 2048;
    double *d_out;
    PIB_CUDA(cudaMalloc(&d_out, sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1;
    PIB_CUDA(cudaEventCreate(&e0)); PIB_CUDA(cudaEventCreate(&e1));
    double best_ms = 1e30;
    for (int rep = 0; rep < 6; ++rep) {
        PIB_CUDA(cudaEventRecord(e0, 0));
        dfma_peak_kernel<<<blocks, threads>>>(d_out, 1.0000001, 1e-9, iters);
        PIB_CUDA(cudaEventRecord(e1, 0));
        PIB_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        PIB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best_ms) best_ms = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_out);
    const double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
    *tflops_out = flops / (best_ms * 1e-3) / 1e12;
    return PIB_OK;
}

}  // extern "C"
