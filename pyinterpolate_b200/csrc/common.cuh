// Shared helpers for libpyinterp_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/pyinterp_b200.h"

namespace pib {

void set_error(const std::string &msg);

#define PIB_CUDA(expr)                                                                       \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            char _buf[512];                                                                  \
            snprintf(_buf, sizeof(_buf), "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,       \
                     cudaGetErrorString(_e));                                                \
            pib::set_error(_buf);                                                            \
            return PIB_ERR_CUDA;                                                             \
        }                                                                                    \
    } while (0)

#define PIB_CHECK(cond, code, msg)                                                           \
    do {                                                                                     \
        if (!(cond)) {                                                                       \
            pib::set_error(msg);                                                             \
            return (code);                                                                   \
        }                                                                                    \
    } while (0)

#define PIB_TRY(expr)                                                                        \
    do {                                                                                     \
        int _r = (expr);                                                                     \
        if (_r != PIB_OK) return _r;                                                         \
    } while (0)

// Stream-ordered scratch allocation (cudaMallocAsync pool); freed in the destructor on the same
// stream, so a call never synchronises just to release its workspace.
void init_device_pool();

struct Scratch {
    cudaStream_t stream;
    void *ptrs[128];
    int n = 0;
    explicit Scratch(cudaStream_t s) : stream(s) { init_device_pool(); }
    ~Scratch() {
        for (int i = 0; i < n; ++i) cudaFreeAsync(ptrs[i], stream);
    }
    template <typename T>
    int alloc(T **out, size_t count) {
        void *p = nullptr;
        size_t bytes = (count ? count : 1) * sizeof(T);
        PIB_CUDA(cudaMallocAsync(&p, bytes, stream));
        if (n >= 128) { cudaFreeAsync(p, stream); set_error("scratch table full"); return PIB_ERR_CUDA; }
        ptrs[n++] = p;
        *out = static_cast<T *>(p);
        return PIB_OK;
    }
};

int sm_count();

// Per-kernel device timing for bench.py's roofline: CUDA events recorded on the launching stream
// around the hot kernels of the LAST library call (several segments when a call launches the kernel
// more than once).  Read back with pib_last_kernel_ms().
enum { PIB_T_PAIR = 0, PIB_T_KNN = 1, PIB_T_SOLVE = 2, PIB_T_COUNT = 3 };
void timing_reset(int which);
void timing_begin(int which, cudaStream_t stream);
void timing_end(int which, cudaStream_t stream);
// name of the kernel variant the last call launched for a timing slot (bench.py reports it beside the roofline)
void set_kernel_name(int which, const char *name);

// ---- spatial preprocessing (spatial.cu) --------------------------------------------------
// Morton-ordered structure-of-arrays copy of the points, padded to a multiple of `pad_to`
// with far-away sentinels, plus one bounding box per `tile` consecutive points.
struct SortedPoints {
    int64_t n = 0, n_pad = 0;
    double2 *xy = nullptr;                             // [n_pad] (x, y) interleaved: one 16-byte load per point
    double *z = nullptr;                               // [n_pad]
    int32_t *orig = nullptr;                           // [n_pad] original row, -1 for padding
    double4 *tile_box = nullptr;                       // [n_pad / tile]  (xmin, ymin, xmax, ymax)
    int tile = 0;
    double4 *d_box = nullptr;                          // xmin, ymin, xmax, ymax of all points (device)
    double box[4] = {0, 0, 0, 0};                      // the same on the host — valid after wait_box()
    double *h_box = nullptr;
    cudaEvent_t box_ready = nullptr;
    int wait_box();                                    // waits for the host copy of the box; checks that it is finite
};
int build_sorted_points(const double *d_xyz, int64_t n, int tile, int pad_to, Scratch &ws,
                        cudaStream_t stream, SortedPoints *out);

int sort_pairs_u32(const uint32_t *keys_in, uint32_t *keys_out, const int32_t *vals_in, int32_t *vals_out,
                   int64_t n, int end_bit, Scratch &ws, cudaStream_t stream);

// Uniform grid over the known points for neighbour search.
struct GridIndex {
    int64_t n = 0;
    double x0 = 0, y0 = 0, inv_h = 0, h = 0;
    int gx = 0, gy = 0;
    double *x = nullptr, *y = nullptr, *z = nullptr;   // [n] cell-sorted
    int32_t *orig = nullptr;                           // [n] original row
    int32_t *cell_start = nullptr;                     // [gx*gy + 1]
};
int build_grid_index(const double *d_xyz, int64_t n, double target_per_cell, Scratch &ws,
                     cudaStream_t stream, GridIndex *out);
// order[q] = query ids sorted by grid cell (locality only; results are scattered back)
int sort_queries_by_cell(const double *d_unknown, int64_t m, const GridIndex &g, Scratch &ws,
                         cudaStream_t stream, int32_t **order_out);

static __device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Bounds checks of the hot kernels' computed indices, compiled in with -DPIB_BOUNDS_CHECK (scripts/build_dbg.sh builds
// build/libpib_dbg.so; run the GPU tests against it with PIB_LIB_PATH).  compute-sanitizer is closed on the GPU pool, so this
// is the memory-safety evidence: a violated check prints the site and traps, which fails the call with a CUDA error.
#ifdef PIB_BOUNDS_CHECK
#define PIB_DBG_ASSERT(cond)                                                                          \
    do {                                                                                              \
        if (!(cond)) {                                                                                \
            printf("PIB_BOUNDS_CHECK failed: %s  (%s:%d, block %d thread %d)\n", #cond, __FILE__, __LINE__, (int)blockIdx.x, \
                   (int)threadIdx.x);                                                                 \
            __trap();                                                                                 \
        }                                                                                             \
    } while (0)
#else
#define PIB_DBG_ASSERT(cond) do { } while (0)
#endif

}  // namespace pib
