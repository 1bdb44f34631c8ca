// common.cuh -- shared declarations of the pgeof B200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/pgeof_b200.h"

namespace pgeof {

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern thread_local uint64_t g_launches;   // kernels launched by this thread (pgeof_launch_count)

#define PGEOF_CUDA(expr)                                                                          \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            ::pgeof::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return _e == cudaErrorMemoryAllocation ? PGEOF_ENOMEM : PGEOF_ECUDA;                  \
        }                                                                                         \
    } while (0)

#define PGEOF_TRY(expr)            \
    do {                           \
        int _s = (expr);           \
        if (_s != PGEOF_OK) return _s; \
    } while (0)

#define PGEOF_LAUNCH_CHECK()                 \
    do {                                     \
        ++::pgeof::g_launches;               \
        PGEOF_CUDA(cudaGetLastError());      \
    } while (0)

// Brackets a kernel launch with CUDA events when profiling is on (pgeof_profile_enable).
struct KernelTimer {
    int slot = -1;
    cudaStream_t stream = nullptr;
    cudaEvent_t start = nullptr;
    KernelTimer(const char* name, cudaStream_t s);
    ~KernelTimer();
};

// Device scratch from the library's own caching arena (capi.cu): blocks are cudaMalloc'ed once, kept on a per-device
// free list and handed out again by size.  A block released on stream A and reused on stream B waits for the event
// recorded at its release, so reuse is stream ordered.  (The driver's cudaMallocAsync pool was measured to stall
// 50-700 ms at random when 2 GB results and 40-160 MB temporaries alternate: profiles/r1c_summary.md.)
struct DeviceBuffer {
    void* ptr = nullptr;
    cudaStream_t stream = nullptr;
    size_t bytes = 0;         // capacity of the arena block
    int device = -1;
    void* event = nullptr;    // cudaEvent_t owned by the block
    DeviceBuffer() = default;
    DeviceBuffer(const DeviceBuffer&) = delete;
    DeviceBuffer& operator=(const DeviceBuffer&) = delete;
    ~DeviceBuffer() { release(); }
    int alloc(size_t bytes, cudaStream_t s);
    void release();
    template <typename T> T* as() const { return reinterpret_cast<T*>(ptr); }
};

// ---------------------------------------------------------------------------
// Uniform grid index (replaces the nanoflann KD-tree, nn_search.hpp:39).
// Cells are ordered row-major (z, y, x): the cells of one x-run are contiguous in
// `pts`, so the candidates of a query are a handful of contiguous float4 spans.
// ---------------------------------------------------------------------------
struct GridView {
    float lo[3];          // bbox minimum
    float h;              // cell edge along y and z
    float inv_h;          // fl(1/h)
    float hx;             // cell edge along x = h / xf (finer: x-runs are trimmed at this granularity)
    float inv_hx;         // fl(1/hx)
    int xf;               // x refinement factor (>= 1)
    float slack;          // bound on |nominal cell boundary - true assignment boundary|
    int n[3];             // cells per axis
    uint32_t n_pts;
    // A grid built for queries that live in a small part of the cloud indexes only the points inside [clip_lo, clip_hi]
    // (the query bounding box dilated by rmax_safe, cut to the cloud's bounding box): a search is exact on it as long as
    // its ball radius stays <= rmax_safe.  The full grid has clip = +-FLT_MAX and rmax_safe = +inf.
    float clip_lo[3], clip_hi[3];
    float rmax_safe;
    const uint32_t* cell_start;   // [n_cells + 1]
    const float4* pts;            // [n_pts] (x, y, z, bits(original index)), sorted by cell
};

struct Grid {
    GridView view{};
    DeviceBuffer cell_start, pts;
    size_t n_cells = 0;
};

// Builds the grid over `xyz` (device, (n,3) dense).  `cell_edge` <= 0 picks the (y, z) edge
// from `target_occupancy` (mean points per h^3 of the bounding box).  Cells are `x_refine`
// times finer along x (the contiguous axis), which costs nothing at query time: a query
// still reads one contiguous span per (y, z) row, only trimmed more tightly.
struct GridClip {
    float lo[3], hi[3];   // index only the points inside this box (already cut to the cloud's bounding box)
    float cell_edge;      // first guess of the (y, z) cell edge, from the density of the whole cloud
    float rmax_safe;      // ball radius up to which a search of a query inside the undilated box sees every point it needs
};
int grid_build(const float* xyz, size_t n, float cell_edge, float target_occupancy, int x_refine, cudaStream_t stream, Grid* out,
               const GridClip* clip = nullptr);

// bounding boxes of one or two (n,3) clouds with a single host synchronisation; false if a cloud holds non-finite values
int bbox_host(const float* a, size_t na, const float* b, size_t nb, cudaStream_t stream, float (&alo)[3], float (&ahi)[3],
              float (&blo)[3], float (&bhi)[3]);

// Sorts `query` (device, (nq,3)) by the cell of `grid` it falls in (clamped) and
// returns float4 (x, y, z, bits(original index)) records in that order.
int grid_sort_queries(const Grid& grid, const float* query, size_t nq, cudaStream_t stream, DeviceBuffer* out);

// ---------------------------------------------------------------------------
// search / features drivers (device pointers, asynchronous on `stream`)
// ---------------------------------------------------------------------------
enum SearchMode { SEARCH_KNN = 0, SEARCH_RADIUS = 1, SEARCH_RADIUS_COUNT = 2, SEARCH_RADIUS_CSR = 3 };

int search_run(SearchMode mode, const float* data, size_t n_data, const float* query, size_t n_query, uint32_t k,
               float radius, void* indices, float* sqr_dist, uint32_t* nn_ptr, cudaStream_t stream);

// CSR row offsets: uint32 as in the reference (pgeof.hpp:78-79) or uint64 (extension beyond 2^32-1 neighbours)
struct RowPtr {
    const uint32_t* p32 = nullptr;
    const unsigned long long* p64 = nullptr;
    RowPtr(const uint32_t* p) : p32(p) {}
    RowPtr(const unsigned long long* p) : p64(p) {}
    RowPtr(const uint32_t* a, const unsigned long long* b) : p32(a), p64(b) {}
};

int features_run(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz, RowPtr nn_ptr, size_t n_rows,
                 uint32_t k_min, int eig_order, float* out, cudaStream_t stream, const uint32_t* out_rows = nullptr);
// knn_search(xyz, xyz, knn) + compute_features without materialising the neighbour lists (fused knn_features extension);
// *done = 0 if the fused path declined (the caller then runs the two kernels)
int knn_features_fused_run(const float* xyz, size_t n, uint32_t knn, uint32_t k_min, int eig_order, float* features, cudaStream_t stream, int* done);
int features_multiscale_run(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz, RowPtr nn_ptr,
                            size_t n_rows, const uint32_t* k_scales_host, size_t n_scales, int eig_order, float* out,
                            cudaStream_t stream);
int features_optimal_run(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz, RowPtr nn_ptr,
                         size_t n_rows, uint32_t k_min, uint32_t k_step, uint32_t k_min_search, int eig_order, float* out,
                         cudaStream_t stream);
int selected_run_f32(const float* xyz, size_t n, float radius, uint32_t max_knn, const int32_t* ids_host, size_t n_ids,
                     int eig_order, float* out, cudaStream_t stream);
int selected_run_f64(const double* xyz, size_t n, double radius, uint32_t max_knn, const int32_t* ids_host, size_t n_ids,
                     int eig_order, double* out, cudaStream_t stream);

// Window of the caller's nn array that is resident behind the `nn` pointer of the next feature call of this thread: rows may
// only address nn[lo, nnz).  Used by the chunked host pipeline, which passes `slice - lo` as nn; 0 otherwise.
extern thread_local unsigned long long g_nn_window_lo;

// per-block bounding boxes of an (n,3) cloud: out[b*6 + {0..2}] = min, out[b*6 + {3..5}] = max
int bbox_partials(const float* xyz, size_t n, DeviceBuffer* out, int* n_partials, cudaStream_t stream);

// exclusive scan of `n` uint32 counts in place; writes the grand total to data[n]
int exclusive_scan_u32(uint32_t* data, size_t n, cudaStream_t stream);

// spatial query shards (shard.cu)
int slab_plan_run(const float* xyz, size_t n, int rank, int world, int axis, pgeof_slab_plan* out, cudaStream_t stream);
int slab_fill_run(const float* xyz, size_t n, const pgeof_slab_plan* plan, long long* rows, float* query, cudaStream_t stream);

// multiprocessors of the current device (cached per device; 148 on B200)
int sm_count();

// device error flag helpers (EINDEX etc.)
int device_flag_check(const int* d_flag, cudaStream_t stream, const char* what);

}  // namespace pgeof
