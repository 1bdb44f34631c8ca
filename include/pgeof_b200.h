/* pgeof_b200.h -- C ABI of the B200-native pgeof hot path (libpgeof_b200.so).
 *
 * This is the drop-in boundary: one entry point per function the reference's
 * nanobind module binds (src/pgeof_ext.cpp:34-177 of drprojects/point_geometric_features),
 * taking plain pointers and sizes.  The binding TU (point_geometric_features_b200/
 * binding/pgeof_ext.cpp) and any other FFI (ctypes, cgo, JNI ...) call only these.
 *
 * Two flavours per entry point:
 *   pgeof_<fn>          HOST buffers (numpy callers).  The library stages H2D / D2H
 *                       itself on an internal stream of the current device.
 *   pgeof_<fn>_dev      DEVICE buffers (DLPack / torch CUDA tensors) on `stream`
 *                       (a cudaStream_t cast to void*; NULL = legacy default stream).
 *                       Work is enqueued on `stream`, but the calls are NOT fully asynchronous:
 *                       every search synchronises the stream once or twice (bounding box -> grid
 *                       dimensions on the host, count of the queries handed to a second pass) and every
 *                       feature call once at its end (the bad-index flag that becomes PGEOF_EINDEX), so
 *                       results are complete and errors reported when a call returns; do not capture
 *                       them in a CUDA graph ("Host synchronisation").
 *
 * Conventions
 *   - all arrays are dense C-contiguous; xyz-like arrays are (n,3) row-major.
 *   - outputs are caller-allocated (ownership stays with the caller's array object;
 *     the reference hands numpy a capsule-owned buffer, nn_search.hpp:41-45).
 *   - return value: PGEOF_OK or a negative pgeof_status; the message is available
 *     from pgeof_last_error() (thread-local).  PGEOF_EINVAL maps to the reference's
 *     std::invalid_argument sites (-> Python ValueError).
 *   - eig_order: PGEOF_EIG_LITERAL reproduces include/pca.hpp:79-89 read literally
 *     (Eigen returns eigenvalues in increasing order and the snapshot does not
 *     re-sort); PGEOF_EIG_DOCUMENTED is the decreasing order the docs assume.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point
 *     returns PGEOF_ECUDA.
 */
#ifndef PGEOF_B200_H
#define PGEOF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PGEOF_API __attribute__((visibility("default")))
#else
#define PGEOF_API
#endif

typedef enum pgeof_status {
    PGEOF_OK = 0,
    PGEOF_EINVAL = -1,  /* bad argument (reference: std::invalid_argument -> ValueError) */
    PGEOF_ECUDA = -2,   /* CUDA runtime failure / no device                              */
    PGEOF_ENOMEM = -3,  /* device or pinned-host allocation failed                       */
    PGEOF_EINDEX = -4   /* nn holds an index >= n_xyz (reference: undefined behaviour)   */
} pgeof_status;

enum { PGEOF_EIG_LITERAL = 0, PGEOF_EIG_DOCUMENTED = 1 };

/* EFeatureID, include/pca.hpp:47-63 (values are the output column order). */
enum {
    PGEOF_LINEARITY = 0, PGEOF_PLANARITY = 1, PGEOF_SCATTERING = 2, PGEOF_VERTICALITY_PGEOF = 3,
    PGEOF_NORMAL_X = 4, PGEOF_NORMAL_Y = 5, PGEOF_NORMAL_Z = 6, PGEOF_LENGTH = 7, PGEOF_SURFACE = 8,
    PGEOF_VOLUME = 9, PGEOF_CURVATURE = 10, PGEOF_K_OPTIMAL = 11, PGEOF_VERTICALITY = 12,
    PGEOF_EIGENTROPY = 13
};

/* ---- library / device management ---------------------------------------- */
PGEOF_API int pgeof_abi_version(void);
PGEOF_API const char* pgeof_last_error(void);
PGEOF_API int pgeof_device_count(void);                 /* 0 when no usable GPU        */
PGEOF_API int pgeof_set_device(int device);             /* device used by host entry points */
PGEOF_API int pgeof_get_device(void);
/* kernels launched by this library on the calling thread since the last reset */
PGEOF_API uint64_t pgeof_launch_count(void);
PGEOF_API void pgeof_reset_launch_count(void);
/* release cached device workspaces and pinned staging of the current device */
PGEOF_API int pgeof_trim(void);

/* Per-kernel device timing (CUDA events on the launching stream) for roofline reports.
 * Off by default; when on, every hot kernel launch is bracketed by two events.
 * pgeof_profile_read sums the launches of kernel `name` ("knn_search", "radius_search",
 * "features", "multiscale", "optimal", "selected", "grid_build") since the last reset;
 * it synchronises those events.  Returns PGEOF_EINVAL for an unknown name. */
PGEOF_API void pgeof_profile_enable(int on);
PGEOF_API void pgeof_profile_reset(void);
PGEOF_API int pgeof_profile_read(const char* name, double* total_ms, uint64_t* launches);

/* Pinned host memory from the library's caching pool.  The binding backs numpy
 * result arrays with it so the D2H copy of a result is a single DMA. */
PGEOF_API void* pgeof_host_alloc(size_t bytes);
PGEOF_API void pgeof_host_free(void* p);

/* ---- neighbour search (replaces include/nn_search.hpp) ------------------- */
/* nanoflann_knn_search<float>, nn_search.hpp:31-67 / pgeof_ext.cpp:118.
 * indices (n_query,knn) uint32, sqr_dist (n_query,knn) float32, rows ascending by
 * (d2, index).  knn > n_data -> PGEOF_EINVAL (nn_search.hpp:37). */
PGEOF_API int pgeof_knn_search(const float* data, size_t n_data, const float* query, size_t n_query,
                               uint32_t knn, uint32_t* indices, float* sqr_dist);
PGEOF_API int pgeof_knn_search_dev(const float* data, size_t n_data, const float* query, size_t n_query,
                                   uint32_t knn, uint32_t* indices, float* sqr_dist, void* stream);

/* nanoflann_radius_search<float>, nn_search.hpp:85-132 / pgeof_ext.cpp:130.
 * Keeps the max_knn nearest with d2 < fl(r*r) (strict); pads indices with -1 and
 * distances with 0.  max_knn > n_data -> PGEOF_EINVAL (nn_search.hpp:92-95). */
PGEOF_API int pgeof_radius_search(const float* data, size_t n_data, const float* query, size_t n_query,
                                  float search_radius, uint32_t max_knn, int32_t* indices, float* sqr_dist);
PGEOF_API int pgeof_radius_search_dev(const float* data, size_t n_data, const float* query, size_t n_query,
                                      float search_radius, uint32_t max_knn, int32_t* indices, float* sqr_dist,
                                      void* stream);

/* Extension: radius search emitting CSR directly -- what the README glue
 * (README.md:157-163) builds from the padded result.  nn_ptr (n_query+1) uint32 is written; *nnz
 * receives the total.  Call once with nn == NULL to size nn (nn_ptr and *nnz are produced), then
 * again with the same arguments and nn to fill it: the pair runs ONE search (the padded table of
 * the first call stays parked in library scratch, per host thread, until the second call, the next
 * first call or pgeof_trim).  The first call synchronises `stream` to read *nnz back. */
PGEOF_API int pgeof_radius_search_csr(const float* data, size_t n_data, const float* query, size_t n_query,
                                      float search_radius, uint32_t max_knn, uint32_t* nn_ptr, uint32_t* nn,
                                      uint64_t* nnz);
PGEOF_API int pgeof_radius_search_csr_dev(const float* data, size_t n_data, const float* query, size_t n_query,
                                          float search_radius, uint32_t max_knn, uint32_t* nn_ptr, uint32_t* nn,
                                          uint64_t* nnz, void* stream);

/* Extension: kNN emitting CSR directly -- nn = the (n_query, knn) table of
 * nanoflann_knn_search (nn_search.hpp:31-67) flattened, nn_ptr[i] = i * knn with ptr_bits = 32
 * (uint32_t*, the reference's dtype) or 64 (uint64_t*: the README's "known limitation" of
 * 2^32-1 neighbours per CSR, README.md:187-195, does not apply).  No squared distances. */
PGEOF_API int pgeof_knn_search_csr(const float* data, size_t n_data, const float* query, size_t n_query,
                                   uint32_t knn, uint32_t* nn, void* nn_ptr, int ptr_bits);
PGEOF_API int pgeof_knn_search_csr_dev(const float* data, size_t n_data, const float* query, size_t n_query,
                                       uint32_t knn, uint32_t* nn, void* nn_ptr, int ptr_bits, void* stream);

/* ---- neighbourhood-PCA features (replaces include/pca.hpp + pgeof.hpp) --- */
/* Every CSR feature function also exists with uint64 row offsets (_p64: extension; the reference
 * binds uint32 only, pgeof.hpp:78-79, which caps one CSR at 2^32-1 neighbours). */
/* compute_geometric_features<float,11>, pgeof.hpp:75-117 / pgeof_ext.cpp:34.
 * out (n_rows,11); rows shorter than k_min stay 0.  k_min < 1 -> PGEOF_EINVAL. */
PGEOF_API int pgeof_compute_features(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz,
                                     const uint32_t* nn_ptr, size_t n_rows, uint32_t k_min, int eig_order,
                                     float* out);
PGEOF_API int pgeof_compute_features_dev(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz,
                                         const uint32_t* nn_ptr, size_t n_rows, uint32_t k_min, int eig_order,
                                         float* out, void* stream);

/* compute_geometric_features_multiscale<float,11>, pgeof.hpp:159-211 / pgeof_ext.cpp:59.
 * k_scales is a HOST array in both flavours; out (n_rows,n_scales,11).
 * k_scales not non-decreasing from 1 -> PGEOF_EINVAL (pgeof.hpp:123-132,165). */
PGEOF_API int pgeof_compute_features_multiscale(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz,
                                                const uint32_t* nn_ptr, size_t n_rows, const uint32_t* k_scales,
                                                size_t n_scales, int eig_order, float* out);
PGEOF_API int pgeof_compute_features_multiscale_dev(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz,
                                                    const uint32_t* nn_ptr, size_t n_rows, const uint32_t* k_scales,
                                                    size_t n_scales, int eig_order, float* out, void* stream);

/* compute_geometric_features_optimal<float,12>, pgeof.hpp:243-310 / pgeof_ext.cpp:86.
 * out (n_rows,12), column 11 = float(k_optimal).  (k_min<1 && k_min_search<1) ->
 * PGEOF_EINVAL (pgeof.hpp:250, sic); k_step == 0 -> PGEOF_EINVAL (reference: UB). */
PGEOF_API int pgeof_compute_features_optimal(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz,
                                             const uint32_t* nn_ptr, size_t n_rows, uint32_t k_min, uint32_t k_step,
                                             uint32_t k_min_search, int eig_order, float* out);
PGEOF_API int pgeof_compute_features_optimal_dev(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz,
                                                 const uint32_t* nn_ptr, size_t n_rows, uint32_t k_min,
                                                 uint32_t k_step, uint32_t k_min_search, int eig_order, float* out,
                                                 void* stream);

/* compute_geometric_features_selected<float|double>, pgeof.hpp:325-375 / pgeof_ext.cpp:148,163.
 * Fused radius search + PCA + the requested features, out (n,n_features) in the
 * caller's order; < 2 points in the ball -> zeros.  feature_ids is a HOST array. */
PGEOF_API int pgeof_compute_features_selected_f32(const float* xyz, size_t n, float search_radius, uint32_t max_knn,
                                                  const int32_t* feature_ids, size_t n_features, int eig_order,
                                                  float* out);
PGEOF_API int pgeof_compute_features_selected_f32_dev(const float* xyz, size_t n, float search_radius,
                                                      uint32_t max_knn, const int32_t* feature_ids,
                                                      size_t n_features, int eig_order, float* out, void* stream);
PGEOF_API int pgeof_compute_features_selected_f64(const double* xyz, size_t n, double search_radius,
                                                  uint32_t max_knn, const int32_t* feature_ids, size_t n_features,
                                                  int eig_order, double* out);
PGEOF_API int pgeof_compute_features_selected_f64_dev(const double* xyz, size_t n, double search_radius,
                                                      uint32_t max_knn, const int32_t* feature_ids,
                                                      size_t n_features, int eig_order, double* out, void* stream);

/* uint64 row offsets (extension, see above): same semantics as the uint32 functions */
PGEOF_API int pgeof_compute_features_p64(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz,
                                         const uint64_t* nn_ptr, size_t n_rows, uint32_t k_min, int eig_order,
                                         float* out);
PGEOF_API int pgeof_compute_features_p64_dev(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz,
                                             const uint64_t* nn_ptr, size_t n_rows, uint32_t k_min, int eig_order,
                                             float* out, void* stream);
PGEOF_API int pgeof_compute_features_multiscale_p64(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz,
                                                    const uint64_t* nn_ptr, size_t n_rows, const uint32_t* k_scales,
                                                    size_t n_scales, int eig_order, float* out);
PGEOF_API int pgeof_compute_features_multiscale_p64_dev(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz,
                                                        const uint64_t* nn_ptr, size_t n_rows,
                                                        const uint32_t* k_scales, size_t n_scales, int eig_order,
                                                        float* out, void* stream);
PGEOF_API int pgeof_compute_features_optimal_p64(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz,
                                                 const uint64_t* nn_ptr, size_t n_rows, uint32_t k_min,
                                                 uint32_t k_step, uint32_t k_min_search, int eig_order, float* out);
PGEOF_API int pgeof_compute_features_optimal_p64_dev(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz,
                                                     const uint64_t* nn_ptr, size_t n_rows, uint32_t k_min,
                                                     uint32_t k_step, uint32_t k_min_search, int eig_order,
                                                     float* out, void* stream);

/* ---- fused pipeline (extension) --------------------------- */
/* knn_search(xyz, xyz, knn) -> CSR view -> compute_features in one call; the
 * (indices, sqr_dist) outputs are optional (NULL = not materialised for the caller). */
PGEOF_API int pgeof_knn_features_dev(const float* xyz, size_t n, uint32_t knn, uint32_t k_min, int eig_order,
                                     uint32_t* indices, float* sqr_dist, float* features, void* stream);
/* host flavour: 12 B per point up, 44 B per point down (+ 8 knn B per point if the lists are requested) */
PGEOF_API int pgeof_knn_features(const float* xyz, size_t n, uint32_t knn, uint32_t k_min, int eig_order,
                                 uint32_t* indices, float* sqr_dist, float* features);

/* ---- spatial query shards for multi-GPU callers (extension) ---
 * The reference has one Taskflow loop over all points (pgeof.hpp:91-109); across GPUs the independent rows are split
 * into slabs along `axis`: slab `rank` of `world` holds ~n/world points, with edges from a 4096-bin histogram of the
 * (replicated) cloud, identical on every rank without a collective.  _plan synchronises the stream once (the caller
 * needs `count` to allocate); _fill writes the rows of the slab in input order and their coordinates. */
typedef struct pgeof_slab_plan {
    float lo, scale;            /* bin(v) = clamp(floor((v - lo) * scale), 0, 4095) */
    uint32_t bin_lo, bin_hi;    /* the slab holds the points whose bin is in [bin_lo, bin_hi) */
    uint64_t count;             /* rows of the slab */
    int axis;
} pgeof_slab_plan;
PGEOF_API int pgeof_slab_plan_dev(const float* xyz, size_t n, int rank, int world, int axis, pgeof_slab_plan* plan,
                                  void* stream);
PGEOF_API int pgeof_slab_fill_dev(const float* xyz, size_t n, const pgeof_slab_plan* plan, int64_t* rows /* [count] */,
                                  float* query /* [count, 3] */, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PGEOF_B200_H */
