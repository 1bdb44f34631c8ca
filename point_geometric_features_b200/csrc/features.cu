// features.cu -- neighbourhood-PCA feature kernels over a CSR (nn, nn_ptr) neighbour list.
// Replaces compute_geometric_features{,_multiscale,_optimal} (include/pgeof.hpp:75-310)
// and pca_from_neighborhood / pca_from_pointcloud (include/pca.hpp:71-129).
//
// Layout.  Rows are processed in SPATIAL order: a counting sort of the rows by the coarse
// cell of their first neighbour (row_order_*) makes consecutive threads / CTAs work on
// overlapping neighbourhoods, so the xyz gathers of a row hit lines its spatial neighbours
// just pulled into L1 / L2 instead of a random 32-B DRAM sector each (the first version read
// 25.8 GB from DRAM for 8.5 GB of algorithmic bytes, profiles/r1a_summary.md).  The cloud is
// re-packed once into 16-B float4 records so that a gather is ONE 128-bit load.
// One CTA owns kRows rows, ONE THREAD PER ROW:
//   stage   the rows' slices of `nn` go to shared memory -- permuted rows: per-warp cp.async
//           (LDGSTS) of each row's contiguous slice; identity order (small inputs): the whole
//           tile is one contiguous span moved by a single 1-D TMA bulk copy (cp.async.bulk +
//           mbarrier) when 16-B alignment allows;
//   walk    each thread walks its neighbour list in shared memory, gathers float4 points with
//           several loads in flight, accumulates the 9 origin-shifted moments (origin = the
//           row's first neighbour), solves the 3x3 eigenproblem in registers
//           (eig3.cuh) and derives the features;
//   store   the tile's features are staged in shared memory and written as 44-B row segments
//           (permuted) or one TMA bulk store (identity order).
//
// Algorithmic bytes per row of length k: 4k (nn) + 4 (nn_ptr) + 12k (xyz gather) + 44 (out)
// = 48 + 16k.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "eig3.cuh"
#include "common.cuh"
#include "ptx.cuh"

namespace pgeof {

namespace {

constexpr int kRows = 128;              // rows per CTA == threads per CTA
constexpr int kMaxScalesPerPass = 8;

struct FeatArgs {
    const float* xyz; uint32_t n_xyz;
    const float4* xyz4;           // 16-B padded copy of xyz (one LDG.128 per gather)
    const uint32_t* order;        // spatial row permutation, or nullptr = identity
    const uint32_t* nn; uint32_t nnz;
    const uint32_t* nn_ptr; uint32_t n_rows;
    uint32_t k_min; int eig_order;
    float* out;
    int* err;
    uint32_t nn_cap;       // entries of `nn` the shared-memory tile can hold
    int tma_in, tma_out;   // pointer alignment allows bulk copies
    int out_by_position;   // direct kernel: `out` is indexed by position in the row sequence (un-permuted by a second pass)
    const uint32_t* out_rows;   // optional: CSR row r is written to out[out_rows[r]] (rows of a compact sub-problem, fused knn_features)
    // multiscale
    uint32_t scales[kMaxScalesPerPass]; uint32_t n_scales_pass; uint32_t n_scales_total; uint32_t scale_base;
    // optimal
    uint32_t k_step, k_min_search;
};

struct MomentsD {
    double sx = 0, sy = 0, sz = 0, sxx = 0, sxy = 0, sxz = 0, syy = 0, syz = 0, szz = 0;
    __device__ __forceinline__ void add(double dx, double dy, double dz)
    {
        sx += dx; sy += dy; sz += dz;
        sxx = fma(dx, dx, sxx); sxy = fma(dx, dy, sxy); sxz = fma(dx, dz, sxz);
        syy = fma(dy, dy, syy); syz = fma(dy, dz, syz); szz = fma(dz, dz, szz);
    }
    __device__ __forceinline__ void cov(uint32_t k, double (&c)[6]) const
    {
        const double inv = 1.0 / (double)k;
        const double mx = sx * inv, my = sy * inv, mz = sz * inv;
        c[0] = sxx * inv - mx * mx; c[1] = sxy * inv - mx * my; c[2] = sxz * inv - mx * mz;
        c[3] = syy * inv - my * my; c[4] = syz * inv - my * mz; c[5] = szz * inv - mz * mz;
    }
};

// Per-thread view of its row after staging.
struct Row {
    uint32_t row;      // CSR row handled by this thread
    uint32_t b, len;   // nn[b, b + len)
    uint32_t s_off;    // offset of the slice inside the shared-memory tile (when staged)
    bool staged;       // slice lives in shared memory (else read nn from global)
    bool valid;        // thread owns a well-formed row
};

struct Tile {
    uint32_t r0, rows;   // positions [r0, r0 + rows) of the (permuted) row sequence
    bool bulk_store;     // identity order: the tile's output is one contiguous block
};

__device__ __forceinline__ void cp_async4(uint32_t* smem_dst, const uint32_t* gmem_src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(ptx::smem_addr(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Stages the nn slices of the CTA's rows into shared memory and returns this thread's row.
__device__ __forceinline__ Row stage_rows(const FeatArgs& a, uint32_t* s_nn, uint32_t* s_rowid, uint64_t* bar, Tile* tile)
{
    const uint32_t r0 = blockIdx.x * kRows;
    const uint32_t rows = min((uint32_t)kRows, a.n_rows - r0);
    tile->r0 = r0; tile->rows = rows; tile->bulk_store = false;
    Row r{0, 0, 0, 0, false, false};
    if (a.order) {
        // ---- permuted rows: every warp stages its own 32 rows with cp.async ---------------
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (threadIdx.x < rows) {
            r.row = __ldg(a.order + r0 + threadIdx.x);
            const uint32_t b = __ldg(a.nn_ptr + r.row), e = __ldg(a.nn_ptr + r.row + 1);
            if (e < b || e > a.nnz) atomicExch(a.err, 1);         // corrupt nn_ptr -> PGEOF_EINDEX, row left 0
            else { r.b = b; r.len = e - b; r.valid = true; }
        }
        s_rowid[threadIdx.x] = r.row;
        uint32_t inc = r.len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        const uint32_t warp_total = __shfl_sync(0xffffffffu, inc, 31);
        const uint32_t warp_cap = a.nn_cap / (kRows / 32);
        r.staged = warp_total <= warp_cap;
        r.s_off = warp * warp_cap + inc - r.len;
        if (r.staged) {
            for (int s = 0; s < 32; ++s) {
                const uint32_t bs = __shfl_sync(0xffffffffu, r.b, s), ls = __shfl_sync(0xffffffffu, r.len, s);
                const uint32_t os = __shfl_sync(0xffffffffu, r.s_off, s);
                for (uint32_t j = lane; j < ls; j += 32) cp_async4(s_nn + os + j, a.nn + bs + j);
            }
            cp_async_wait_all();
            __syncwarp();
        }
        return r;
    }
    // ---- identity order: the tile's nn is one contiguous span -> 1-D TMA bulk copy ---------
    const uint32_t p0 = __ldg(a.nn_ptr + r0), p1 = __ldg(a.nn_ptr + r0 + rows);
    const bool ok = p0 <= p1 && p1 <= a.nnz;
    const uint32_t a0 = p0 & ~3u;
    const bool staged = ok && (p1 - a0) <= a.nn_cap;
    tile->bulk_store = true;
    s_rowid[threadIdx.x] = r0 + threadIdx.x;
    if (!ok) { if (threadIdx.x == 0) atomicExch(a.err, 1); return r; }
    if (staged) {
        const uint32_t a1 = max(p1 & ~3u, a0);     // end of the 16-B aligned body
        if (a.tma_in && a1 > a0) {
            if (threadIdx.x == 0) {
                ptx::mbarrier_init(bar, 1);
                ptx::fence_mbarrier_init();
                const uint32_t bytes = (a1 - a0) * 4u;
                ptx::mbarrier_arrive_expect_tx(bar, bytes);
                ptx::bulk_g2s(s_nn, a.nn + a0, bytes, bar);
            }
            for (uint32_t p = a1 + threadIdx.x; p < p1; p += kRows) s_nn[p - a0] = __ldg(a.nn + p);   // <= 3 entries
            __syncthreads();                      // barrier init visible to the waiters + tail stored
            ptx::mbarrier_wait(bar, 0);
        } else {
            for (uint32_t p = p0 + threadIdx.x; p < p1; p += kRows) s_nn[p - a0] = __ldg(a.nn + p);
            __syncthreads();
        }
    }
    if (threadIdx.x < rows) {
        r.row = r0 + threadIdx.x;
        const uint32_t b = __ldg(a.nn_ptr + r.row), e = __ldg(a.nn_ptr + r.row + 1);
        if (e < b || e > a.nnz) atomicExch(a.err, 1);
        else { r.b = b; r.len = e - b; r.valid = true; r.staged = staged; r.s_off = b - a0; }
    }
    return r;
}

// Writes the tile's staged features: F floats per row.
template <int F>
__device__ __forceinline__ void store_rows(const FeatArgs& a, const Tile& t, const float* s_out, const uint32_t* s_rowid)
{
    __syncthreads();
    const uint32_t total = t.rows * F;
    if (t.bulk_store && a.tma_out && (total & 3u) == 0) {
        if (threadIdx.x == 0) {
            ptx::fence_proxy_async_smem();
            ptx::bulk_s2g(a.out + (size_t)t.r0 * F, s_out, total * 4u);
            ptx::bulk_commit();
            ptx::bulk_wait_read0();
        }
    } else {
        for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
            const uint32_t r = i / F, f = i - r * F;
            a.out[(size_t)s_rowid[r] * F + f] = s_out[i];
        }
    }
}

__device__ __forceinline__ float3 load_point(const FeatArgs& a, uint32_t i)
{
    if (a.xyz4) { const float4 p = __ldg(a.xyz4 + i); return make_float3(p.x, p.y, p.z); }
    return make_float3(__ldg(a.xyz + 3 * (size_t)i), __ldg(a.xyz + 3 * (size_t)i + 1), __ldg(a.xyz + 3 * (size_t)i + 2));
}

// row walker: calls fn(j, dx, dy, dz) for neighbour j = 0..len-1 with coordinates
// relative to the first neighbour; returns false when an index is out of range.
template <typename NnPtr, typename Fn>
__device__ __forceinline__ bool walk_row(const FeatArgs& a, NnPtr src, uint32_t len, Fn&& fn)
{
    const uint32_t i0 = src[0];
    if (i0 >= a.n_xyz) return false;
    const float3 o = load_point(a, i0);
    bool ok = true;
#pragma unroll 4
    for (uint32_t j = 0; j < len; ++j) {
        uint32_t i = src[j];
        if (i >= a.n_xyz) { ok = false; i = i0; }
        const float3 p = load_point(a, i);
        fn(j, p.x - o.x, p.y - o.y, p.z - o.z);
    }
    return ok;
}

struct GlobalNn {   // fallback when the slice does not fit shared memory
    const uint32_t* p;
    __device__ __forceinline__ uint32_t operator[](uint32_t j) const { return __ldg(p + j); }
};

template <typename Fn>
__device__ __forceinline__ bool walk(const FeatArgs& a, const Row& r, const uint32_t* s_nn, uint32_t len, Fn&& fn)
{
    return r.staged ? walk_row(a, s_nn + r.s_off, len, fn) : walk_row(a, GlobalNn{a.nn + r.b}, len, fn);
}

// shared-memory carve-up: [0,128) mbarrier | s_rowid[kRows] | s_out[kRows * F] | s_nn[nn_cap]
template <int F>
struct Smem {
    static constexpr size_t kRowId = 128;
    static constexpr size_t kOut = kRowId + kRows * sizeof(uint32_t);
    static constexpr size_t kNn = kOut + (size_t)kRows * F * sizeof(float);
};

// ----------------------------------------------------------------------------------
// compute_features (pgeof.hpp:75-117)
// ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRows) features_kernel(const FeatArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    uint32_t* s_rowid = reinterpret_cast<uint32_t*>(smem + Smem<11>::kRowId);
    float* s_out = reinterpret_cast<float*>(smem + Smem<11>::kOut);
    uint32_t* s_nn = reinterpret_cast<uint32_t*>(smem + Smem<11>::kNn);
    Tile t;
    const Row r = stage_rows(a, s_nn, s_rowid, bar, &t);
    float f[11];
#pragma unroll
    for (int i = 0; i < 11; ++i) f[i] = 0.f;
    if (r.valid && r.len >= a.k_min && r.len > 0) {          // pgeof.hpp:103
        Moments m;
        auto acc = [&](uint32_t, float dx, float dy, float dz) { m.add(dx, dy, dz); };
        if (!walk(a, r, s_nn, r.len, acc)) atomicExch(a.err, 2);
        else features11<float>(m.pca(r.len, a.eig_order), f);
    }
#pragma unroll
    for (int i = 0; i < 11; ++i) s_out[threadIdx.x * 11 + i] = f[i];
    store_rows<11>(a, t, s_out, s_rowid);
}

// compute_features, second layout: NO shared-memory tile for nn.  Every thread streams its own row of
// nn straight from global memory with 64-bit loads (a row is a contiguous 4k-byte run; its 128-B
// lines stay in L1 between the thread's visits) and keeps 8 independent 128-bit gathers in flight.
// Without the 45 KB tile a CTA needs 6 KB of shared memory (output staging), so occupancy is bound by
// registers only and the SM holds 3x more gathers in flight -- the kernel is bound by the latency of
// random 32-B sector reads, not by issue slots.  HINT selects the gather's cache policy.
// L2 eviction policies: the gathered cloud is re-used ~k times while the rows around a point are processed,
// nn and the outputs stream through once; without hints the stream evicts the cloud (29 % of the gathers missed L2)
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

template <int HINT>
__device__ __forceinline__ float4 gather_point(const float4* p, uint64_t pol)
{
    float4 v;
    if (HINT == 1)
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    else if (HINT == 2)
        asm volatile("ld.global.nc.L1::evict_first.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    else if (HINT >= 3)
        asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    else
        v = __ldg(p);
    return v;
}

// eight consecutive entries of nn with ONE 256-bit load that does not allocate in L1 (sm_100 LDG.256): the stream
// is read exactly once, and 1024 resident rows x one 128-B line each would otherwise crowd the gathered
// points out of L1.  p must be 32-byte aligned.
template <int HINT>
__device__ __forceinline__ void stream_nn8(const uint32_t* p, uint32_t (&i)[8])
{
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(i[0]), "=r"(i[1]), "=r"(i[2]), "=r"(i[3]), "=r"(i[4]), "=r"(i[5]), "=r"(i[6]), "=r"(i[7]) : "l"(p));
}

// up to N (<= 7) consecutive entries: indices first, then their gathers together, then the moments in order
template <int HINT, int N, typename Acc>
__device__ __forceinline__ void walk_some(const FeatArgs& a, const uint32_t* __restrict__ p, uint32_t j0, uint32_t cnt, uint32_t i0, const float4& o,
                                          uint64_t pol_keep, Acc& acc, bool& ok)
{
    uint32_t i[N];
    float4 q[N];
#pragma unroll
    for (int u = 0; u < N; ++u) {
        i[u] = (uint32_t)u < cnt ? __ldg(p + u) : i0;
        if (i[u] >= a.n_xyz) { ok = false; i[u] = i0; }
    }
#pragma unroll
    for (int u = 0; u < N; ++u) q[u] = gather_point<HINT>(a.xyz4 + i[u], pol_keep);
#pragma unroll
    for (int u = 0; u < N; ++u) if ((uint32_t)u < cnt) acc(j0 + u, q[u].x - o.x, q[u].y - o.y, q[u].z - o.z);
}

// acc(j, dx, dy, dz) is called for j = 0 .. len-1 in order, offsets relative to the row's first neighbour
template <int HINT, typename Acc>
__device__ __forceinline__ bool walk_direct(const FeatArgs& a, uint32_t b, uint32_t len, Acc& acc)
{
    const uint32_t* __restrict__ p = a.nn + b;
    const uint32_t n = a.n_xyz;
    const uint32_t i0 = __ldg(p);
    if (i0 >= n) return false;
    const uint64_t pol_keep = HINT >= 3 ? l2_policy_evict_last() : 0;
    const float4 o = gather_point<HINT>(a.xyz4 + i0, pol_keep);   // origin of the shifted moments; its own term is zero
    bool ok = true;
    acc(0u, 0.f, 0.f, 0.f);
    uint32_t j = 1;
    // head: up to the next 32-byte boundary of the stream (whatever the alignment of nn itself)
    const uint32_t nh = min((uint32_t)((0u - (uint32_t)reinterpret_cast<uintptr_t>(p + 1)) & 31u) >> 2, len - 1u);
    if (nh) { walk_some<HINT, 7>(a, p + j, j, nh, i0, o, pol_keep, acc, ok); j += nh; }
    for (; j + 8 <= len; j += 8) {
        uint32_t i[8];
        stream_nn8<HINT>(p + j, i);
        float4 q[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (i[u] >= n) { ok = false; i[u] = i0; }
            q[u] = gather_point<HINT>(a.xyz4 + i[u], pol_keep);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) acc(j + u, q[u].x - o.x, q[u].y - o.y, q[u].z - o.z);
    }
    if (j < len) walk_some<HINT, 7>(a, p + j, j, len - j, i0, o, pol_keep, acc, ok);
    return ok;
}

// THREADS rows per tile, one thread per row.  Every CTA owns a CONTIGUOUS chunk of the (spatially ordered) row
// sequence and walks it tile by tile: the rows an SM works on at any time form one or two compact blobs, so a
// gathered point is re-used out of L1 by the blob's other rows (small independent CTAs spread over the whole
// in-flight window shared their gathers only through L2).
template <int HINT, int THREADS>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS) features_direct_kernel(const FeatArgs a)
{
    __shared__ uint32_t s_rowid[THREADS];
    __shared__ float s_out[THREADS * 11];
    const uint32_t tiles = (a.n_rows + THREADS - 1) / THREADS;
    const uint32_t per = (tiles + gridDim.x - 1) / gridDim.x;
    const uint32_t t_end = min(tiles, (blockIdx.x + 1) * per);
    for (uint32_t tile = blockIdx.x * per; tile < t_end; ++tile) {
        const uint32_t r0 = tile * THREADS;
        // out_by_position: a.out is indexed by POSITION in the (permuted) row sequence, the tile's output is one contiguous block
        Tile t{r0, min((uint32_t)THREADS, a.n_rows - r0), (a.order == nullptr && a.out_rows == nullptr) || a.out_by_position != 0};
        float f[11];
#pragma unroll
        for (int i = 0; i < 11; ++i) f[i] = 0.f;
        uint32_t row = r0 + threadIdx.x;
        if (threadIdx.x < t.rows && a.order) row = __ldg(a.order + row);
        s_rowid[threadIdx.x] = a.out_by_position ? r0 + threadIdx.x : ((a.out_rows && threadIdx.x < t.rows) ? __ldg(a.out_rows + row) : row);
        if (threadIdx.x < t.rows) {
            const uint32_t b = __ldg(a.nn_ptr + row), e = __ldg(a.nn_ptr + row + 1);
            if (e < b || e > a.nnz) atomicExch(a.err, 1);        // corrupt nn_ptr -> PGEOF_EINDEX, row left 0
            else if (e - b >= a.k_min && e > b) {                // pgeof.hpp:103
                Moments m;
                auto acc = [&](uint32_t, float dx, float dy, float dz) { m.add(dx, dy, dz); };
                if (!walk_direct<HINT>(a, b, e - b, acc)) atomicExch(a.err, 2);
                else features11<float>(m.pca(e - b, a.eig_order), f);
            }
        }
#pragma unroll
        for (int i = 0; i < 11; ++i) s_out[threadIdx.x * 11 + i] = f[i];
        store_rows<11>(a, t, s_out, s_rowid);
        __syncthreads();                                         // the staging buffers are re-used by the next tile
    }
}

// compute_features_multiscale on the direct walker: 512 rows per CTA, no shared memory (all of it stays L1 for the
// gathers); a row's 11 floats of a scale are stored straight from the thread when the walk reaches that prefix
// length (44 contiguous bytes; rows too short for a scale get zeros, pgeof.hpp:175,193)
constexpr int kMsThreads = 256;

__device__ __forceinline__ void store11(float* dst, const float (&f)[11])
{
#pragma unroll
    for (int i = 0; i < 11; ++i) dst[i] = f[i];
}

// kept out of line: the walker inlines its accumulator at 22 sites, the eigen solve + 11 formulas are ~400 instructions
__device__ __noinline__ void emit_scale(Moments m, uint32_t k, int eig_order, float* dst)
{
    float f[11];
    features11<float>(m.pca(k, eig_order), f);
    store11(dst, f);
}

__global__ void __launch_bounds__(kMsThreads, 3) multiscale_direct_kernel(const FeatArgs a)
{
    uint32_t row = blockIdx.x * kMsThreads + threadIdx.x;
    if (row >= a.n_rows) return;
    if (a.order) row = __ldg(a.order + row);
    float* out = a.out + ((size_t)row * a.n_scales_total + a.scale_base) * 11;
    const float zero[11] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const uint32_t b = __ldg(a.nn_ptr + row), e = __ldg(a.nn_ptr + row + 1);
    uint32_t n_fit = 0, s = 0;                                  // scales of this pass the row is long enough for / written so far
    if (e < b || e > a.nnz) atomicExch(a.err, 1);
    else {
        const uint32_t len = e - b;
        while (n_fit < a.n_scales_pass && a.scales[n_fit] <= len) ++n_fit;
        if (n_fit && a.scales[n_fit - 1] > 0) {
            Moments m;
            while (s < n_fit && a.scales[s] == 0) { store11(out + s * 11, zero); ++s; }   // k_s = 0 is rejected on the host; defensive
            uint32_t next_k = s < n_fit ? a.scales[s] : 0xffffffffu;      // prefix length of the next scale to emit (register, not a.scales[s])
            auto acc = [&](uint32_t j, float dx, float dy, float dz) {
                m.add(dx, dy, dz);
                if (j + 1 == next_k) {
                    do {
                        emit_scale(m, j + 1, a.eig_order, out + s * 11);
                        ++s;
                        next_k = s < n_fit ? a.scales[s] : 0xffffffffu;
                    } while (next_k == j + 1);
                }
            };
            if (!walk_direct<0>(a, b, a.scales[n_fit - 1], acc)) { atomicExch(a.err, 2); s = 0; }
        }
    }
    for (; s < a.n_scales_pass; ++s) store11(out + s * 11, zero);
}

// compute_features_optimal on the direct walker.  Prefix moments in double as in optimal_kernel; the entropy of every
// candidate size is first evaluated in FLOAT from Jacobi eigenvalues (accurate to a few ulp of the matrix norm for
// every eigenvalue, so the float and double entropies differ by < 3e-5 in the worst case, ~1e-6 typically).  A
// candidate that wins or loses by more than kOptMargin is decided by the float values; anything closer is decided by
// the double values exactly as the all-double scan does (strict '<', smallest k wins ties, pgeof.hpp:289), so the
// chosen k is identical.  The evaluation is kept out of line: the walker inlines its accumulator at 22 sites.
constexpr float kOptMargin = 2e-4f;

struct OptState {
    double best_c[6];
    double best_h64;
    float best_h32;
    uint32_t best_k;
    int have64;
};

__device__ __forceinline__ double entropy_f64(const double (&c)[6])
{
    double w[3];
    eigvals3_f64(c[0], c[1], c[2], c[3], c[4], c[5], w);
    return eigentropy_of<double>(w[0], w[1], w[2]);
}

__device__ __noinline__ void optimal_eval(MomentsD m, uint32_t k, int first, OptState* st)
{
    double c[6];
    m.cov(k, c);
    float w[3];
    jacobi_eigvals_f32((float)c[0], (float)c[1], (float)c[2], (float)c[3], (float)c[4], (float)c[5], w);
    const float h32 = eigentropy_of<float>(w[0], w[1], w[2]);
    bool take = first || h32 < st->best_h32 - kOptMargin;
    int exact = 0;
    double h64 = 0.0;
    if (!take && !(h32 > st->best_h32 + kOptMargin)) {                      // too close to call in float
        if (!st->have64) { st->best_h64 = entropy_f64(st->best_c); st->have64 = 1; }
        h64 = entropy_f64(c);
        take = h64 < st->best_h64;                                          // pgeof.hpp:289
        exact = 1;
    }
    if (take) {
        st->best_k = k; st->best_h32 = h32; st->best_h64 = h64; st->have64 = exact;
#pragma unroll
        for (int i = 0; i < 6; ++i) st->best_c[i] = c[i];
    }
}

__global__ void __launch_bounds__(kRows, 4) optimal_direct_kernel(const FeatArgs a)
{
    __shared__ uint32_t s_rowid[kRows];
    __shared__ float s_out[kRows * 12];
    const uint32_t r0 = blockIdx.x * kRows;
    Tile t{r0, min((uint32_t)kRows, a.n_rows - r0), a.order == nullptr};
    float f[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) f[i] = 0.f;
    uint32_t row = r0 + threadIdx.x;
    if (threadIdx.x < t.rows && a.order) row = __ldg(a.order + row);
    s_rowid[threadIdx.x] = row;
    if (threadIdx.x < t.rows) {
        const uint32_t b = __ldg(a.nn_ptr + row), e = __ldg(a.nn_ptr + row + 1);
        if (e < b || e > a.nnz) atomicExch(a.err, 1);
        else {
            const uint32_t len = e - b;
            if (len >= a.k_min && len >= a.k_min_search && len > 0) {                     // pgeof.hpp:272
                const uint32_t k0 = min(max(max(a.k_min, a.k_min_search), 1u), len);    // :274
                MomentsD m;
                OptState st;
                st.best_k = len; st.best_h32 = 0.f; st.best_h64 = 0.0; st.have64 = 0;
#pragma unroll
                for (int i = 0; i < 6; ++i) st.best_c[i] = 0.0;
                uint32_t rem = 0;                                                        // k % k_step, kept incrementally
                auto acc = [&](uint32_t j, float dx, float dy, float dz) {
                    m.add((double)dx, (double)dy, (double)dz);
                    const uint32_t k = j + 1;
                    rem = rem + 1 == a.k_step ? 0u : rem + 1;
                    if (k < k0 || (k > k0 && rem != 0 && k != len)) return;              // :283
                    optimal_eval(m, k, k == k0, &st);
                };
                if (!walk_direct<0>(a, b, len, acc)) atomicExch(a.err, 2);
                else {
                    float g[11];
                    features11<float>(pca_from_cov<float>((float)st.best_c[0], (float)st.best_c[1], (float)st.best_c[2], (float)st.best_c[3],
                                                         (float)st.best_c[4], (float)st.best_c[5], a.eig_order), g);
#pragma unroll
                    for (int i = 0; i < 11; ++i) f[i] = g[i];
                    f[11] = (float)st.best_k;
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) s_out[threadIdx.x * 12 + i] = f[i];
    store_rows<12>(a, t, s_out, s_rowid);
}

// out[row] = tmp[inv[row]]: undoes the spatial row permutation with random 4F-byte READS and fully coalesced
// writes (scattered 44-B row writes from the feature kernel cost more than the whole neighbourhood walk:
// partial-sector writes, profiles/r1f)
template <int F>
__global__ void __launch_bounds__(kRows) unpermute_kernel(const float* __restrict__ tmp, const uint32_t* __restrict__ inv, uint32_t n_rows,
                                                          float* __restrict__ out)
{
    __shared__ float s[kRows * F];
    const uint32_t r0 = blockIdx.x * kRows;
    const uint32_t rows = min((uint32_t)kRows, n_rows - r0);
    if (threadIdx.x < rows) {
        const float* src = tmp + (size_t)__ldg(inv + r0 + threadIdx.x) * F;
#pragma unroll
        for (int f = 0; f < F; ++f) s[threadIdx.x * F + f] = __ldg(src + f);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < rows * F; i += kRows) out[(size_t)r0 * F + i] = s[i];
}

// ----------------------------------------------------------------------------------
// compute_features_multiscale (pgeof.hpp:159-211): scale s uses the first k_s entries of
// the row; one walk yields every scale from the running (prefix) moments.
// ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRows) multiscale_kernel(const FeatArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    uint32_t* s_rowid = reinterpret_cast<uint32_t*>(smem + Smem<0>::kRowId);
    uint32_t* s_nn = reinterpret_cast<uint32_t*>(smem + Smem<0>::kNn);
    Tile t;
    const Row r = stage_rows(a, s_nn, s_rowid, bar, &t);
    if (!r.valid) return;
    // rows are only walked up to the largest scale of this pass that fits (pgeof.hpp:193 early break)
    uint32_t n_fit = 0;
    while (n_fit < a.n_scales_pass && a.scales[n_fit] <= r.len) ++n_fit;
    if (n_fit == 0 || a.scales[n_fit - 1] == 0) return;
    const uint32_t walk_len = a.scales[n_fit - 1];
    float* out = a.out + ((size_t)r.row * a.n_scales_total + a.scale_base) * 11;
    Moments m;
    uint32_t s = 0;
    while (s < n_fit && a.scales[s] == 0) ++s;   // k_s = 0 is rejected on the host; defensive
    auto acc = [&](uint32_t j, float dx, float dy, float dz) {
        m.add(dx, dy, dz);
        while (s < n_fit && a.scales[s] == j + 1) {
            float f[11];
            features11<float>(m.pca(j + 1, a.eig_order), f);
#pragma unroll
            for (int i = 0; i < 11; ++i) out[s * 11 + i] = f[i];
            ++s;
        }
    };
    if (!walk(a, r, s_nn, walk_len, acc)) {
        atomicExch(a.err, 2);
        for (uint32_t i = 0; i < n_fit * 11; ++i) out[i] = 0.f;
    }
}

// ----------------------------------------------------------------------------------
// compute_features_optimal (pgeof.hpp:243-310): Weinmann eigenentropy scan.  The prefix
// moments are accumulated in double so that the entropy of every evaluated k is accurate
// to ~1e-13 and the arg-min (strict '<', smallest k wins ties) is reproducible.
// ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRows) optimal_kernel(const FeatArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    uint32_t* s_rowid = reinterpret_cast<uint32_t*>(smem + Smem<12>::kRowId);
    float* s_out = reinterpret_cast<float*>(smem + Smem<12>::kOut);
    uint32_t* s_nn = reinterpret_cast<uint32_t*>(smem + Smem<12>::kNn);
    Tile t;
    const Row r = stage_rows(a, s_nn, s_rowid, bar, &t);
    float f[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) f[i] = 0.f;
    const uint32_t len = r.len;
    if (r.valid && len >= a.k_min && len >= a.k_min_search && len > 0) {               // pgeof.hpp:272
        const uint32_t k0 = min(max(max(a.k_min, a.k_min_search), 1u), len);           // :274
        MomentsD m;
        double best_h = 1.0, best_c[6] = {0, 0, 0, 0, 0, 0};
        uint32_t best_k = len;
        auto acc = [&](uint32_t j, float dx, float dy, float dz) {
            m.add((double)dx, (double)dy, (double)dz);
            const uint32_t k = j + 1;
            if (k < k0) return;
            if (k > k0 && (k % a.k_step) != 0 && k != len) return;                      // :283
            double c[6], w[3];
            m.cov(k, c);
            eigvals3_f64(c[0], c[1], c[2], c[3], c[4], c[5], w);
            const double h = eigentropy_of<double>(w[0], w[1], w[2]);
            if (k == k0 || h < best_h) {                                                // :289
                best_h = h; best_k = k;
#pragma unroll
                for (int i = 0; i < 6; ++i) best_c[i] = c[i];
            }
        };
        if (!walk(a, r, s_nn, len, acc)) atomicExch(a.err, 2);
        else {
            float g[11];
            features11<float>(pca_from_cov<float>((float)best_c[0], (float)best_c[1], (float)best_c[2], (float)best_c[3],
                                                 (float)best_c[4], (float)best_c[5], a.eig_order), g);
#pragma unroll
            for (int i = 0; i < 11; ++i) f[i] = g[i];
            f[11] = (float)best_k;
        }
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) s_out[threadIdx.x * 12 + i] = f[i];
    store_rows<12>(a, t, s_out, s_rowid);
}

// ----------------------------------------------------------------------------------
// pre-passes: float4 re-pack of the cloud, spatial ordering of the rows
// ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pad_xyz_kernel(const float* __restrict__ xyz, size_t n, float4* __restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_float4(__ldg(xyz + 3 * i), __ldg(xyz + 3 * i + 1), __ldg(xyz + 3 * i + 2), 0.f);
}

struct RowGrid { float lo[3]; float scale[3]; int cells; int morton; };

// one warp: reduce the bbox partials and derive the coarse row-ordering grid (no host sync)
__global__ void row_grid_kernel(const float* __restrict__ partial, int n_partial, int cells, int morton, RowGrid* __restrict__ g)
{
    const int lane = threadIdx.x;
    float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (int b = lane; b < n_partial; b += 32)
        for (int d = 0; d < 3; ++d) { mn[d] = fminf(mn[d], partial[b * 6 + d]); mx[d] = fmaxf(mx[d], partial[b * 6 + 3 + d]); }
    for (int d = 0; d < 3; ++d)
        for (int o = 16; o > 0; o >>= 1) {
            mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
            mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
        }
    if (lane == 0) {
        for (int d = 0; d < 3; ++d) {
            const float ext = mx[d] - mn[d];
            g->lo[d] = mn[d];
            g->scale[d] = (ext > 0.f && ext < 3.0e38f) ? (float)cells / ext : 0.f;
        }
        g->cells = cells;
        g->morton = morton;
    }
}

// bits of a 10-bit value spread to every third position
__device__ __forceinline__ uint32_t spread3(uint32_t v)
{
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

__device__ __forceinline__ uint32_t row_key(const FeatArgs& a, const RowGrid& g, uint32_t row)
{
    const uint32_t b = __ldg(a.nn_ptr + row), e = __ldg(a.nn_ptr + row + 1);
    if (e <= b || b >= a.nnz) return 0;
    const uint32_t i = __ldg(a.nn + b);
    if (i >= a.n_xyz) return 0;
    const float4 p = __ldg(a.xyz4 + i);
    const int c = g.cells;
    const int cx = min(max(__float2int_rd((p.x - g.lo[0]) * g.scale[0]), 0), c - 1);
    const int cy = min(max(__float2int_rd((p.y - g.lo[1]) * g.scale[1]), 0), c - 1);
    const int cz = min(max(__float2int_rd((p.z - g.lo[2]) * g.scale[2]), 0), c - 1);
    // Morton order of the coarse cells: rows that share neighbours are visited close in time along all three
    // axes, so a gathered point is still in L2 when the next row needs it (a (z, y, x) raster order revisits the
    // neighbouring plane only after a whole plane of rows: more gather traffic than the L2 holds)
    if (g.morton) return spread3((uint32_t)cx) | (spread3((uint32_t)cy) << 1) | (spread3((uint32_t)cz) << 2);
    return ((uint32_t)cz * c + cy) * c + cx;
}

__global__ void __launch_bounds__(256) row_count_kernel(const FeatArgs a, const RowGrid* __restrict__ gp, uint32_t* __restrict__ counts,
                                                        uint32_t* __restrict__ keys, uint32_t* __restrict__ rank)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < a.n_rows;
    const RowGrid g = *gp;
    uint32_t key = 0xffffffffu;
    if (valid) key = row_key(a, g, i);
    const unsigned active = __ballot_sync(0xffffffffu, valid);
    if (!valid) return;
    const unsigned peers = __match_any_sync(active, key);
    const int leader = __ffs(peers) - 1;
    const int lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counts + key, (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    unsigned lt;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt));
    keys[i] = key;
    rank[i] = base + (uint32_t)__popc(peers & lt);
}

__global__ void __launch_bounds__(256) row_scatter_kernel(uint32_t n_rows, const uint32_t* __restrict__ starts, const uint32_t* __restrict__ keys,
                                                          const uint32_t* __restrict__ rank, uint32_t* __restrict__ order,
                                                          uint32_t* __restrict__ inverse)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_rows) {
        const uint32_t pos = __ldg(starts + keys[i]) + rank[i];
        order[pos] = i;
        inverse[i] = pos;
    }
}

// shared-memory tile for `nn`: mean row length with 50 % head-room, at least 32 entries a row
uint32_t pick_nn_cap(size_t nnz, size_t n_rows, size_t fixed_bytes)
{
    const double mean = n_rows ? (double)nnz / (double)n_rows : 0.0;
    size_t cap = (size_t)(kRows * std::max(32.0, mean * 1.5)) + 16;
    const size_t max_bytes = 96 * 1024 - fixed_bytes;          // keep >= 2 CTAs per SM
    cap = std::min(cap, max_bytes / 4);
    return (uint32_t)(cap & ~(size_t)15);                      // per-warp quarter stays 16-B aligned
}

int make_args(FeatArgs* a, const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz, const uint32_t* nn_ptr, size_t n_rows,
              int eig_order, float* out, int* err, size_t floats_per_row)
{
    if (n_xyz > 0xffffffffull || nnz > 0xffffffffull || n_rows > 0xfffffffeull) { set_error("array too large for uint32 CSR"); return PGEOF_EINVAL; }
    if (eig_order != PGEOF_EIG_LITERAL && eig_order != PGEOF_EIG_DOCUMENTED) { set_error("bad eig_order %d", eig_order); return PGEOF_EINVAL; }
    std::memset(a, 0, sizeof(*a));
    a->xyz = xyz; a->n_xyz = (uint32_t)n_xyz; a->nn = nn; a->nnz = (uint32_t)nnz; a->nn_ptr = nn_ptr; a->n_rows = (uint32_t)n_rows;
    a->eig_order = eig_order; a->out = out; a->err = err; a->k_min = 1; a->k_step = 1; a->k_min_search = 1;
    a->tma_in = ((uintptr_t)nn % 16 == 0);
    a->tma_out = ((uintptr_t)out % 16 == 0) && ((kRows * floats_per_row * 4) % 16 == 0);
    return PGEOF_OK;
}

// Device buffers of the two pre-passes; they live until the feature kernel was enqueued
// (stream-ordered frees).
struct Prepass {
    DeviceBuffer xyz4, order, inverse;
};

int env_int(const char* name, int dflt)
{
    const char* e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}


int prepare(FeatArgs* a, Prepass* p, cudaStream_t stream)
{
    // 1. float4 re-pack of the cloud: one 128-bit load per gathered neighbour
    PGEOF_TRY(p->xyz4.alloc((size_t)std::max<uint32_t>(a->n_xyz, 1) * sizeof(float4), stream));
    if (a->n_xyz) {
        pad_xyz_kernel<<<(a->n_xyz + 255) / 256, 256, 0, stream>>>(a->xyz, a->n_xyz, p->xyz4.as<float4>());
        PGEOF_LAUNCH_CHECK();
    }
    a->xyz4 = p->xyz4.as<float4>();
    // 2. spatial row order (counting sort by the coarse cell of the first neighbour)
    const int min_rows = env_int("PGEOF_FEATURES_SORT_MIN_ROWS", 32768);
    if (a->n_xyz == 0 || a->nnz == 0 || (int64_t)a->n_rows < (int64_t)min_rows || env_int("PGEOF_FEATURES_SORT", 1) == 0) return PGEOF_OK;
    KernelTimer timer("row_order", stream);
    DeviceBuffer partial, grid, counts, keys, rank;
    int n_partial = 0;
    PGEOF_TRY(bbox_partials(a->xyz, a->n_xyz, &partial, &n_partial, stream));
    int cells = (int)std::lround(std::cbrt((double)a->n_rows / 6.0));
    const int morton = env_int("PGEOF_FEATURES_MORTON", 1);
    cells = std::min(std::max(cells, 8), morton ? 128 : 160);
    size_t n_cells = (size_t)cells * cells * cells;
    if (morton) { int bits = 3; while ((1 << bits) < cells) ++bits; n_cells = (size_t)1 << (3 * bits); }
    PGEOF_TRY(grid.alloc(sizeof(RowGrid), stream));
    PGEOF_TRY(counts.alloc((n_cells + 1) * sizeof(uint32_t), stream));
    PGEOF_TRY(keys.alloc((size_t)a->n_rows * sizeof(uint32_t), stream));
    PGEOF_TRY(rank.alloc((size_t)a->n_rows * sizeof(uint32_t), stream));
    PGEOF_TRY(p->order.alloc((size_t)a->n_rows * sizeof(uint32_t), stream));
    PGEOF_TRY(p->inverse.alloc((size_t)a->n_rows * sizeof(uint32_t), stream));
    row_grid_kernel<<<1, 32, 0, stream>>>(partial.as<float>(), n_partial, cells, morton, grid.as<RowGrid>());
    PGEOF_LAUNCH_CHECK();
    PGEOF_CUDA(cudaMemsetAsync(counts.ptr, 0, (n_cells + 1) * sizeof(uint32_t), stream));
    const unsigned blocks = (a->n_rows + 255) / 256;
    row_count_kernel<<<blocks, 256, 0, stream>>>(*a, grid.as<RowGrid>(), counts.as<uint32_t>(), keys.as<uint32_t>(), rank.as<uint32_t>());
    PGEOF_LAUNCH_CHECK();
    PGEOF_TRY(exclusive_scan_u32(counts.as<uint32_t>(), n_cells, stream));
    row_scatter_kernel<<<blocks, 256, 0, stream>>>(a->n_rows, counts.as<uint32_t>(), keys.as<uint32_t>(), rank.as<uint32_t>(), p->order.as<uint32_t>(), p->inverse.as<uint32_t>());
    PGEOF_LAUNCH_CHECK();
    a->order = p->order.as<uint32_t>();
    return PGEOF_OK;
}

template <typename K>
int launch_tiles(K kern, const char* name, const FeatArgs& a, size_t smem, cudaStream_t stream)
{
    if (smem > 48 * 1024) PGEOF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned blocks = (a.n_rows + kRows - 1) / kRows;
    {
        KernelTimer timer(name, stream);
        kern<<<blocks, kRows, smem, stream>>>(a);
    }
    PGEOF_LAUNCH_CHECK();
    return PGEOF_OK;
}

}  // namespace

int device_flag_check(const int* d_flag, cudaStream_t stream, const char* what)
{
    int h = 0;
    PGEOF_CUDA(cudaMemcpyAsync(&h, d_flag, sizeof(int), cudaMemcpyDeviceToHost, stream));
    PGEOF_CUDA(cudaStreamSynchronize(stream));
    if (h == 1) { set_error("%s: nn_ptr is not a non-decreasing offset array into nn", what); return PGEOF_EINDEX; }
    if (h != 0) { set_error("%s: nn holds an index >= len(xyz)", what); return PGEOF_EINDEX; }
    return PGEOF_OK;
}

int features_run(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz, const uint32_t* nn_ptr, size_t n_rows,
                 uint32_t k_min, int eig_order, float* out, cudaStream_t stream, const uint32_t* out_rows)
{
    if (n_rows == 0) return PGEOF_OK;
    DeviceBuffer err;
    PGEOF_TRY(err.alloc(sizeof(int), stream));
    PGEOF_CUDA(cudaMemsetAsync(err.ptr, 0, sizeof(int), stream));
    FeatArgs a;
    PGEOF_TRY(make_args(&a, xyz, n_xyz, nn, nnz, nn_ptr, n_rows, eig_order, out, err.as<int>(), 11));
    a.k_min = k_min;
    a.out_rows = out_rows;
    Prepass pre;
    PGEOF_TRY(prepare(&a, &pre, stream));
    const int layout = out_rows ? 1 : env_int("PGEOF_FEATURES_LAYOUT", 1);   // 1: direct nn stream (default), 0: shared-memory nn tile
    if (layout == 0) {
        const size_t fixed = Smem<11>::kNn;
        a.nn_cap = pick_nn_cap(nnz, n_rows, fixed);
        PGEOF_TRY(launch_tiles(features_kernel, "features", a, fixed + (size_t)a.nn_cap * 4, stream));
    } else {
        // permuted rows: features land in a position-indexed scratch block, then one gather pass restores row order
        DeviceBuffer tmp;
        const bool unpermute = a.order && !out_rows && env_int("PGEOF_FEATURES_UNPERMUTE", 0) != 0;   // measured: no gain over scattered row writes
        if (unpermute) {
            a.out_by_position = 1;
            PGEOF_TRY(tmp.alloc(n_rows * 11 * sizeof(float), stream));
            a.out = tmp.as<float>();
            a.tma_out = 1;
        }
        const int hint = env_int("PGEOF_FEATURES_HINT", 0);   // 3: L2 evict_last gathers (no gain, with or without a persisting set-aside)
        const int cta = env_int("PGEOF_FEATURES_CTA", 512);
        int dev = 0, sms = 148;
        PGEOF_CUDA(cudaGetDevice(&dev));
        PGEOF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        auto launch = [&](auto kern, int threads) -> int {
            const unsigned tiles = (unsigned)((n_rows + threads - 1) / threads);
            const unsigned blocks = std::min<unsigned>(tiles, (unsigned)(sms * (1024 / threads) * env_int("PGEOF_FEATURES_WAVES", 1 << 20)));
            {
                KernelTimer timer("features", stream);
                kern<<<blocks, threads, 0, stream>>>(a);
            }
            PGEOF_LAUNCH_CHECK();
            return PGEOF_OK;
        };
        a.tma_out = a.tma_out && ((cta * 11 * 4) % 16 == 0);
        if (hint == 3 && cta == 512) PGEOF_TRY(launch(features_direct_kernel<3, 512>, 512));
        else if (cta == 128) PGEOF_TRY(launch(features_direct_kernel<0, 128>, 128));
        else if (cta == 256) PGEOF_TRY(launch(features_direct_kernel<0, 256>, 256));
        else if (cta == 1024) PGEOF_TRY(launch(features_direct_kernel<0, 1024>, 1024));
        else PGEOF_TRY(launch(features_direct_kernel<0, 512>, 512));
        if (unpermute) {
            KernelTimer timer("row_order", stream);
            unpermute_kernel<11><<<(unsigned)((n_rows + kRows - 1) / kRows), kRows, 0, stream>>>(tmp.as<float>(), pre.inverse.as<uint32_t>(), (uint32_t)n_rows, out);
            PGEOF_LAUNCH_CHECK();
        }
    }
    return device_flag_check(err.as<int>(), stream, "compute_features");
}

int features_multiscale_run(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz, const uint32_t* nn_ptr,
                            size_t n_rows, const uint32_t* k_scales_host, size_t n_scales, int eig_order, float* out,
                            cudaStream_t stream)
{
    if (n_rows == 0 || n_scales == 0) return PGEOF_OK;
    if (n_scales > 0xffffffffull / 11) { set_error("too many scales"); return PGEOF_EINVAL; }
    DeviceBuffer err;
    PGEOF_TRY(err.alloc(sizeof(int), stream));
    PGEOF_CUDA(cudaMemsetAsync(err.ptr, 0, sizeof(int), stream));
    if (env_int("PGEOF_FEATURES_LAYOUT", 1) == 0)   // the direct kernel writes every element itself (zeros where a row is too short)
        PGEOF_CUDA(cudaMemsetAsync(out, 0, n_rows * n_scales * 11 * sizeof(float), stream));   // calloc semantics, pgeof.hpp:175
    FeatArgs a;
    PGEOF_TRY(make_args(&a, xyz, n_xyz, nn, nnz, nn_ptr, n_rows, eig_order, out, err.as<int>(), 11));
    a.n_scales_total = (uint32_t)n_scales;
    Prepass pre;
    PGEOF_TRY(prepare(&a, &pre, stream));
    const size_t fixed = Smem<0>::kNn;
    a.nn_cap = pick_nn_cap(nnz, n_rows, fixed);
    for (size_t base = 0; base < n_scales; base += kMaxScalesPerPass) {
        a.scale_base = (uint32_t)base;
        a.n_scales_pass = (uint32_t)std::min<size_t>(kMaxScalesPerPass, n_scales - base);
        for (uint32_t s = 0; s < a.n_scales_pass; ++s) a.scales[s] = k_scales_host[base + s];
        if (env_int("PGEOF_FEATURES_LAYOUT", 1) == 0) PGEOF_TRY(launch_tiles(multiscale_kernel, "multiscale", a, fixed + (size_t)a.nn_cap * 4, stream));
        else {
            KernelTimer timer("multiscale", stream);
            multiscale_direct_kernel<<<(unsigned)((n_rows + kMsThreads - 1) / kMsThreads), kMsThreads, 0, stream>>>(a);
            PGEOF_LAUNCH_CHECK();
        }
    }
    return device_flag_check(err.as<int>(), stream, "compute_features_multiscale");
}

int features_optimal_run(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz, const uint32_t* nn_ptr,
                         size_t n_rows, uint32_t k_min, uint32_t k_step, uint32_t k_min_search, int eig_order, float* out,
                         cudaStream_t stream)
{
    if (n_rows == 0) return PGEOF_OK;
    DeviceBuffer err;
    PGEOF_TRY(err.alloc(sizeof(int), stream));
    PGEOF_CUDA(cudaMemsetAsync(err.ptr, 0, sizeof(int), stream));
    FeatArgs a;
    PGEOF_TRY(make_args(&a, xyz, n_xyz, nn, nnz, nn_ptr, n_rows, eig_order, out, err.as<int>(), 12));
    a.k_min = k_min; a.k_step = k_step; a.k_min_search = k_min_search;
    Prepass pre;
    PGEOF_TRY(prepare(&a, &pre, stream));
    const size_t fixed = Smem<12>::kNn;
    a.nn_cap = pick_nn_cap(nnz, n_rows, fixed);
    // layout 1: float Jacobi filter + double recheck of near ties on the direct walker; layout 0: all-double scan over a
    // shared-memory nn tile.  (A float CLOSED-FORM filter was tried first and dropped: it loses ~1e-3 when two
    // eigenvalues nearly coincide, which no fixed margin covers.)
    if (env_int("PGEOF_FEATURES_LAYOUT", 1) == 0) PGEOF_TRY(launch_tiles(optimal_kernel, "optimal", a, fixed + (size_t)a.nn_cap * 4, stream));
    else PGEOF_TRY(launch_tiles(optimal_direct_kernel, "optimal", a, 0, stream));
    return device_flag_check(err.as<int>(), stream, "compute_features_optimal");
}

}  // namespace pgeof
