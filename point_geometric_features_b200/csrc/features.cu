// features.cu -- neighbourhood-PCA feature kernels over a CSR (nn, nn_ptr) neighbour list.
// Replaces compute_geometric_features{,_multiscale,_optimal} (include/pgeof.hpp:75-310)
// and pca_from_neighborhood / pca_from_pointcloud (include/pca.hpp:71-129).
//
// Layout: one CTA owns a tile of kRows consecutive CSR rows.  The tile's slice of `nn`
// is one contiguous span, so it is staged into shared memory with a single 1-D TMA bulk
// copy (cp.async.bulk + mbarrier) when 16-B alignment allows, else with coalesced loads.
// Then ONE THREAD PER ROW walks its neighbour list: gathers xyz (3 x 4-B read-only loads,
// several rows of loads in flight per thread), accumulates the 9 origin-shifted moments
// (origin = the row's first neighbour), solves the 3x3 eigenproblem in
// registers (eig3.cuh) and stages its 11 floats in shared memory; the tile's output is one
// contiguous block written with a TMA bulk store (or coalesced stores).
//
// Algorithmic bytes per row of length k: 4k (nn) + 4 (nn_ptr) + 12k (xyz gather) + 44 (out)
// = 48 + 16k; HBM-bound (random 12-B gathers cost a 32-B sector each).
#include <algorithm>
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
    const uint32_t* nn; uint32_t nnz;
    const uint32_t* nn_ptr; uint32_t n_rows;
    uint32_t k_min; int eig_order;
    float* out;
    int* err;
    uint32_t nn_cap;       // entries of `nn` the shared-memory tile can hold
    int tma_in, tma_out;   // pointer alignment allows bulk copies
    // multiscale
    uint32_t scales[kMaxScalesPerPass]; uint32_t n_scales_pass; uint32_t n_scales_total; uint32_t scale_base;
    // optimal
    uint32_t k_step, k_min_search;
};

struct Moments {
    float sx = 0, sy = 0, sz = 0, sxx = 0, sxy = 0, sxz = 0, syy = 0, syz = 0, szz = 0;
    __device__ __forceinline__ void add(float dx, float dy, float dz)
    {
        sx += dx; sy += dy; sz += dz;
        sxx = fmaf(dx, dx, sxx); sxy = fmaf(dx, dy, sxy); sxz = fmaf(dx, dz, sxz);
        syy = fmaf(dy, dy, syy); syz = fmaf(dy, dz, syz); szz = fmaf(dz, dz, szz);
    }
    // population covariance of the first k points (pca.hpp:75-76), shift invariant
    __device__ __forceinline__ Pca<float> pca(uint32_t k, int eig_order) const
    {
        const float inv = 1.f / (float)k;
        const float mx = sx * inv, my = sy * inv, mz = sz * inv;
        return pca_from_cov<float>(fmaf(-mx, mx, sxx * inv), fmaf(-mx, my, sxy * inv), fmaf(-mx, mz, sxz * inv),
                                   fmaf(-my, my, syy * inv), fmaf(-my, mz, syz * inv), fmaf(-mz, mz, szz * inv), eig_order);
    }
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

// Tile prologue shared by the three kernels: stages nn[p0, p1) of the CTA's rows.
// Returns the tile-relative base (entry p lives at s_nn[p - a0]) or staged = false.
struct Tile {
    uint32_t r0, rows, p0, p1, a0;
    bool staged, ok;
};

__device__ __forceinline__ Tile stage_tile(const FeatArgs& a, uint32_t* s_nn, uint64_t* bar)
{
    Tile t;
    t.r0 = blockIdx.x * kRows;
    t.rows = min((uint32_t)kRows, a.n_rows - t.r0);
    t.p0 = __ldg(a.nn_ptr + t.r0);
    t.p1 = __ldg(a.nn_ptr + t.r0 + t.rows);
    t.ok = t.p0 <= t.p1 && t.p1 <= a.nnz;          // corrupt nn_ptr -> PGEOF_EINDEX, rows left 0
    t.a0 = t.p0 & ~3u;
    t.staged = t.ok && (t.p1 - t.a0) <= a.nn_cap;
    if (!t.ok) { if (threadIdx.x == 0) atomicExch(a.err, 1); return t; }
    if (!t.staged) return t;
    const uint32_t a1 = max(t.p1 & ~3u, t.a0);     // end of the 16-B aligned body
    if (a.tma_in && a1 > t.a0) {
        if (threadIdx.x == 0) {
            ptx::mbarrier_init(bar, 1);
            ptx::fence_mbarrier_init();
            const uint32_t bytes = (a1 - t.a0) * 4u;
            ptx::mbarrier_arrive_expect_tx(bar, bytes);
            ptx::bulk_g2s(s_nn, a.nn + t.a0, bytes, bar);
        }
        for (uint32_t p = a1 + threadIdx.x; p < t.p1; p += kRows) s_nn[p - t.a0] = __ldg(a.nn + p);   // <= 3 entries
        __syncthreads();                      // barrier init visible to the waiters + tail stored
        ptx::mbarrier_wait(bar, 0);
    } else {
        for (uint32_t p = t.p0 + threadIdx.x; p < t.p1; p += kRows) s_nn[p - t.a0] = __ldg(a.nn + p);
        __syncthreads();
    }
    return t;
}

__device__ __forceinline__ void store_tile(const FeatArgs& a, const Tile& t, const float* s_out, uint32_t floats_per_row, float* gdst)
{
    __syncthreads();
    const uint32_t total = t.rows * floats_per_row;
    if (a.tma_out && (total & 3u) == 0) {
        if (threadIdx.x == 0) {
            ptx::fence_proxy_async_smem();
            ptx::bulk_s2g(gdst, s_out, total * 4u);
            ptx::bulk_commit();
            ptx::bulk_wait_read0();
        }
    } else {
        for (uint32_t i = threadIdx.x; i < total; i += kRows) gdst[i] = s_out[i];
    }
}

// row walker: calls fn(j, dx, dy, dz) for neighbour j = 0..len-1 with coordinates
// relative to the first neighbour; returns false when an index is out of range.
template <typename NnPtr, typename Fn>
__device__ __forceinline__ bool walk_row(const FeatArgs& a, NnPtr src, uint32_t len, Fn&& fn)
{
    const uint32_t i0 = src[0];
    if (i0 >= a.n_xyz) return false;
    const float px = __ldg(a.xyz + 3 * (size_t)i0), py = __ldg(a.xyz + 3 * (size_t)i0 + 1), pz = __ldg(a.xyz + 3 * (size_t)i0 + 2);
    bool ok = true;
#pragma unroll 4
    for (uint32_t j = 0; j < len; ++j) {
        uint32_t i = src[j];
        if (i >= a.n_xyz) { ok = false; i = i0; }
        const float x = __ldg(a.xyz + 3 * (size_t)i), y = __ldg(a.xyz + 3 * (size_t)i + 1), z = __ldg(a.xyz + 3 * (size_t)i + 2);
        fn(j, x - px, y - py, z - pz);
    }
    return ok;
}

struct GlobalNn {   // fallback when the tile does not fit shared memory
    const uint32_t* p;
    __device__ __forceinline__ uint32_t operator[](uint32_t j) const { return __ldg(p + j); }
};

// ----------------------------------------------------------------------------------
// compute_features (pgeof.hpp:75-117)
// ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRows) features_kernel(const FeatArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    float* s_out = reinterpret_cast<float*>(smem + 128);
    uint32_t* s_nn = reinterpret_cast<uint32_t*>(smem + 128 + kRows * 11 * sizeof(float));
    const Tile t = stage_tile(a, s_nn, bar);
    float f[11];
#pragma unroll
    for (int i = 0; i < 11; ++i) f[i] = 0.f;
    if (t.ok && threadIdx.x < t.rows) {
        const uint32_t b = __ldg(a.nn_ptr + t.r0 + threadIdx.x), e = __ldg(a.nn_ptr + t.r0 + threadIdx.x + 1);
        if (e < b || e > a.nnz) atomicExch(a.err, 1);
        else {
            const uint32_t len = e - b;
            if (len >= a.k_min && len > 0) {          // pgeof.hpp:103
                Moments m;
                auto acc = [&](uint32_t, float dx, float dy, float dz) { m.add(dx, dy, dz); };
                const bool ok = t.staged ? walk_row(a, s_nn + (b - t.a0), len, acc) : walk_row(a, GlobalNn{a.nn + b}, len, acc);
                if (!ok) atomicExch(a.err, 2);
                else features11<float>(m.pca(len, a.eig_order), f);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 11; ++i) s_out[threadIdx.x * 11 + i] = f[i];
    store_tile(a, t, s_out, 11, a.out + (size_t)t.r0 * 11);
}

// ----------------------------------------------------------------------------------
// compute_features_multiscale (pgeof.hpp:159-211): scale s uses the first k_s entries of
// the row; one walk yields every scale from the running (prefix) moments.
// ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRows) multiscale_kernel(const FeatArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    uint32_t* s_nn = reinterpret_cast<uint32_t*>(smem + 128);
    const Tile t = stage_tile(a, s_nn, bar);
    if (!t.ok || threadIdx.x >= t.rows) return;
    const uint32_t row = t.r0 + threadIdx.x;
    const uint32_t b = __ldg(a.nn_ptr + row), e = __ldg(a.nn_ptr + row + 1);
    if (e < b || e > a.nnz) { atomicExch(a.err, 1); return; }
    const uint32_t len = e - b;
    // rows are only walked up to the largest scale of this pass that fits (pgeof.hpp:193 early break)
    uint32_t n_fit = 0;
    while (n_fit < a.n_scales_pass && a.scales[n_fit] <= len) ++n_fit;
    if (n_fit == 0 || a.scales[n_fit - 1] == 0) return;
    const uint32_t walk = a.scales[n_fit - 1];
    float* out = a.out + ((size_t)row * a.n_scales_total + a.scale_base) * 11;
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
    const bool ok = t.staged ? walk_row(a, s_nn + (b - t.a0), walk, acc) : walk_row(a, GlobalNn{a.nn + b}, walk, acc);
    if (!ok) {
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
    float* s_out = reinterpret_cast<float*>(smem + 128);
    uint32_t* s_nn = reinterpret_cast<uint32_t*>(smem + 128 + kRows * 12 * sizeof(float));
    const Tile t = stage_tile(a, s_nn, bar);
    float f[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) f[i] = 0.f;
    if (t.ok && threadIdx.x < t.rows) {
        const uint32_t b = __ldg(a.nn_ptr + t.r0 + threadIdx.x), e = __ldg(a.nn_ptr + t.r0 + threadIdx.x + 1);
        if (e < b || e > a.nnz) atomicExch(a.err, 1);
        else {
            const uint32_t len = e - b;
            if (len >= a.k_min && len >= a.k_min_search && len > 0) {               // pgeof.hpp:272
                const uint32_t k0 = min(max(max(a.k_min, a.k_min_search), 1u), len);   // :274
                MomentsD m;
                double best_h = 1.0, best_c[6] = {0, 0, 0, 0, 0, 0};
                uint32_t best_k = len;
                auto acc = [&](uint32_t j, float dx, float dy, float dz) {
                    m.add((double)dx, (double)dy, (double)dz);
                    const uint32_t k = j + 1;
                    if (k < k0) return;
                    if (k > k0 && (k % a.k_step) != 0 && k != len) return;          // :283
                    double c[6], w[3];
                    m.cov(k, c);
                    eigvals3_f64(c[0], c[1], c[2], c[3], c[4], c[5], w);
                    const double h = eigentropy_of<double>(w[0], w[1], w[2]);
                    if (k == k0 || h < best_h) {                                    // :289
                        best_h = h; best_k = k;
#pragma unroll
                        for (int i = 0; i < 6; ++i) best_c[i] = c[i];
                    }
                };
                const bool ok = t.staged ? walk_row(a, s_nn + (b - t.a0), len, acc) : walk_row(a, GlobalNn{a.nn + b}, len, acc);
                if (!ok) atomicExch(a.err, 2);
                else {
                    float g[11];
                    features11<float>(pca_from_cov<float>((float)best_c[0], (float)best_c[1], (float)best_c[2], (float)best_c[3],
                                                         (float)best_c[4], (float)best_c[5], a.eig_order), g);
#pragma unroll
                    for (int i = 0; i < 11; ++i) f[i] = g[i];
                    f[11] = (float)best_k;
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) s_out[threadIdx.x * 12 + i] = f[i];
    store_tile(a, t, s_out, 12, a.out + (size_t)t.r0 * 12);
}

// shared-memory tile for `nn`: mean row length with 50 % head-room, at least 32 entries a row
uint32_t pick_nn_cap(size_t nnz, size_t n_rows, size_t fixed_bytes)
{
    const double mean = n_rows ? (double)nnz / (double)n_rows : 0.0;
    size_t cap = (size_t)(kRows * std::max(32.0, mean * 1.5)) + 8;
    const size_t max_bytes = 96 * 1024 - fixed_bytes;          // keep >= 2 CTAs per SM
    cap = std::min(cap, max_bytes / 4);
    return (uint32_t)(cap & ~(size_t)3);
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
                 uint32_t k_min, int eig_order, float* out, cudaStream_t stream)
{
    if (n_rows == 0) return PGEOF_OK;
    DeviceBuffer err;
    PGEOF_TRY(err.alloc(sizeof(int), stream));
    PGEOF_CUDA(cudaMemsetAsync(err.ptr, 0, sizeof(int), stream));
    FeatArgs a;
    PGEOF_TRY(make_args(&a, xyz, n_xyz, nn, nnz, nn_ptr, n_rows, eig_order, out, err.as<int>(), 11));
    a.k_min = k_min;
    const size_t fixed = 128 + kRows * 11 * sizeof(float);
    a.nn_cap = pick_nn_cap(nnz, n_rows, fixed);
    PGEOF_TRY(launch_tiles(features_kernel, "features", a, fixed + (size_t)a.nn_cap * 4, stream));
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
    PGEOF_CUDA(cudaMemsetAsync(out, 0, n_rows * n_scales * 11 * sizeof(float), stream));   // calloc semantics, pgeof.hpp:175
    FeatArgs a;
    PGEOF_TRY(make_args(&a, xyz, n_xyz, nn, nnz, nn_ptr, n_rows, eig_order, out, err.as<int>(), 11));
    a.n_scales_total = (uint32_t)n_scales;
    const size_t fixed = 128;
    a.nn_cap = pick_nn_cap(nnz, n_rows, fixed);
    for (size_t base = 0; base < n_scales; base += kMaxScalesPerPass) {
        a.scale_base = (uint32_t)base;
        a.n_scales_pass = (uint32_t)std::min<size_t>(kMaxScalesPerPass, n_scales - base);
        for (uint32_t s = 0; s < a.n_scales_pass; ++s) a.scales[s] = k_scales_host[base + s];
        PGEOF_TRY(launch_tiles(multiscale_kernel, "multiscale", a, fixed + (size_t)a.nn_cap * 4, stream));
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
    const size_t fixed = 128 + kRows * 12 * sizeof(float);
    a.nn_cap = pick_nn_cap(nnz, n_rows, fixed);
    PGEOF_TRY(launch_tiles(optimal_kernel, "optimal", a, fixed + (size_t)a.nn_cap * 4, stream));
    return device_flag_check(err.as<int>(), stream, "compute_features_optimal");
}

}  // namespace pgeof
