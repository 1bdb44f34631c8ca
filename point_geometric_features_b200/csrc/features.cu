// features.cu -- neighbourhood-PCA feature kernels over a CSR (nn, nn_ptr) neighbour list.
// Replaces compute_geometric_features{,_multiscale,_optimal} (include/pgeof.hpp:75-310)
// and pca_from_neighborhood / pca_from_pointcloud (include/pca.hpp:71-129).
//
// Layout.  Rows are processed in SPATIAL order: a counting sort of the rows by the Morton code of the coarse
// cell of their first neighbour makes consecutive threads / CTAs work on overlapping neighbourhoods, so the gathers
// of a row hit lines its spatial neighbours just pulled into L1 / L2.  The cloud is re-packed once into 16-B float4
// records so that a gather is ONE 128-bit load.  One CTA owns a tile of rows, ONE THREAD PER ROW: every thread
// streams its own slice of `nn` with 256-bit no-allocate loads (the stream is read exactly once and must not crowd
// the gathered points out of L1), keeps 8 gathers in flight, accumulates the 9 origin-shifted moments (origin = the
// row's first neighbour), solves the 3x3 eigenproblem in registers (eig3.cuh) and the tile's features
// leave through shared memory as 44-B row segments (permuted rows) or one TMA bulk store.
//
// What bounds it (profiles/r2_summary.md): not HBM.  The cloud is in input order, so every gathered point sits in its own
// 128-B line: a 512-row CTA touches ~1950 distinct lines, two CTAs need 3900 L1 tags where ~1500 exist, 38 % of the 500 M
// gathers of a 10 M x 50 launch miss L1 (3.8e8 L2 -> L1 sectors) and the kernel runs at the rate the SMs keep those misses in
// flight (long_scoreboard: 17 stalled warps per issue); its time does not move when the DRAM traffic is changed by 60 % with
// L2 policies.  A Morton-sorted copy of the cloud behind a rank table (rank[nn[j]] then the record) was built and measured in
// round 2: DRAM reads 6.1 -> 5.6 GB, but a table lookup has the same one-line-per-lookup footprint as the gather it
// replaces -- same time, +0.3 ms of pre-pass -- and was dropped.
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

constexpr int kRows = 128;              // rows per CTA of the optimal-k kernel
constexpr int kMaxScalesPerPass = 8;

struct FeatArgs {
    uint32_t n_xyz;
    const float4* pts;            // the cloud as 16-B float4 records (input order): one LDG.128 per gather
    cudaTextureObject_t tex;      // the same records behind a 1-D linear texture (0: none): the gathers of the walkers go through the TEX pipe
    const uint32_t* order;        // spatial row permutation, or nullptr = identity
    const uint32_t* nn; unsigned long long nnz;
    unsigned long long nn_lo;              // rows may only address nn[nn_lo, nnz): a host pipeline hands the kernels one slice of nn at a time
    const uint32_t* nn_ptr;                // row offsets, uint32 (the reference's dtype) ...
    const unsigned long long* nn_ptr64;    // ... or uint64 (extension: more than 2^32-1 neighbours in one CSR, README "known limitations")
    uint32_t n_rows;
    uint32_t k_min; int eig_order;
    float* out;
    int* err;
    int tma_out;           // pointer alignment allows bulk stores
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
        const double inv = __drcp_rn((double)k);          // = 1.0 / k, correctly rounded, without the division's slow-path checks
        const double mx = sx * inv, my = sy * inv, mz = sz * inv;
        c[0] = sxx * inv - mx * mx; c[1] = sxy * inv - mx * my; c[2] = sxz * inv - mx * mz;
        c[3] = syy * inv - my * my; c[4] = syz * inv - my * mz; c[5] = szz * inv - mz * mz;
    }
};

struct Tile {
    uint32_t r0, rows;   // positions [r0, r0 + rows) of the (permuted) row sequence
    bool bulk_store;     // identity order: the tile's output is one contiguous block
};

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


// eight consecutive entries of nn with ONE 256-bit load that does not allocate in L1 (sm_100 LDG.256): the stream
// is read exactly once, and 1024 resident rows x one 128-B line each would otherwise crowd the gathered
// points out of L1.  p must be 32-byte aligned.
__device__ __forceinline__ void stream_nn8(const uint32_t* p, uint32_t (&i)[8])
{
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(i[0]), "=r"(i[1]), "=r"(i[2]), "=r"(i[3]), "=r"(i[4]), "=r"(i[5]), "=r"(i[6]), "=r"(i[7]) : "l"(p));
}

// nn[b, e) of a row, whichever width the offsets have
__device__ __forceinline__ void row_span(const FeatArgs& a, uint32_t row, unsigned long long& b, unsigned long long& e)
{
    if (a.nn_ptr64) { b = __ldg(a.nn_ptr64 + row); e = __ldg(a.nn_ptr64 + row + 1); }
    else { b = __ldg(a.nn_ptr + row); e = __ldg(a.nn_ptr + row + 1); }
}

// one gathered point: through the texture units when the call made a texture of the cloud.  A warp's 32 scattered 16-byte
// records are 32 wavefronts of the LSU data pipe, which the gather kernels run at 60-74 % of; as texels they cost the
// feature / multiscale / optimal kernels 2 / 5 / 5 % less time (alternating the two pipes: 7 % MORE -- what the gathers wait
// for is the miss path behind L1, not a pipe; profiles/r2_summary.md).  PGEOF_FEATURES_TEX = 0 keeps the LDG path.
__device__ __forceinline__ float4 fetch_pt(const FeatArgs& a, uint32_t i)
{
    if (a.tex) return tex1Dfetch<float4>(a.tex, (int)i);
    return __ldg(a.pts + i);
}

// up to N (<= 7) consecutive entries: indices first, then their gathers together, then the moments in order
template <int N, typename Acc>
__device__ __forceinline__ void walk_some(const FeatArgs& a, const uint32_t* __restrict__ p, uint32_t j0, uint32_t cnt, uint32_t i0, const float4& o,
                                          Acc& acc, bool& ok)
{
    uint32_t i[N];
    float4 q[N];
#pragma unroll
    for (int u = 0; u < N; ++u) {
        i[u] = (uint32_t)u < cnt ? __ldg(p + u) : i0;
        if (i[u] >= a.n_xyz) { ok = false; i[u] = i0; }
    }
#pragma unroll
    for (int u = 0; u < N; ++u) q[u] = fetch_pt(a, i[u]);
#pragma unroll
    for (int u = 0; u < N; ++u) if ((uint32_t)u < cnt) acc(j0 + u, q[u].x - o.x, q[u].y - o.y, q[u].z - o.z);
}

// acc(j, dx, dy, dz) is called for j = 0 .. len-1 in order, offsets relative to the row's first neighbour
template <typename Acc>
__device__ __forceinline__ bool walk_direct(const FeatArgs& a, unsigned long long b, uint32_t len, Acc& acc)
{
    const uint32_t* __restrict__ p = a.nn + b;
    const uint32_t n = a.n_xyz;
    const uint32_t i0 = __ldg(p);
    if (i0 >= n) return false;
    const float4 o = fetch_pt(a, i0);   // origin of the shifted moments; its own term is zero
    bool ok = true;
    acc(0u, 0.f, 0.f, 0.f);
    uint32_t j = 1;
    // head: up to the next 32-byte boundary of the stream (whatever the alignment of nn itself)
    const uint32_t nh = min((uint32_t)((0u - (uint32_t)reinterpret_cast<uintptr_t>(p + 1)) & 31u) >> 2, len - 1u);
    if (nh) { walk_some<7>(a, p + j, j, nh, i0, o, acc, ok); j += nh; }
    for (; j + 8 <= len; j += 8) {
        uint32_t i[8];
        stream_nn8(p + j, i);
        float4 q[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) if (i[u] >= n) { ok = false; i[u] = i0; }
#pragma unroll
        for (int u = 0; u < 8; ++u) q[u] = fetch_pt(a, i[u]);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc(j + u, q[u].x - o.x, q[u].y - o.y, q[u].z - o.z);
    }
    if (j < len) walk_some<7>(a, p + j, j, len - j, i0, o, acc, ok);
    return ok;
}

// THREADS rows per tile, one thread per row.  Every CTA owns a CONTIGUOUS chunk of the (spatially ordered) row
// sequence and walks it tile by tile: the rows an SM works on at any time form one or two compact blobs, so a
// gathered point is re-used out of L1 by the blob's other rows (small independent CTAs spread over the whole
// in-flight window shared their gathers only through L2).
// PHASES = 2: the tile's rows are staged and written half at a time, so two resident CTAs need 25 KB of shared memory instead
// of 49 KB and the 32 KB carve-out leaves 196 KB of L1 to the gathers
template <int THREADS, int PHASES = 1>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS) features_direct_kernel(const FeatArgs a)
{
    constexpr int STAGED = THREADS / PHASES;
    __shared__ uint32_t s_rowid[STAGED];
    __shared__ __align__(128) float s_out[STAGED * 11];   // source of a TMA bulk store: 16-B alignment required
    const uint32_t tiles = (a.n_rows + THREADS - 1) / THREADS;
    const uint32_t per = (tiles + gridDim.x - 1) / gridDim.x;
    const uint32_t t_end = min(tiles, (blockIdx.x + 1) * per);
    for (uint32_t tile = blockIdx.x * per; tile < t_end; ++tile) {
        const uint32_t r0 = tile * THREADS;
        Tile t{r0, min((uint32_t)THREADS, a.n_rows - r0), a.order == nullptr && a.out_rows == nullptr};
        float f[11];
#pragma unroll
        for (int i = 0; i < 11; ++i) f[i] = 0.f;
        uint32_t row = r0 + threadIdx.x;
        if (threadIdx.x < t.rows && a.order) row = __ldg(a.order + row);
        const uint32_t out_row = (a.out_rows && threadIdx.x < t.rows) ? __ldg(a.out_rows + row) : row;
        if (threadIdx.x < t.rows) {
            unsigned long long b, e;
            row_span(a, row, b, e);
            if (e < b || e > a.nnz || b < a.nn_lo || e - b > 0xffffffffull) atomicExch(a.err, 1);   // corrupt nn_ptr -> PGEOF_EINDEX, row left 0
            else if (e - b >= a.k_min && e > b) {                // pgeof.hpp:103
                Moments m;
                auto acc = [&](uint32_t, float dx, float dy, float dz) { m.add(dx, dy, dz); };
                if (!walk_direct(a, b, (uint32_t)(e - b), acc)) atomicExch(a.err, 2);
                else features11<float>(m.pca((uint32_t)(e - b), a.eig_order), f);
            }
        }
#pragma unroll
        for (int ph = 0; ph < PHASES; ++ph) {
            if ((int)threadIdx.x / STAGED == ph) {
                const uint32_t slot = threadIdx.x - ph * STAGED;
                s_rowid[slot] = out_row;
#pragma unroll
                for (int i = 0; i < 11; ++i) s_out[slot * 11 + i] = f[i];
            }
            const uint32_t done = ph * STAGED;
            Tile th{r0 + done, t.rows > done ? min((uint32_t)STAGED, t.rows - done) : 0u, t.bulk_store};
            store_rows<11>(a, th, s_out, s_rowid);
            __syncthreads();                                     // the staging buffers are re-used by the next phase / tile
        }
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
    unsigned long long b, e;
    row_span(a, row, b, e);
    uint32_t n_fit = 0, s = 0;                                  // scales of this pass the row is long enough for / written so far
    if (e < b || e > a.nnz || b < a.nn_lo || e - b > 0xffffffffull) atomicExch(a.err, 1);
    else {
        const uint32_t len = (uint32_t)(e - b);
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
            if (!walk_direct(a, b, a.scales[n_fit - 1], acc)) { atomicExch(a.err, 2); s = 0; }
        }
    }
    for (; s < a.n_scales_pass; ++s) store11(out + s * 11, zero);
}

// compute_features_multiscale in two passes.  The kernel above waits on its gathers (ncu r2ab_ms: long_scoreboard 6 stalled
// warps per issue) while 73 % of its instructions are the eigen solves, which its 80 registers pay for with a quarter less
// occupancy: gathering and solving do not overlap well inside one thread.  Pass 1 is the walker alone (moments only, 32 warps
// per SM) and parks the nine prefix moments of every (row, scale) in scratch (36 B); pass 2 is one thread per (row, scale):
// load, solve, 11 formulas, store -- no gathers, no divergence between scales.  Same operations on the same values as the
// one-pass kernel: bit identical.  Scratch: 36 B x rows x scales per chunk of rows (<= 1.5 GB).
__global__ void __launch_bounds__(kMsThreads, 4) multiscale_moments_kernel(const FeatArgs a, float* __restrict__ mom, uint32_t pos0, uint32_t pos1)
{
    const uint32_t pos = pos0 + blockIdx.x * kMsThreads + threadIdx.x;      // position in the (spatially ordered) row sequence
    if (pos >= pos1) return;
    const uint32_t row = a.order ? __ldg(a.order + pos) : pos;
    float* const dst = mom + (size_t)(pos - pos0) * a.n_scales_pass * 9;
    unsigned long long b, e;
    row_span(a, row, b, e);
    if (e < b || e > a.nnz || b < a.nn_lo || e - b > 0xffffffffull) { atomicExch(a.err, 1); return; }
    const uint32_t len = (uint32_t)(e - b);
    uint32_t n_fit = 0, s = 0;
    while (n_fit < a.n_scales_pass && a.scales[n_fit] <= len) ++n_fit;
    if (!n_fit || a.scales[n_fit - 1] == 0) return;
    while (s < n_fit && a.scales[s] == 0) ++s;
    Moments m;
    uint32_t next_k = s < n_fit ? a.scales[s] : 0xffffffffu;
    auto acc = [&](uint32_t j, float dx, float dy, float dz) {
        m.add(dx, dy, dz);
        while (j + 1 == next_k) {
            float* d = dst + s * 9;
            d[0] = m.sx; d[1] = m.sy; d[2] = m.sz; d[3] = m.sxx; d[4] = m.sxy; d[5] = m.sxz; d[6] = m.syy; d[7] = m.syz; d[8] = m.szz;
            ++s;
            next_k = s < n_fit ? a.scales[s] : 0xffffffffu;
        }
    };
    if (!walk_direct(a, b, a.scales[n_fit - 1], acc)) atomicExch(a.err, 2);
}

__global__ void __launch_bounds__(256) multiscale_eigen_kernel(const FeatArgs a, const float* __restrict__ mom, uint32_t pos0, uint32_t pos1)
{
    const size_t item = (size_t)blockIdx.x * 256 + threadIdx.x;
    const uint32_t rel = (uint32_t)(item / a.n_scales_pass), s = (uint32_t)(item - (size_t)rel * a.n_scales_pass);
    if (rel >= pos1 - pos0) return;
    const uint32_t pos = pos0 + rel;
    const uint32_t row = a.order ? __ldg(a.order + pos) : pos;
    float f[11];
#pragma unroll
    for (int i = 0; i < 11; ++i) f[i] = 0.f;
    unsigned long long b, e;
    row_span(a, row, b, e);
    const uint32_t k = a.scales[s];
    if (!(e < b || e > a.nnz || b < a.nn_lo || e - b > 0xffffffffull) && k > 0 && k <= e - b && *reinterpret_cast<volatile int*>(a.err) == 0) {
        const float* src = mom + item * 9;                                   // pgeof.hpp:175,193: rows too short for a scale stay 0
        Moments m;
        m.sx = __ldg(src); m.sy = __ldg(src + 1); m.sz = __ldg(src + 2); m.sxx = __ldg(src + 3); m.sxy = __ldg(src + 4); m.sxz = __ldg(src + 5);
        m.syy = __ldg(src + 6); m.syz = __ldg(src + 7); m.szz = __ldg(src + 8);
        features11<float>(m.pca(k, a.eig_order), f);
    }
    store11(a.out + ((size_t)row * a.n_scales_total + a.scale_base + s) * 11, f);
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
    __shared__ __align__(128) float s_out[kRows * 12];     // source of a TMA bulk store: 16-B alignment required
    const uint32_t r0 = blockIdx.x * kRows;
    Tile t{r0, min((uint32_t)kRows, a.n_rows - r0), a.order == nullptr};
    float f[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) f[i] = 0.f;
    uint32_t row = r0 + threadIdx.x;
    if (threadIdx.x < t.rows && a.order) row = __ldg(a.order + row);
    s_rowid[threadIdx.x] = row;
    if (threadIdx.x < t.rows) {
        unsigned long long b, e;
        row_span(a, row, b, e);
        if (e < b || e > a.nnz || b < a.nn_lo || e - b > 0xffffffffull) atomicExch(a.err, 1);
        else {
            const uint32_t len = (uint32_t)(e - b);
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
                if (!walk_direct(a, b, len, acc)) atomicExch(a.err, 2);
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

// ----------------------------------------------------------------------------------
// compute_features_optimal, second version: one ROLLED scan loop per thread, cheap float filter.
// The walker above inlines its accumulator at 22 sites, which forced the ~600-instruction evaluation out of line with its
// state in local memory.  Here the offsets of 8 gathered neighbours go through shared memory ([slot][thread]: conflict
// free) and one rolled loop consumes them, so the evaluation exists once, inline, with all state in registers:
//   * prefix moments in double (as before: the covariance of every k is exact to ~1e-16 relative);
//   * the FILTER entropy comes from float eigenvalues of a cyclic Jacobi iteration with approximate rotations
//     (profiles/r2_summary.md: the IEEE divisions / square roots of the rotations were 68 % of the kernel's instructions);
//   * a candidate within kOptMargin of the best is decided by the double closed-form entropies exactly as the all-double
//     scan decides (strict '<', smallest k wins, pgeof.hpp:289): k_opt is identical.
// ----------------------------------------------------------------------------------
__device__ __noinline__ double entropy_f64_nv(double c0, double c1, double c2, double c3, double c4, double c5)
{
    const double c[6] = {c0, c1, c2, c3, c4, c5};
    return entropy_f64(c);
}

// filter entropy in float with the fast log (|error| ~1e-7: three orders below kOptMargin)
__device__ __forceinline__ float eigentropy_fast(float l0, float l1, float l2)
{
    const float eps = 1e-3f;
    const float inv = __frcp_rn(l0 + l1 + l2 + eps);
    const float e0 = l0 * inv, e1 = l1 * inv, e2 = l2 * inv;
    return -0.69314718056f * (e0 * __log2f(e0 + eps) + e1 * __log2f(e1 + eps) + e2 * __log2f(e2 + eps));
}

// Filter eigenvalues in closed form (trigonometric solution of the characteristic cubic, float, MUFU cosine): ~90 instructions
// against ~250 for three Jacobi sweeps.  The form loses digits where two eigenvalues coincide (phi = acos(r) / 3 is ill
// conditioned at |r| -> 1), so it returns a bound on the entropy error it may have caused and the caller widens its
// too-close-to-call window by it (the double evaluation then decides, as for any near tie): k_opt stays exact.
// Input scaled to max |a_ij| = 1, so the trace is >= 1; w is unsorted and clamped at 0.
__device__ __forceinline__ float eigvals3_closed_f32(float a00, float a01, float a02, float a11, float a12, float a22, float (&w)[3])
{
    const float q = (a00 + a11 + a22) * (1.f / 3.f);
    const float b00 = a00 - q, b11 = a11 - q, b22 = a22 - q;
    const float off2 = fmaf(a01, a01, fmaf(a02, a02, a12 * a12));
    const float p2 = fmaf(b00, b00, fmaf(b11, b11, fmaf(b22, b22, 2.f * off2)));
    if (!(p2 > 1e-30f)) { w[0] = w[1] = w[2] = fmaxf(q, 0.f); return 1e-6f; }
    const float ip = rsqrtf(p2 * (1.f / 6.f));
    const float p = p2 * (1.f / 6.f) * ip;
    const float c00 = b00 * ip, c11 = b11 * ip, c22 = b22 * ip, c01 = a01 * ip, c02 = a02 * ip, c12 = a12 * ip;
    float r = 0.5f * (c00 * fmaf(c11, c22, -c12 * c12) - c01 * fmaf(c01, c22, -c12 * c02) + c02 * fmaf(c01, c12, -c11 * c02));
    r = fminf(1.f, fmaxf(-1.f, r));
    const float phi = acosf(r) * (1.f / 3.f);
    const float hi = fmaf(2.f * p, __cosf(phi), q);
    const float lo = fmaf(2.f * p, __cosf(phi + 2.0943951f), q);
    const float mid = 3.f * q - hi - lo;
    w[0] = fmaxf(lo, 0.f); w[1] = fmaxf(mid, 0.f); w[2] = fmaxf(hi, 0.f);
    // |dr| <= 4e-6 (the determinant of entries <= sqrt(6) in float, rsqrt and cancellation included, four times over);
    // dphi = dr / (3 sqrt(1 - r^2)), never more than sqrt(2 dr) / 3; an eigenvalue moves by <= 2 p (dphi + cosine error 5e-7)
    // plus the rounding of the inputs (1e-6 of the unit norm); the entropy by <= 3 (|ln 1e-3| + 1) = 24 times that over the
    // trace (>= 1).
    const float dr = 4e-6f;
    const float dphi = fminf(dr * rsqrtf(fmaxf(fmaf(-r, r, 1.f), 1e-12f)), 2.9e-3f) * (1.f / 3.f);
    return 24.f * (2.f * p * (dphi + 5e-7f) + 1e-6f);
}

template <int MINB, bool CLOSED>
__global__ void __launch_bounds__(kRows, MINB) optimal_scan_kernel(const FeatArgs a)
{
    __shared__ uint32_t s_rowid[kRows];
    __shared__ __align__(128) float s_out[kRows * 12];     // source of a TMA bulk store: 16-B alignment required
    __shared__ float s_d[24][kRows];                       // offsets of the 8 neighbours in flight: [3 u + axis][thread]
    const uint32_t r0 = blockIdx.x * kRows;
    Tile t{r0, min((uint32_t)kRows, a.n_rows - r0), a.order == nullptr};
    float f[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) f[i] = 0.f;
    uint32_t row = r0 + threadIdx.x;
    if (threadIdx.x < t.rows && a.order) row = __ldg(a.order + row);
    s_rowid[threadIdx.x] = row;
    if (threadIdx.x < t.rows) {
        unsigned long long b, e;
        row_span(a, row, b, e);
        if (e < b || e > a.nnz || b < a.nn_lo || e - b > 0xffffffffull) atomicExch(a.err, 1);
        else {
            const uint32_t len = (uint32_t)(e - b);
            const uint32_t* __restrict__ p = a.nn + b;
            const bool eligible = len >= a.k_min && len >= a.k_min_search && len > 0;     // pgeof.hpp:272
            const uint32_t i0 = eligible ? __ldg(p) : 0u;
            bool ok = i0 < a.n_xyz;
            if (eligible && !ok) atomicExch(a.err, 2);
            if (eligible && ok) {
                const uint32_t k0 = min(max(max(a.k_min, a.k_min_search), 1u), len);    // :274
                const float4 o = fetch_pt(a, i0);
                MomentsD m;
                double best_c[6] = {0, 0, 0, 0, 0, 0}, best_h64 = 0.0;
                float best_h32 = 0.f, best_err = 0.f;
                uint32_t best_k = len;
                bool have64 = false;
                uint32_t rem = 0;                                                        // k % k_step, kept incrementally
                float* const sd = &s_d[0][threadIdx.x];
                uint32_t j = 0;
                while (j < len) {
                    // ---- next group of the nn stream: up to the next 32-byte boundary, then 8 at a time ----
                    uint32_t i[8];
                    const uint32_t lead = (uint32_t)(reinterpret_cast<uintptr_t>(p + j) & 31u) >> 2;
                    uint32_t cnt;
                    if (lead == 0 && j + 8 <= len) { stream_nn8(p + j, i); cnt = 8; }
                    else {
                        cnt = min(8u - lead, len - j);
#pragma unroll
                        for (int u = 0; u < 8; ++u) i[u] = (uint32_t)u < cnt ? __ldg(p + j + u) : i0;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) if (i[u] >= a.n_xyz) { ok = false; i[u] = i0; }
                    float4 q[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) q[u] = fetch_pt(a, i[u]);
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        sd[(3 * u + 0) * kRows] = q[u].x - o.x; sd[(3 * u + 1) * kRows] = q[u].y - o.y; sd[(3 * u + 2) * kRows] = q[u].z - o.z;
                    }
                    // ---- the scan proper: one rolled loop, the evaluation inline ----
#pragma unroll 1
                    for (uint32_t u = 0; u < cnt; ++u) {
                        m.add((double)sd[(3 * u + 0) * kRows], (double)sd[(3 * u + 1) * kRows], (double)sd[(3 * u + 2) * kRows]);
                        const uint32_t k = j + u + 1;
                        rem = rem + 1 == a.k_step ? 0u : rem + 1;
                        if (k < k0 || (k > k0 && rem != 0 && k != len)) continue;        // :283
                        double c[6];
                        m.cov(k, c);
                        // float eigenvalues by cyclic Jacobi from the identity with APPROXIMATE rotations (MUFU reciprocal /
                        // rsqrt: ~20 instructions per rotation instead of ~65 with IEEE division and square root).  A rotation
                        // whose (c, s) is off by a few ulp is a similarity up to a factor 1 + O(1e-7): after <= 24 rotations the
                        // diagonal holds the eigenvalues to ~3e-6 relative, most of it a COMMON scale error that cancels in the
                        // ratios the entropy is made of; an off-diagonal left behind moves an eigenvalue by at most its size
                        // (<= 1e-6 |C|).  |h32 - h64| <= ~5e-5, kOptMargin is 2e-4.
                        float a00 = (float)c[0], a01 = (float)c[1], a02 = (float)c[2], a11 = (float)c[3], a12 = (float)c[4], a22 = (float)c[5];
                        float scale = fmaxf(fmaxf(fabsf(a00), fabsf(a11)), fabsf(a22));
                        scale = fmaxf(scale, fmaxf(fmaxf(fabsf(a01), fabsf(a02)), fabsf(a12)));
                        float w0 = 0.f, w1 = 0.f, w2 = 0.f, err = 0.f;
                        if (CLOSED && scale > 0.f) {
                            const float inv = __frcp_rn(scale);
                            float w[3];
                            err = eigvals3_closed_f32(a00 * inv, a01 * inv, a02 * inv, a11 * inv, a12 * inv, a22 * inv, w);
                            w0 = w[0] * scale; w1 = w[1] * scale; w2 = w[2] * scale;
                        }
                        if (!CLOSED && scale > 0.f) {
                            const float inv = __frcp_rn(scale);
                            a00 *= inv; a01 *= inv; a02 *= inv; a11 *= inv; a12 *= inv; a22 *= inv;
#define PGEOF_ROTF(app, aqq, apq, arp, arq)                                                   \
                            if (fabsf(apq) > fmaxf(1e-9f * (fabsf(app) + fabsf(aqq)), 1e-18f)) {   \
                                /* t = tan(phi) = sign(h apq) 2|apq| / (|h| + sqrt(h^2 + 4 apq^2)): three MUFU ops, no division */ \
                                const float hh = aqq - app, t2a = 2.f * apq;                  \
                                const float r1 = fmaf(hh, hh, t2a * t2a);                     \
                                const float ta = fabsf(t2a) * __frcp_rn(fabsf(hh) + r1 * rsqrtf(r1)); \
                                const float tt = __uint_as_float(__float_as_uint(ta) | ((__float_as_uint(hh) ^ __float_as_uint(apq)) & 0x80000000u)); \
                                const float cs = rsqrtf(fmaf(tt, tt, 1.f)), sn = tt * cs;     \
                                app = fmaf(-tt, apq, app); aqq = fmaf(tt, apq, aqq); apq = 0.f; \
                                const float rp = arp, rq = arq;                               \
                                arp = cs * rp - sn * rq; arq = sn * rp + cs * rq;             \
                            } else apq = 0.f;
#pragma unroll 1
                            for (int sweep = 0; sweep < 8; ++sweep) {
                                const float off = fabsf(a01) + fabsf(a02) + fabsf(a12);
                                if (off <= 1e-6f * (fabsf(a00) + fabsf(a11) + fabsf(a22))) break;
                                PGEOF_ROTF(a00, a11, a01, a02, a12)
                                PGEOF_ROTF(a00, a22, a02, a01, a12)
                                PGEOF_ROTF(a11, a22, a12, a01, a02)
                            }
#undef PGEOF_ROTF
                            w0 = fmaxf(a00 * scale, 0.f); w1 = fmaxf(a11 * scale, 0.f); w2 = fmaxf(a22 * scale, 0.f);
                        }
                        const float h32 = eigentropy_fast(w0, w1, w2);
                        // window inside which float cannot call it: Jacobi values are good to ~5e-5, the closed form to its own bound
                        const float margin = CLOSED ? 1e-4f + err + best_err : kOptMargin;
                        bool take = k == k0 || h32 < best_h32 - margin;
                        bool exact = false;
                        double h64 = 0.0;
                        if (!take && !(h32 > best_h32 + margin)) {                       // too close to call in float
                            if (!have64) { best_h64 = entropy_f64_nv(best_c[0], best_c[1], best_c[2], best_c[3], best_c[4], best_c[5]); have64 = true; }
                            h64 = entropy_f64_nv(c[0], c[1], c[2], c[3], c[4], c[5]);
                            take = h64 < best_h64;                                       // pgeof.hpp:289
                            exact = true;
                        }
                        if (take) {
                            best_k = k; best_h32 = h32; best_h64 = h64; have64 = exact; best_err = err;
#pragma unroll
                            for (int z = 0; z < 6; ++z) best_c[z] = c[z];
                        }
                    }
                    j += cnt;
                }
                if (!ok) atomicExch(a.err, 2);
                else {
                    float g[11];
                    features11<float>(pca_from_cov<float>((float)best_c[0], (float)best_c[1], (float)best_c[2], (float)best_c[3],
                                                         (float)best_c[4], (float)best_c[5], a.eig_order), g);
#pragma unroll
                    for (int z = 0; z < 11; ++z) f[z] = g[z];
                    f[11] = (float)best_k;
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) s_out[threadIdx.x * 12 + i] = f[i];
    store_rows<12>(a, t, s_out, s_rowid);
}
// ----------------------------------------------------------------------------------
// pre-passes: Morton ordering of the cloud (float4 records + rank table) and of the rows
// ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pad_xyz_kernel(const float* __restrict__ xyz, size_t n, float4* __restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_float4(__ldg(xyz + 3 * i), __ldg(xyz + 3 * i + 1), __ldg(xyz + 3 * i + 2), 0.f);
}

struct RowGrid { float lo[3]; float scale[3]; int cells; };

// one warp: reduce the bbox partials and derive the coarse ordering grid (no host sync)
__global__ void row_grid_kernel(const float* __restrict__ partial, int n_partial, int cells, RowGrid* __restrict__ g)
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

// Morton code of the coarse cell of a point: neighbourhoods that overlap are visited close in time along all three
// axes (a (z, y, x) raster order revisits the neighbouring plane only after a whole plane of rows)
__device__ __forceinline__ uint32_t morton_key(const RowGrid& g, float x, float y, float z)
{
    const int c = g.cells;
    const int cx = min(max(__float2int_rd((x - g.lo[0]) * g.scale[0]), 0), c - 1);   // NaN -> 0
    const int cy = min(max(__float2int_rd((y - g.lo[1]) * g.scale[1]), 0), c - 1);
    const int cz = min(max(__float2int_rd((z - g.lo[2]) * g.scale[2]), 0), c - 1);
    return spread3((uint32_t)cx) | (spread3((uint32_t)cy) << 1) | (spread3((uint32_t)cz) << 2);
}

// histogram with warp-aggregated atomics; rank[i] = position of item i inside its bucket
__device__ __forceinline__ uint32_t bucket_rank(uint32_t* __restrict__ counts, uint32_t key, bool valid)
{
    const unsigned active = __ballot_sync(0xffffffffu, valid);
    if (!valid) return 0;
    const unsigned peers = __match_any_sync(active, key);
    const int leader = __ffs(peers) - 1;
    const int lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counts + key, (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    unsigned lt;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt));
    return base + (uint32_t)__popc(peers & lt);
}

// a row sorts with the cell of its first neighbour (itself, for the rows of a self kNN / radius search)
__global__ void __launch_bounds__(256) row_count_kernel(const FeatArgs a, const RowGrid* __restrict__ gp, uint32_t* __restrict__ counts,
                                                        uint32_t* __restrict__ keys, uint32_t* __restrict__ rank)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < a.n_rows;
    uint32_t key = 0;
    if (valid) {
        unsigned long long b, e;
        row_span(a, i, b, e);
        if (e > b && b < a.nnz && b >= a.nn_lo) {
            const uint32_t first = __ldg(a.nn + b);
            if (first < a.n_xyz) {
                const float4 p = __ldg(a.pts + first);
                key = morton_key(*gp, p.x, p.y, p.z);
            }
        }
    }
    const uint32_t r = bucket_rank(counts, key, valid);
    if (valid) { keys[i] = key; rank[i] = r; }
}

__global__ void __launch_bounds__(256) row_scatter_kernel(uint32_t n_rows, const uint32_t* __restrict__ starts, const uint32_t* __restrict__ keys,
                                                          const uint32_t* __restrict__ rank, uint32_t* __restrict__ order)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_rows) order[__ldg(starts + keys[i]) + rank[i]] = i;
}

int env_int(const char* name, int dflt)
{
    const char* e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}

}  // namespace
// set by the host pipeline (capi.cu) around a call that hands the kernels a slice nn[lo, nnz) of the caller's array
thread_local unsigned long long g_nn_window_lo = 0;
namespace {

int make_args(FeatArgs* a, size_t n_xyz, const uint32_t* nn, size_t nnz, RowPtr nn_ptr, size_t n_rows,
              int eig_order, float* out, int* err)
{
    if (n_xyz > 0xffffffffull || n_rows > 0xfffffffeull) { set_error("more than 2^32-1 points or rows"); return PGEOF_EINVAL; }
    if (!nn_ptr.p64 && nnz > 0xffffffffull) { set_error("nn holds more than 2^32-1 entries: uint32 nn_ptr cannot address them (pass uint64 offsets or shard the rows)"); return PGEOF_EINVAL; }
    if (eig_order != PGEOF_EIG_LITERAL && eig_order != PGEOF_EIG_DOCUMENTED) { set_error("bad eig_order %d", eig_order); return PGEOF_EINVAL; }
    std::memset(a, 0, sizeof(*a));
    a->n_xyz = (uint32_t)n_xyz; a->nn = nn; a->nnz = nnz; a->nn_lo = g_nn_window_lo; a->nn_ptr = nn_ptr.p32; a->nn_ptr64 = nn_ptr.p64; a->n_rows = (uint32_t)n_rows;
    a->eig_order = eig_order; a->out = out; a->err = err; a->k_min = 1; a->k_step = 1; a->k_min_search = 1;
    a->tma_out = ((uintptr_t)out % 16 == 0);
    return PGEOF_OK;
}

// Device buffers of the pre-passes; they live until the feature kernel was enqueued (stream-ordered frees).
struct Prepass {
    DeviceBuffer pts, order;
    cudaTextureObject_t tex = 0;
    ~Prepass() { if (tex) cudaDestroyTextureObject(tex); }
};

// PGEOF_FEATURES_SORT = 0 keeps the rows in input order (debugging; small inputs do that anyway)
int prepare(FeatArgs* a, const float* xyz, Prepass* p, cudaStream_t stream)
{
    const uint32_t n = a->n_xyz;
    // 1. float4 re-pack of the cloud: one 128-bit load per gathered neighbour
    PGEOF_TRY(p->pts.alloc((size_t)std::max<uint32_t>(n, 1) * sizeof(float4), stream));
    a->pts = p->pts.as<float4>();
    if (n) {
        pad_xyz_kernel<<<(n + 255) / 256, 256, 0, stream>>>(xyz, n, p->pts.as<float4>());
        PGEOF_LAUNCH_CHECK();
    }
    a->tex = 0;
    if (n && n <= (1u << 27) && env_int("PGEOF_FEATURES_TEX", 1) != 0) {    // 1-D linear textures hold 2^27 texels
        cudaResourceDesc rd{};
        rd.resType = cudaResourceTypeLinear;
        rd.res.linear.devPtr = p->pts.ptr;
        rd.res.linear.desc = cudaCreateChannelDesc<float4>();
        rd.res.linear.sizeInBytes = (size_t)n * sizeof(float4);
        cudaTextureDesc td{};
        td.readMode = cudaReadModeElementType;
        if (cudaCreateTextureObject(&p->tex, &rd, &td, nullptr) == cudaSuccess) a->tex = p->tex;
        else { cudaGetLastError(); p->tex = 0; }
    }
    // 2. spatial row order (counting sort by the Morton cell of the first neighbour)
    const int min_rows = env_int("PGEOF_FEATURES_SORT_MIN_ROWS", 32768);
    if (n == 0 || a->nnz == 0 || (int64_t)a->n_rows < (int64_t)min_rows || env_int("PGEOF_FEATURES_SORT", 1) == 0) return PGEOF_OK;
    KernelTimer timer("row_order", stream);
    DeviceBuffer partial, grid, counts, keys, rrank;
    int n_partial = 0;
    PGEOF_TRY(bbox_partials(xyz, n, &partial, &n_partial, stream));
    int cells = (int)std::lround(std::cbrt((double)a->n_rows / 6.0));
    cells = std::min(std::max(cells, 8), 128);
    int bits = 3;
    while ((1 << bits) < cells) ++bits;
    const size_t n_cells = (size_t)1 << (3 * bits);
    PGEOF_TRY(grid.alloc(sizeof(RowGrid), stream));
    PGEOF_TRY(counts.alloc((n_cells + 1) * sizeof(uint32_t), stream));
    PGEOF_TRY(keys.alloc((size_t)a->n_rows * sizeof(uint32_t), stream));
    PGEOF_TRY(rrank.alloc((size_t)a->n_rows * sizeof(uint32_t), stream));
    PGEOF_TRY(p->order.alloc((size_t)a->n_rows * sizeof(uint32_t), stream));
    row_grid_kernel<<<1, 32, 0, stream>>>(partial.as<float>(), n_partial, cells, grid.as<RowGrid>());
    PGEOF_LAUNCH_CHECK();
    PGEOF_CUDA(cudaMemsetAsync(counts.ptr, 0, (n_cells + 1) * sizeof(uint32_t), stream));
    const unsigned rblocks = (a->n_rows + 255) / 256;
    row_count_kernel<<<rblocks, 256, 0, stream>>>(*a, grid.as<RowGrid>(), counts.as<uint32_t>(), keys.as<uint32_t>(), rrank.as<uint32_t>());
    PGEOF_LAUNCH_CHECK();
    PGEOF_TRY(exclusive_scan_u32(counts.as<uint32_t>(), n_cells, stream));
    row_scatter_kernel<<<rblocks, 256, 0, stream>>>(a->n_rows, counts.as<uint32_t>(), keys.as<uint32_t>(), rrank.as<uint32_t>(), p->order.as<uint32_t>());
    PGEOF_LAUNCH_CHECK();
    a->order = p->order.as<uint32_t>();
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

int features_run(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz, RowPtr nn_ptr, size_t n_rows,
                 uint32_t k_min, int eig_order, float* out, cudaStream_t stream, const uint32_t* out_rows)
{
    if (n_rows == 0) return PGEOF_OK;
    DeviceBuffer err;
    PGEOF_TRY(err.alloc(sizeof(int), stream));
    PGEOF_CUDA(cudaMemsetAsync(err.ptr, 0, sizeof(int), stream));
    FeatArgs a;
    PGEOF_TRY(make_args(&a, n_xyz, nn, nnz, nn_ptr, n_rows, eig_order, out, err.as<int>()));
    a.k_min = k_min;
    a.out_rows = out_rows;
    Prepass pre;
    PGEOF_TRY(prepare(&a, xyz, &pre, stream));
    const int cta = env_int("PGEOF_FEATURES_CTA", 512);
    auto launch = [&](auto kern, int threads) -> int {
        const unsigned tiles = (unsigned)((n_rows + threads - 1) / threads);
        {
            KernelTimer timer("features", stream);
            kern<<<tiles, threads, 0, stream>>>(a);
        }
        PGEOF_LAUNCH_CHECK();
        return PGEOF_OK;
    };
    // shared-memory carve-out: the 1024 resident rows stage 48 KB of output rows; everything else should be L1 for the gathers.
    // Asking for 25 % (the 64 KB configuration) instead of the driver's choice: 2.70 -> 2.63 ms at 10 M x 50 (0 %: the CTAs no
    // longer fit two per SM, 3.17 ms; 50 %: 2.78 ms).  Also measured and not kept: one CTA of 1024 rows (3.03 ms) and per-thread
    // 44-byte stores without any staging (3.72 ms).
    const int carve = env_int("PGEOF_FEATURES_CARVEOUT", 25);
    if (cta == 128) {
        a.tma_out = a.tma_out && ((128 * 11 * 4) % 16 == 0);
        if (carve >= 0) PGEOF_CUDA(cudaFuncSetAttribute(features_direct_kernel<128>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        PGEOF_TRY(launch(features_direct_kernel<128>, 128));
    } else if (cta == 256) {
        a.tma_out = a.tma_out && ((256 * 11 * 4) % 16 == 0);
        if (carve >= 0) PGEOF_CUDA(cudaFuncSetAttribute(features_direct_kernel<256>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        PGEOF_TRY(launch(features_direct_kernel<256>, 256));
    } else if (env_int("PGEOF_FEATURES_PHASES", 2) == 2) {     // 2.63 -> 2.60 ms at 10 M x 50 (1: one phase, 64 KB carve-out)
        a.tma_out = a.tma_out && ((256 * 11 * 4) % 16 == 0);
        PGEOF_CUDA(cudaFuncSetAttribute(features_direct_kernel<512, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, env_int("PGEOF_FEATURES_CARVEOUT", 14)));
        PGEOF_TRY(launch(features_direct_kernel<512, 2>, 512));
    } else {
        a.tma_out = a.tma_out && ((512 * 11 * 4) % 16 == 0);
        if (carve >= 0) PGEOF_CUDA(cudaFuncSetAttribute(features_direct_kernel<512>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        PGEOF_TRY(launch(features_direct_kernel<512>, 512));
    }
    return device_flag_check(err.as<int>(), stream, "compute_features");
}

int features_multiscale_run(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz, RowPtr nn_ptr,
                            size_t n_rows, const uint32_t* k_scales_host, size_t n_scales, int eig_order, float* out,
                            cudaStream_t stream)
{
    if (n_rows == 0 || n_scales == 0) return PGEOF_OK;
    if (n_scales > 0xffffffffull / 11) { set_error("too many scales"); return PGEOF_EINVAL; }
    DeviceBuffer err;
    PGEOF_TRY(err.alloc(sizeof(int), stream));
    PGEOF_CUDA(cudaMemsetAsync(err.ptr, 0, sizeof(int), stream));
    FeatArgs a;
    PGEOF_TRY(make_args(&a, n_xyz, nn, nnz, nn_ptr, n_rows, eig_order, out, err.as<int>()));
    a.n_scales_total = (uint32_t)n_scales;
    Prepass pre;
    PGEOF_TRY(prepare(&a, xyz, &pre, stream));
    // the kernel writes every element itself (zeros where a row is too short: calloc semantics, pgeof.hpp:175)
    for (size_t base = 0; base < n_scales; base += kMaxScalesPerPass) {
        a.scale_base = (uint32_t)base;
        a.n_scales_pass = (uint32_t)std::min<size_t>(kMaxScalesPerPass, n_scales - base);
        for (uint32_t s = 0; s < a.n_scales_pass; ++s) a.scales[s] = k_scales_host[base + s];
        KernelTimer timer("multiscale", stream);
        // PGEOF_MULTISCALE_SPLIT = 0: the one-pass kernel (kept as an A/B switch)
        if (env_int("PGEOF_MULTISCALE_SPLIT", 1) != 0) {
            const size_t per_row = (size_t)a.n_scales_pass * 9 * sizeof(float);
            size_t chunk = std::max<size_t>(1, std::min<size_t>(n_rows, (size_t(3) << 29) / per_row));
            const int chunk_rows = env_int("PGEOF_MULTISCALE_CHUNK_ROWS", 0);   // tests: several chunks on a small input
            if (chunk_rows > 0) chunk = std::min<size_t>(chunk, (size_t)chunk_rows);
            DeviceBuffer mom;
            PGEOF_TRY(mom.alloc(chunk * per_row, stream));
            for (size_t r0 = 0; r0 < n_rows; r0 += chunk) {
                const size_t r1 = std::min(n_rows, r0 + chunk);
                multiscale_moments_kernel<<<(unsigned)((r1 - r0 + kMsThreads - 1) / kMsThreads), kMsThreads, 0, stream>>>(a, mom.as<float>(), (uint32_t)r0, (uint32_t)r1);
                PGEOF_LAUNCH_CHECK();
                multiscale_eigen_kernel<<<(unsigned)(((r1 - r0) * a.n_scales_pass + 255) / 256), 256, 0, stream>>>(a, mom.as<float>(), (uint32_t)r0, (uint32_t)r1);
                PGEOF_LAUNCH_CHECK();
            }
        } else {
        multiscale_direct_kernel<<<(unsigned)((n_rows + kMsThreads - 1) / kMsThreads), kMsThreads, 0, stream>>>(a);
        PGEOF_LAUNCH_CHECK();
        }
    }
    return device_flag_check(err.as<int>(), stream, "compute_features_multiscale");
}

int features_optimal_run(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz, RowPtr nn_ptr,
                         size_t n_rows, uint32_t k_min, uint32_t k_step, uint32_t k_min_search, int eig_order, float* out,
                         cudaStream_t stream)
{
    if (n_rows == 0) return PGEOF_OK;
    DeviceBuffer err;
    PGEOF_TRY(err.alloc(sizeof(int), stream));
    PGEOF_CUDA(cudaMemsetAsync(err.ptr, 0, sizeof(int), stream));
    FeatArgs a;
    PGEOF_TRY(make_args(&a, n_xyz, nn, nnz, nn_ptr, n_rows, eig_order, out, err.as<int>()));
    a.k_min = k_min; a.k_step = k_step; a.k_min_search = k_min_search;
    a.tma_out = a.tma_out && ((kRows * 12 * 4) % 16 == 0);
    Prepass pre;
    PGEOF_TRY(prepare(&a, xyz, &pre, stream));
    {
        KernelTimer timer("optimal", stream);
        // PGEOF_OPTIMAL_SCAN = 0: the first version (walker with the evaluation out of line), kept as an A/B switch
        const int scan = env_int("PGEOF_OPTIMAL_SCAN", 2);   // 2: closed-form float filter, 1: Jacobi float filter
        // CTAs of 128 rows per SM.  Jacobi filter: 26.8 / 23.4 / 22.5 / 23.0 ms per 10 M rows at 4 / 5 / 6 / 8 -> 6 (80 registers);
        // closed-form filter: 11.9 / 11.2 / 13.2 / 17.4 ms -> 5 (96 registers)
        if (scan == 1) optimal_scan_kernel<6, false><<<(unsigned)((n_rows + kRows - 1) / kRows), kRows, 0, stream>>>(a);
        else if (scan != 0) {
            const unsigned blocks = (unsigned)((n_rows + kRows - 1) / kRows);
            switch (env_int("PGEOF_OPTIMAL_CTAS", 5)) {
                case 4: optimal_scan_kernel<4, true><<<blocks, kRows, 0, stream>>>(a); break;
                case 6: optimal_scan_kernel<6, true><<<blocks, kRows, 0, stream>>>(a); break;
                case 8: optimal_scan_kernel<8, true><<<blocks, kRows, 0, stream>>>(a); break;
                default: optimal_scan_kernel<5, true><<<blocks, kRows, 0, stream>>>(a); break;
            }
        }
        else optimal_direct_kernel<<<(unsigned)((n_rows + kRows - 1) / kRows), kRows, 0, stream>>>(a);
    }
    PGEOF_LAUNCH_CHECK();
    return device_flag_check(err.as<int>(), stream, "compute_features_optimal");
}

}  // namespace pgeof
