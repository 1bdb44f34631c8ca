// grid_index.cu -- uniform-grid index build: bbox reduce, cell histogram (warp-aggregated
// atomics that also hand every point its rank inside the cell), exclusive scan, scatter.
// Replaces the per-call nanoflann KD-tree build (nn_search.hpp:39,97; pgeof.hpp:333).
//
// HBM traffic per point (n points, c cells): read 12 B (bbox) + read 12 B, write 4 B rank
// (histogram) + read 12 B + 4 B, write 16 B (scatter) = 60 B/pt, plus 12 B per cell for the
// scan.  All passes are streaming; the scatter writes 16-B records to random cells.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>

#include "grid.cuh"

namespace pgeof {

namespace {

constexpr int kBBoxBlocks = 592;   // partial bounding boxes per cloud (a fixed count, sized as 4 CTAs per SM of a B200)
constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) bbox_kernel(const float* __restrict__ xyz, size_t n, float* __restrict__ partial)
{
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float v = __ldg(xyz + 3 * i + d);
            mn[d] = fminf(mn[d], v);
            mx[d] = fmaxf(mx[d], v);
        }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
            mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
        }
    __shared__ float s[kThreads / 32][6];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { for (int d = 0; d < 3; ++d) { s[w][d] = mn[d]; s[w][3 + d] = mx[d]; } }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = s[0][threadIdx.x];
        for (int i = 1; i < kThreads / 32; ++i) v = threadIdx.x < 3 ? fminf(v, s[i][threadIdx.x]) : fmaxf(v, s[i][threadIdx.x]);
        partial[blockIdx.x * 6 + threadIdx.x] = v;
    }
}

// Histogram with warp-aggregated atomics.  rank[i] = position of point i inside its cell.
__device__ __forceinline__ bool inside_clip(const GridView& g, float x, float y, float z)
{
    return x >= g.clip_lo[0] && x <= g.clip_hi[0] && y >= g.clip_lo[1] && y <= g.clip_hi[1] && z >= g.clip_lo[2] && z <= g.clip_hi[2];
}

// use_clip: points outside the grid's clip box are not indexed (data); queries are always kept
__global__ void __launch_bounds__(kThreads) cell_count_kernel(GridView g, const float* __restrict__ xyz, size_t n,
                                                              uint32_t* __restrict__ counts, uint32_t* __restrict__ rank, int use_clip)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = i < n;
    uint32_t cid = 0xffffffffu;
    if (valid) {
        const float x = __ldg(xyz + 3 * i), y = __ldg(xyz + 3 * i + 1), z = __ldg(xyz + 3 * i + 2);
        valid = !use_clip || inside_clip(g, x, y, z);
        if (valid) cid = cell_index(g, x, y, z);
    }
    const unsigned active = __ballot_sync(0xffffffffu, valid);
    if (!valid) return;
    const unsigned peers = __match_any_sync(active, cid);
    const int leader = __ffs(peers) - 1;
    const int lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counts + cid, (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    rank[i] = base + (uint32_t)__popc(peers & lanemask_lt());
}

// sum of c^2 over the cell counts: sum c^2 / n is the occupancy of the cell an average POINT lives in.  Uniform data:
// mean occupancy + 1; a cloud of surfaces / clusters in a mostly empty bounding box: orders of magnitude more.
__global__ void __launch_bounds__(kThreads) occupancy_kernel(const uint32_t* __restrict__ counts, size_t n_cells, unsigned long long* __restrict__ sums)
{
    unsigned long long acc = 0, cnt = 0;    // sums[0] = sum c^2, sums[1] = sum c (points indexed: < n on a clipped grid)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_cells; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long c = __ldg(counts + i);
        acc += c * c;
        cnt += c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { acc += __shfl_xor_sync(0xffffffffu, acc, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o); }
    if ((threadIdx.x & 31) == 0 && cnt) { atomicAdd(sums, acc); atomicAdd(sums + 1, cnt); }
}

__global__ void __launch_bounds__(kThreads) scatter_kernel(GridView g, const float* __restrict__ xyz, size_t n,
                                                           const uint32_t* __restrict__ cell_start,
                                                           const uint32_t* __restrict__ rank, float4* __restrict__ out, int use_clip)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = __ldg(xyz + 3 * i), y = __ldg(xyz + 3 * i + 1), z = __ldg(xyz + 3 * i + 2);
    if (use_clip && !inside_clip(g, x, y, z)) return;
    const uint32_t cid = cell_index(g, x, y, z);
    const uint32_t pos = __ldg(cell_start + cid) + __ldg(rank + i);
    out[pos] = make_float4(x, y, z, __uint_as_float((uint32_t)i));
}

// The scatter in two steps: 16-byte records written to random places of a 160 MB array (10 M points) leave L2 as half-written
// sectors; a 4-byte permutation (40 MB, stays in L2) written at random and a second pass that READS the cloud at random
// (120 MB, in L2 from the count pass) and WRITES the records in order moves the same data with coalesced stores.
__global__ void __launch_bounds__(kThreads) perm_kernel(GridView g, const float* __restrict__ xyz, size_t n, const uint32_t* __restrict__ cell_start,
                                                        const uint32_t* __restrict__ rank, uint32_t* __restrict__ perm, int use_clip)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = __ldg(xyz + 3 * i), y = __ldg(xyz + 3 * i + 1), z = __ldg(xyz + 3 * i + 2);
    if (use_clip && !inside_clip(g, x, y, z)) return;
    perm[__ldg(cell_start + cell_index(g, x, y, z)) + __ldg(rank + i)] = (uint32_t)i;
}

__global__ void __launch_bounds__(kThreads) gather_sorted_kernel(const float* __restrict__ xyz, size_t n, const uint32_t* __restrict__ perm,
                                                                 const uint32_t* __restrict__ total, float4* __restrict__ out)
{
    const size_t pos = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= n || pos >= __ldg(total)) return;               // a clipped grid indexes fewer than n points
    const uint32_t i = __ldg(perm + pos);
    out[pos] = make_float4(__ldg(xyz + 3 * (size_t)i), __ldg(xyz + 3 * (size_t)i + 1), __ldg(xyz + 3 * (size_t)i + 2), __uint_as_float(i));
}

// ---------------------------------------------------------------------------
// exclusive scan, three passes: tile sums -> scan of tile sums -> rescan + offset
// ---------------------------------------------------------------------------
constexpr int kScanItems = 16;
constexpr int kScanTile = kThreads * kScanItems;   // 4096 counts per block

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total)
{
    __shared__ uint32_t warp_sums[kThreads / 32];
    __shared__ uint32_t block_total;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t ws = lane < kThreads / 32 ? warp_sums[lane] : 0;
        uint32_t winc = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < kThreads / 32) warp_sums[lane] = winc - ws;
        if (lane == kThreads / 32 - 1) block_total = winc;
    }
    __syncthreads();
    const uint32_t r = warp_sums[w] + inc - v;
    *total = block_total;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kThreads) scan_tile_sums(const uint32_t* __restrict__ data, size_t n, uint32_t* __restrict__ tile_sums)
{
    const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) if (base + j < n) s += data[base + j];
    uint32_t total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kThreads) scan_tile_offsets(uint32_t* __restrict__ tile_sums, size_t n_tiles, uint32_t* __restrict__ grand_total)
{
    uint32_t carry = 0;
    for (size_t base = 0; base < n_tiles; base += kThreads) {
        const size_t i = base + threadIdx.x;
        const uint32_t v = i < n_tiles ? tile_sums[i] : 0;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(v, &total);
        if (i < n_tiles) tile_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) *grand_total = carry;
}

__global__ void __launch_bounds__(kThreads) scan_apply(uint32_t* __restrict__ data, size_t n, const uint32_t* __restrict__ tile_offsets)
{
    const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) { v[j] = base + j < n ? data[base + j] : 0; s += v[j]; }
    uint32_t total;
    uint32_t run = block_exclusive_scan(s, &total) + tile_offsets[blockIdx.x];
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
        if (base + j < n) data[base + j] = run;
        run += v[j];
    }
}

}  // namespace

int bbox_partials(const float* xyz, size_t n, DeviceBuffer* out, int* n_partials, cudaStream_t stream)
{
    PGEOF_TRY(out->alloc(kBBoxBlocks * 6 * sizeof(float), stream));
    bbox_kernel<<<kBBoxBlocks, kThreads, 0, stream>>>(xyz, n, out->as<float>());
    PGEOF_LAUNCH_CHECK();
    *n_partials = kBBoxBlocks;
    return PGEOF_OK;
}

int exclusive_scan_u32(uint32_t* data, size_t n, cudaStream_t stream)
{
    if (n == 0) { PGEOF_CUDA(cudaMemsetAsync(data, 0, sizeof(uint32_t), stream)); return PGEOF_OK; }
    const size_t n_tiles = (n + kScanTile - 1) / kScanTile;
    DeviceBuffer sums;
    PGEOF_TRY(sums.alloc(n_tiles * sizeof(uint32_t), stream));
    scan_tile_sums<<<(unsigned)n_tiles, kThreads, 0, stream>>>(data, n, sums.as<uint32_t>());
    PGEOF_LAUNCH_CHECK();
    scan_tile_offsets<<<1, kThreads, 0, stream>>>(sums.as<uint32_t>(), n_tiles, data + n);
    PGEOF_LAUNCH_CHECK();
    scan_apply<<<(unsigned)n_tiles, kThreads, 0, stream>>>(data, n, sums.as<uint32_t>());
    PGEOF_LAUNCH_CHECK();
    return PGEOF_OK;
}

static int reduce_bbox(const float* hp, float (&lo)[3], float (&hi)[3])
{
    for (int d = 0; d < 3; ++d) { lo[d] = FLT_MAX; hi[d] = -FLT_MAX; }
    for (int b = 0; b < kBBoxBlocks; ++b)
        for (int d = 0; d < 3; ++d) { lo[d] = std::min(lo[d], hp[b * 6 + d]); hi[d] = std::max(hi[d], hp[b * 6 + 3 + d]); }
    for (int d = 0; d < 3; ++d)
        if (!(lo[d] <= hi[d]) || !std::isfinite(lo[d]) || !std::isfinite(hi[d])) { set_error("point cloud holds non-finite coordinates"); return PGEOF_EINVAL; }
    return PGEOF_OK;
}

int bbox_host(const float* a, size_t na, const float* b, size_t nb, cudaStream_t stream, float (&alo)[3], float (&ahi)[3],
              float (&blo)[3], float (&bhi)[3])
{
    DeviceBuffer partial;
    PGEOF_TRY(partial.alloc(2 * kBBoxBlocks * 6 * sizeof(float), stream));
    bbox_kernel<<<kBBoxBlocks, kThreads, 0, stream>>>(a, na, partial.as<float>());
    PGEOF_LAUNCH_CHECK();
    if (b) {
        bbox_kernel<<<kBBoxBlocks, kThreads, 0, stream>>>(b, nb, partial.as<float>() + kBBoxBlocks * 6);
        PGEOF_LAUNCH_CHECK();
    }
    static thread_local float hp[2 * kBBoxBlocks * 6];
    PGEOF_CUDA(cudaMemcpyAsync(hp, partial.ptr, (b ? 2 : 1) * kBBoxBlocks * 6 * sizeof(float), cudaMemcpyDeviceToHost, stream));
    PGEOF_CUDA(cudaStreamSynchronize(stream));
    PGEOF_TRY(reduce_bbox(hp, alo, ahi));
    if (b) PGEOF_TRY(reduce_bbox(hp + kBBoxBlocks * 6, blo, bhi));
    return PGEOF_OK;
}

static size_t max_cells_for(size_t n)
{
    // the dense cell table costs 12 B of scan traffic per cell; keep it comparable to the point data
    // (C3, 10 M LiDAR-like points, r = 0.2: 2 n / 4 n / 8 n / 16 n cells -> 12.73 / 12.45 / 12.70 / 13.04 ms per step)
    size_t cap = std::max<size_t>(4 * n, (size_t)1 << 20);
    if (const char* e = std::getenv("PGEOF_MAX_CELLS")) cap = (size_t)std::strtoull(e, nullptr, 10);
    return std::min<size_t>(cap, (size_t)1 << 28);
}

int grid_build(const float* xyz, size_t n, float cell_edge, float target_occupancy, int x_refine, cudaStream_t stream, Grid* out,
               const GridClip* clip)
{
    if (n == 0 || n > 0xfffffff0ull) { set_error("grid_build: n=%zu out of range", n); return PGEOF_EINVAL; }
    KernelTimer timer("grid_build", stream);
    // 1. bounding box (of the clip box when the caller restricts the index; it already knows the cloud's)
    float lo[3], hi[3];
    if (clip) {
        for (int d = 0; d < 3; ++d) { lo[d] = clip->lo[d]; hi[d] = clip->hi[d]; }
    } else {
        float qlo[3], qhi[3];
        PGEOF_TRY(bbox_host(xyz, n, nullptr, 0, stream, lo, hi, qlo, qhi));
    }
    double ext[3], maxabs = 0;
    for (int d = 0; d < 3; ++d) {
        ext[d] = (double)hi[d] - (double)lo[d];
        maxabs = std::max(maxabs, std::max(std::fabs((double)lo[d]), std::fabs((double)hi[d])));
    }
    // 2. cell edge
    double h = cell_edge;
    const bool adaptive = !(h > 0);
    if (adaptive) {
        if (clip) h = clip->cell_edge;     // density of the whole cloud (the number of points inside the box is not known yet)
        else {
            double vol = 1; int dims = 0;
            for (int d = 0; d < 3; ++d) if (ext[d] > 0) { vol *= ext[d]; ++dims; }
            h = dims ? std::pow(vol * std::max(1.0f, target_occupancy) / (double)n, 1.0 / dims) : 1.0;
        }
    }
    const size_t cap = max_cells_for(n);
    GridView& g = out->view;
    DeviceBuffer rank, stat;
    PGEOF_TRY(rank.alloc(n * sizeof(uint32_t), stream));
    const unsigned blocks = (unsigned)((n + kThreads - 1) / kThreads);
    uint32_t* cs = nullptr;
    for (int attempt = 0;; ++attempt) {
        // float32 cannot resolve cells much finer than an ulp of the coordinates
        h = std::max(h, std::max(maxabs * 1e-5, 1e-30));
        int xf = std::max(1, std::min(x_refine, 16));
        int nc[3];
        for (;;) {
            double cells = 1;
            for (int d = 0; d < 3; ++d) {
                const double edge = d == 0 ? h / xf : h;
                const double c = std::floor(ext[d] / edge) + 1;
                nc[d] = (int)std::min(c, 2e9);
                cells *= c;
            }
            if (cells <= (double)cap) break;
            if (xf > 1) xf >>= 1; else h *= 1.26;
        }
        for (int d = 0; d < 3; ++d) { g.lo[d] = lo[d]; g.n[d] = nc[d]; }
        g.h = (float)h;
        g.inv_h = 1.0f / g.h;
        g.hx = (float)(h / xf);
        g.inv_hx = 1.0f / g.hx;
        g.xf = xf;
        g.slack = (float)((2.0 * maxabs + h) * 9.5367431640625e-7);   // 2^-20
        g.n_pts = (uint32_t)n;
        for (int d = 0; d < 3; ++d) { g.clip_lo[d] = clip ? clip->lo[d] : -FLT_MAX; g.clip_hi[d] = clip ? clip->hi[d] : FLT_MAX; }
        g.rmax_safe = clip ? clip->rmax_safe : INFINITY;
        out->n_cells = (size_t)nc[0] * nc[1] * nc[2];
        // 3. histogram
        PGEOF_TRY(out->cell_start.alloc((out->n_cells + 1) * sizeof(uint32_t), stream));
        cs = out->cell_start.as<uint32_t>();
        PGEOF_CUDA(cudaMemsetAsync(cs, 0, (out->n_cells + 1) * sizeof(uint32_t), stream));
        cell_count_kernel<<<blocks, kThreads, 0, stream>>>(g, xyz, n, cs, rank.as<uint32_t>(), clip ? 1 : 0);
        PGEOF_LAUNCH_CHECK();
        if (!adaptive || attempt >= 2) break;
        // 3b. the edge above assumes the points fill their bounding box.  If the cell an average point lives in is far
        // fuller than intended (surfaces / clusters in a mostly empty box), shrink the cells and count again.
        if (!stat.ptr) PGEOF_TRY(stat.alloc(2 * sizeof(unsigned long long), stream));
        PGEOF_CUDA(cudaMemsetAsync(stat.ptr, 0, 2 * sizeof(unsigned long long), stream));
        occupancy_kernel<<<kBBoxBlocks, kThreads, 0, stream>>>(cs, out->n_cells, stat.as<unsigned long long>());
        PGEOF_LAUNCH_CHECK();
        unsigned long long sums[2] = {0, 0};
        PGEOF_CUDA(cudaMemcpyAsync(sums, stat.ptr, sizeof(sums), cudaMemcpyDeviceToHost, stream));
        PGEOF_CUDA(cudaStreamSynchronize(stream));
        const double seen = (double)sums[0] / (double)std::max<unsigned long long>(sums[1], 1);                                  // occupancy seen by a point (x-refined cells)
        const double want = std::max(1.0f, target_occupancy) / xf + 1.0;                  // uniform data: mean + 1
        if (seen <= 2.5 * want || out->n_cells * 2 > cap) break;
        h *= std::max(0.25, std::pow(want / seen, 1.0 / 2.5));                            // between a surface (1/2) and a volume (1/3)
    }
    // 4. scan -> scatter
    PGEOF_TRY(out->pts.alloc(n * sizeof(float4), stream));
    PGEOF_TRY(exclusive_scan_u32(cs, out->n_cells, stream));
    // (the second pass reads the cloud at random: it only pays while the cloud fits in L2 -- 10 M points: 0.71 -> 0.63 ms,
    // 50 M points: 4.0 -> 4.7 ms -- so larger clouds keep the one-step scatter; PGEOF_GRID_SCATTER = 1 / 2 forces either)
    const char* sc = std::getenv("PGEOF_GRID_SCATTER");
    const int sc_mode = sc ? std::atoi(sc) : 0;
    // (a clipped grid indexes a fraction of the cloud: two passes over all n points cost it more than they save, 0.17 -> 0.20 ms
    // per rank of an 8-way run)
    if (sc_mode == 2 || (sc_mode != 1 && !clip && n * 12 <= ((size_t)160 << 20))) {
        DeviceBuffer perm;
        PGEOF_TRY(perm.alloc(n * sizeof(uint32_t), stream));
        perm_kernel<<<blocks, kThreads, 0, stream>>>(g, xyz, n, cs, rank.as<uint32_t>(), perm.as<uint32_t>(), clip ? 1 : 0);
        PGEOF_LAUNCH_CHECK();
        gather_sorted_kernel<<<blocks, kThreads, 0, stream>>>(xyz, n, perm.as<uint32_t>(), cs + out->n_cells, out->pts.as<float4>());
        PGEOF_LAUNCH_CHECK();
    } else {
    scatter_kernel<<<blocks, kThreads, 0, stream>>>(g, xyz, n, cs, rank.as<uint32_t>(), out->pts.as<float4>(), clip ? 1 : 0);
    PGEOF_LAUNCH_CHECK();
    }
    g.cell_start = cs;
    g.pts = out->pts.as<float4>();
    return PGEOF_OK;
}

int grid_sort_queries(const Grid& grid, const float* query, size_t nq, cudaStream_t stream, DeviceBuffer* out)
{
    PGEOF_TRY(out->alloc(std::max<size_t>(nq, 1) * sizeof(float4), stream));
    if (nq == 0) return PGEOF_OK;
    DeviceBuffer counts, rank;
    PGEOF_TRY(counts.alloc((grid.n_cells + 1) * sizeof(uint32_t), stream));
    PGEOF_TRY(rank.alloc(nq * sizeof(uint32_t), stream));
    PGEOF_CUDA(cudaMemsetAsync(counts.ptr, 0, (grid.n_cells + 1) * sizeof(uint32_t), stream));
    const unsigned blocks = (unsigned)((nq + kThreads - 1) / kThreads);
    cell_count_kernel<<<blocks, kThreads, 0, stream>>>(grid.view, query, nq, counts.as<uint32_t>(), rank.as<uint32_t>(), 0);
    PGEOF_LAUNCH_CHECK();
    PGEOF_TRY(exclusive_scan_u32(counts.as<uint32_t>(), grid.n_cells, stream));
    scatter_kernel<<<blocks, kThreads, 0, stream>>>(grid.view, query, nq, counts.as<uint32_t>(), rank.as<uint32_t>(), out->as<float4>(), 0);
    PGEOF_LAUNCH_CHECK();
    return PGEOF_OK;
}

}  // namespace pgeof
