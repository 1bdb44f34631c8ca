// shard.cu -- spatial query shards for multi-GPU callers.
//
// The path shards with no exchange step (nn_search.hpp:51-58, pgeof.hpp:95-108): the cloud is replicated and every
// rank owns a SLAB of the queries along one axis, which keeps the query density the kNN tile kernel needs (a random
// 1/N of the rows costs almost as much as all of them).  The slab edges come from a 4096-bin histogram of the
// replicated cloud, so every rank derives the same edges without a collective; the rows of a slab are compacted in
// input order.  Two calls, no hidden state: pgeof_slab_plan_dev (bounding interval, histogram, edges, row count ->
// one host synchronisation) and pgeof_slab_fill_dev (row ids + query coordinates).
#include <cfloat>

#include "common.cuh"

namespace pgeof {

namespace {

constexpr int kBins = 4096;
constexpr int kThreads = 256;
constexpr int kItems = 8;
constexpr int kTile = kThreads * kItems;

struct SlabPlanDev { float lo, scale; uint32_t b_lo, b_hi; unsigned long long count; };

__device__ __forceinline__ int slab_bin(float v, float lo, float scale)
{
    const int b = __float2int_rd(__fmul_rn(__fsub_rn(v, lo), scale));   // same rounding as the torch restatement in shard.py
    return min(max(b, 0), kBins - 1);
}

// one warp: interval of the axis from the per-block bounding boxes
__global__ void slab_interval_kernel(const float* __restrict__ partial, int n_partial, int axis, SlabPlanDev* __restrict__ plan)
{
    const int lane = threadIdx.x;
    float mn = FLT_MAX, mx = -FLT_MAX;
    for (int b = lane; b < n_partial; b += 32) { mn = fminf(mn, partial[b * 6 + axis]); mx = fmaxf(mx, partial[b * 6 + 3 + axis]); }
    for (int o = 16; o > 0; o >>= 1) { mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
    if (lane == 0) {
        const float ext = __fsub_rn(mx, mn);
        plan->lo = mn;
        plan->scale = (ext > 0.f && ext < 3.0e38f) ? __fdiv_rn((float)kBins, ext) : 0.f;
    }
}

__global__ void __launch_bounds__(kThreads) slab_hist_kernel(const float* __restrict__ xyz, size_t n, int axis, const SlabPlanDev* __restrict__ plan,
                                                             uint32_t* __restrict__ hist)
{
    __shared__ uint32_t s_hist[kBins];
    for (int i = threadIdx.x; i < kBins; i += kThreads) s_hist[i] = 0;
    __syncthreads();
    const float lo = plan->lo, scale = plan->scale;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads)
        atomicAdd(s_hist + slab_bin(__ldg(xyz + 3 * i + axis), lo, scale), 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < kBins; i += kThreads) if (s_hist[i]) atomicAdd(hist + i, s_hist[i]);
}

// one block: prefix over the bins -> slab r = bins [b_r, b_{r+1}), b_r = first bin whose exclusive prefix reaches r n / world
__global__ void __launch_bounds__(1024) slab_edges_kernel(const uint32_t* __restrict__ hist, unsigned long long n, int rank, int world,
                                                          SlabPlanDev* __restrict__ plan)
{
    __shared__ unsigned long long s_cum[kBins + 1];
    __shared__ unsigned long long s_part[1024];
    const int t = threadIdx.x;
    unsigned long long local[4], sum = 0;
    for (int j = 0; j < 4; ++j) { local[j] = hist[t * 4 + j]; sum += local[j]; }
    s_part[t] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const unsigned long long v = t >= o ? s_part[t - o] : 0ull;
        __syncthreads();
        s_part[t] += v;
        __syncthreads();
    }
    unsigned long long run = s_part[t] - sum;
    for (int j = 0; j < 4; ++j) { s_cum[t * 4 + j] = run; run += local[j]; }
    if (t == 1023) s_cum[kBins] = run;
    __syncthreads();
    if (t < 2) {
        const unsigned long long target = (unsigned long long)(rank + t) * n / (unsigned long long)world;
        int lo = 0, hi = kBins;                     // smallest b with cum[b] >= target
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_cum[mid] >= target) hi = mid; else lo = mid + 1; }
        if (t == 0) plan->b_lo = (uint32_t)lo; else plan->b_hi = (uint32_t)lo;
    }
    __syncthreads();
    if (t == 0) plan->count = s_cum[plan->b_hi] - s_cum[plan->b_lo];
}

__device__ __forceinline__ bool in_slab(float v, float lo, float scale, uint32_t b_lo, uint32_t b_hi)
{
    const uint32_t b = (uint32_t)slab_bin(v, lo, scale);
    return b >= b_lo && b < b_hi;
}

__global__ void __launch_bounds__(kThreads) slab_count_kernel(const float* __restrict__ xyz, size_t n, int axis, float lo, float scale,
                                                              uint32_t b_lo, uint32_t b_hi, uint32_t* __restrict__ tile_counts)
{
    const size_t base = (size_t)blockIdx.x * kTile;
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < kItems; ++j) {
        const size_t i = base + (size_t)j * kThreads + threadIdx.x;
        if (i < n && in_slab(__ldg(xyz + 3 * i + axis), lo, scale, b_lo, b_hi)) ++c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    __shared__ uint32_t s[kThreads / 32];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t total = 0;
        for (int w = 0; w < kThreads / 32; ++w) total += s[w];
        tile_counts[blockIdx.x] = total;
    }
}

// rows in input order: item (j, thread) of a tile is point base + j * kThreads + thread, so the tile is walked j-major
__global__ void __launch_bounds__(kThreads) slab_write_kernel(const float* __restrict__ xyz, size_t n, int axis, float lo, float scale,
                                                              uint32_t b_lo, uint32_t b_hi, const uint32_t* __restrict__ tile_offsets,
                                                              long long* __restrict__ rows, float* __restrict__ query)
{
    __shared__ uint32_t s_warp[kThreads / 32];
    __shared__ uint32_t s_run;
    const size_t base = (size_t)blockIdx.x * kTile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_run = tile_offsets[blockIdx.x];
    __syncthreads();
    for (int j = 0; j < kItems; ++j) {
        const size_t i = base + (size_t)j * kThreads + threadIdx.x;
        float x = 0.f, y = 0.f, z = 0.f;
        bool in = false;
        if (i < n) {
            x = __ldg(xyz + 3 * i); y = __ldg(xyz + 3 * i + 1); z = __ldg(xyz + 3 * i + 2);
            in = in_slab(axis == 0 ? x : (axis == 1 ? y : z), lo, scale, b_lo, b_hi);
        }
        const unsigned m = __ballot_sync(0xffffffffu, in);
        if (lane == 0) s_warp[warp] = (uint32_t)__popc(m);
        __syncthreads();
        uint32_t off = s_run;
        for (int w = 0; w < warp; ++w) off += s_warp[w];
        if (in) {
            const size_t pos = (size_t)off + (size_t)__popc(m & ((1u << lane) - 1u));
            rows[pos] = (long long)i;
            query[3 * pos] = x; query[3 * pos + 1] = y; query[3 * pos + 2] = z;
        }
        __syncthreads();
        if (threadIdx.x == 0) { uint32_t t = 0; for (int w = 0; w < kThreads / 32; ++w) t += s_warp[w]; s_run += t; }
        __syncthreads();
    }
}

}  // namespace

int slab_plan_run(const float* xyz, size_t n, int rank, int world, int axis, pgeof_slab_plan* out, cudaStream_t stream)
{
    DeviceBuffer partial, hist, plan;
    int n_partial = 0;
    PGEOF_TRY(bbox_partials(xyz, n, &partial, &n_partial, stream));
    PGEOF_TRY(hist.alloc(kBins * sizeof(uint32_t), stream));
    PGEOF_TRY(plan.alloc(sizeof(SlabPlanDev), stream));
    PGEOF_CUDA(cudaMemsetAsync(hist.ptr, 0, kBins * sizeof(uint32_t), stream));
    slab_interval_kernel<<<1, 32, 0, stream>>>(partial.as<float>(), n_partial, axis, plan.as<SlabPlanDev>());
    PGEOF_LAUNCH_CHECK();
    slab_hist_kernel<<<sm_count() * 4, kThreads, 0, stream>>>(xyz, n, axis, plan.as<SlabPlanDev>(), hist.as<uint32_t>());
    PGEOF_LAUNCH_CHECK();
    slab_edges_kernel<<<1, 1024, 0, stream>>>(hist.as<uint32_t>(), (unsigned long long)n, rank, world, plan.as<SlabPlanDev>());
    PGEOF_LAUNCH_CHECK();
    SlabPlanDev h;
    PGEOF_CUDA(cudaMemcpyAsync(&h, plan.ptr, sizeof(h), cudaMemcpyDeviceToHost, stream));
    PGEOF_CUDA(cudaStreamSynchronize(stream));
    out->lo = h.lo; out->scale = h.scale; out->bin_lo = h.b_lo; out->bin_hi = h.b_hi; out->count = h.count; out->axis = axis;
    return PGEOF_OK;
}

int slab_fill_run(const float* xyz, size_t n, const pgeof_slab_plan* plan, long long* rows, float* query, cudaStream_t stream)
{
    if (n == 0 || plan->count == 0) return PGEOF_OK;
    const size_t tiles = (n + kTile - 1) / kTile;
    DeviceBuffer counts;
    PGEOF_TRY(counts.alloc((tiles + 1) * sizeof(uint32_t), stream));
    slab_count_kernel<<<(unsigned)tiles, kThreads, 0, stream>>>(xyz, n, plan->axis, plan->lo, plan->scale, plan->bin_lo, plan->bin_hi, counts.as<uint32_t>());
    PGEOF_LAUNCH_CHECK();
    PGEOF_TRY(exclusive_scan_u32(counts.as<uint32_t>(), tiles, stream));
    slab_write_kernel<<<(unsigned)tiles, kThreads, 0, stream>>>(xyz, n, plan->axis, plan->lo, plan->scale, plan->bin_lo, plan->bin_hi,
                                                                  counts.as<uint32_t>(), rows, query);
    PGEOF_LAUNCH_CHECK();
    return PGEOF_OK;
}

}  // namespace pgeof
