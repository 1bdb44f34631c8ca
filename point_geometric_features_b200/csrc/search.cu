// search.cu -- knn_search / radius_search on the uniform grid, one warp per query.
// Replaces nanoflann_knn_search / nanoflann_radius_search (include/nn_search.hpp:31-132).
//
// Per query: (1) seed a radius from the local density of the 3x3x3 cell block,
// (2) scan the cells that intersect the ball, keeping keys under the threshold in a
// per-warp shared-memory buffer, (3) if the ball holds fewer than k points grow it, if it
// overflows the buffer shrink / bisect the threshold, (4) pick the k smallest keys with a
// register-resident counting bisection, (5) bitonic-sort them with warp shuffles and write
// the row.  Exactness: a point is only ever rejected by the 64-bit key threshold, and the
// cell coverage of a ball is computed with directed rounding (see search_core.cuh).
//
// Roofline: algorithmic bytes/query = 12 (query) + 12 (data, once) + 8k (idx + d2) = 24 + 8k.
// The kernel is issue bound (distance + compaction + sort), not HBM bound: DESIGN.md.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "search_core.cuh"

namespace pgeof {

namespace {

constexpr int kWarps = 8;

template <int NSORT, int MODE>
struct SearchCfg {
    static constexpr int M = NSORT / 32;
    static constexpr int CAP = (MODE == SEARCH_KNN) ? (NSORT <= 128 ? 256 : 2 * NSORT) : (NSORT <= 256 ? 512 : 2 * NSORT);
    static constexpr int MC = CAP / 32;
};

struct SearchArgs {
    const float4* queries;   // sorted query records (x, y, z, bits(original row))
    uint32_t n_query;
    uint32_t k;              // knn or max_knn
    float radius;            // radius modes
    float target;            // kNN: candidate count the seeded ball should hold
    void* indices;           // uint32 (knn) / int32 (radius) dense rows, or CSR nn
    float* sqr_dist;
    uint32_t* nn_ptr;        // RADIUS_COUNT: per-row counts out; RADIUS_CSR: row offsets in
};

template <int NSORT, int MODE>
__global__ void __launch_bounds__(kWarps * 32) search_kernel(const GridView g, const SearchArgs a)
{
    using Cfg = SearchCfg<NSORT, MODE>;
    constexpr int M = Cfg::M, CAP = Cfg::CAP, MC = Cfg::MC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64* keybuf = reinterpret_cast<u64*>(smem_raw) + (size_t)warp * CAP;
    const uint32_t w = blockIdx.x * kWarps + warp;
    if (w >= a.n_query) return;

    const float4 q4 = __ldg(a.queries + w);
    const float qx = q4.x, qy = q4.y, qz = q4.z;
    const uint32_t row = __float_as_uint(q4.w);
    const uint32_t k = a.k;

    uint32_t c = 0;     // keys under `tau` (<= CAP once the loops below finish)
    u64 tau = 0;
    if (MODE == SEARCH_KNN) {
        uint32_t n27, cells27;
        block27_count(g, qx, qy, qz, lane, &n27, &cells27);
        const float rho = fmaxf((float)n27, 1.f) / ((float)max(cells27, 1u) * g.h * g.h * g.h);
        float R = cbrtf(a.target / (4.18879f * rho)) + bbox_distance(g, qx, qy, qz);
        float Rg = R, Rg_hi = 0.f;
        u64 tau_lo = 0, tau_hi = 0;
        bool have_lo = false, have_hi = false;
        int shrinks = 0;
        tau = tau_from_radius(R);
        for (int it = 0; it < 512; ++it) {
            c = scan_ball<CAP, true>(g, qx, qy, qz, Rg, tau, keybuf, lane);
            if (c < k) {
                tau_lo = tau; have_lo = true;
                if (have_hi) { tau = tau_lo + (tau_hi - tau_lo) / 2; Rg = Rg_hi; }
                else {
                    const float f = fminf(fmaxf(cbrtf(1.2f * a.target / fmaxf((float)c, 0.5f)), 1.2f), 2.5f);
                    R *= f; Rg = R; tau = tau_from_radius(R);
                }
            } else if (c > (uint32_t)CAP) {
                tau_hi = tau; have_hi = true; Rg_hi = Rg;
                // dense spot: shrink the ball by the density estimate (twice at most) ...
                bool shrunk = false;
                if (!have_lo && shrinks < 2) {
                    const float Rn = R * fminf(fmaxf(cbrtf(a.target / (float)c), 0.3f), 0.9f);
                    const u64 tn = tau_from_radius(Rn);
                    if (tn < tau_hi && (tn >> 32) != 0) { R = Rn; Rg = R; tau = tn; ++shrinks; shrunk = true; }
                }
                // ... else bisect the key space between tau_lo (0: nothing is below it) and tau_hi:
                // always converges because keys are distinct (many duplicates / exact ties land here)
                if (!shrunk) { have_lo = true; tau = tau_lo + (tau_hi - tau_lo) / 2; Rg = Rg_hi; }
            } else break;
        }
    } else {
        const float r2 = __fmul_rn(a.radius, a.radius);                       // nn_search.hpp:98
        const uint32_t r2b = __float_as_uint(r2);
        if (!(r2 > 0.f)) c = 0;                                               // nothing is < 0
        else {
            tau = ((u64)r2b << 32) - 1;                                       // d2 < r2, strict
            const float Rg = __fmul_ru(__fsqrt_ru(r2), 1.0001f);
            if (MODE == SEARCH_RADIUS_COUNT) {
                c = scan_ball<CAP, false>(g, qx, qy, qz, Rg, tau, keybuf, lane);
            } else {
                c = scan_ball<CAP, true>(g, qx, qy, qz, Rg, tau, keybuf, lane);
                if (c > (uint32_t)CAP) {                                      // very dense ball: bisect by rescans
                    u64 tau_lo = 0, tau_hi = tau;
                    float cur_cnt = (float)c, cur_d2 = r2;
                    const float target = 0.5f * (float)(k + CAP);
                    for (int it = 0; it < 512; ++it) {
                        u64 mid = tau_lo + (tau_hi - tau_lo) / 2;
                        if (it < 3) {
                            const float gd2 = cur_d2 * exp2f(0.6666667f * log2f(target / cur_cnt));
                            const u64 guess = ((u64)__float_as_uint(gd2) << 32) | 0xffffffffull;
                            if (guess > tau_lo && guess < tau_hi) mid = guess;
                        }
                        tau = mid;
                        c = scan_ball<CAP, true>(g, qx, qy, qz, Rg, tau, keybuf, lane);
                        if (c < k) tau_lo = tau; else if (c > (uint32_t)CAP) tau_hi = tau; else break;
                        cur_cnt = fmaxf((float)c, 0.5f); cur_d2 = key_d2(tau);
                    }
                }
            }
        }
    }

    if (MODE == SEARCH_RADIUS_COUNT) {
        if (lane == 0) a.nn_ptr[row] = min(c, k);
        return;
    }

    // ---- select the `need` smallest keys and sort them -------------------------------
    const uint32_t need = (MODE == SEARCH_KNN) ? k : min(c, k);
    u64 v[M];
    if (c > (uint32_t)NSORT) {
        u64 key[MC];
#pragma unroll
        for (int r = 0; r < MC; ++r) {
            const uint32_t e = r * 32 + lane;
            key[r] = ((uint32_t)(r * 32) < c && e < c) ? keybuf[e] : kKeyMax;
        }
        const u64 t = select_threshold<MC>(key, c, tau, need, NSORT);
        __syncwarp();
        const unsigned lt = lanemask_lt();
        uint32_t off = 0;
#pragma unroll
        for (int r = 0; r < MC; ++r) {
            if ((uint32_t)(r * 32) < c) {
                const bool acc = key[r] <= t;
                const unsigned m = __ballot_sync(kFull, acc);
                if (acc) keybuf[off + __popc(m & lt)] = key[r];
                off += __popc(m);
            }
        }
        __syncwarp();
        c = off;
    }
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const uint32_t e = m * 32 + lane;
        v[m] = e < c ? keybuf[e] : kKeyMax;
    }
    warp_bitonic_sort<M>(v, lane);

    // ---- write the row ----------------------------------------------------------------
    if (MODE == SEARCH_KNN) {
        uint32_t* idx = reinterpret_cast<uint32_t*>(a.indices) + (size_t)row * k;
        float* d2 = a.sqr_dist + (size_t)row * k;
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const uint32_t e = m * 32 + lane;
            if (e < k) { idx[e] = key_idx(v[m]); d2[e] = key_d2(v[m]); }
        }
    } else if (MODE == SEARCH_RADIUS) {
        int32_t* idx = reinterpret_cast<int32_t*>(a.indices) + (size_t)row * k;
        float* d2 = a.sqr_dist + (size_t)row * k;
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const uint32_t e = m * 32 + lane;
            if (e < k) {                                   // pad: -1 / 0 (nn_search.hpp:104,108)
                const bool hit = e < need;
                idx[e] = hit ? (int32_t)key_idx(v[m]) : -1;
                d2[e] = hit ? key_d2(v[m]) : 0.f;
            }
        }
    } else {   // SEARCH_RADIUS_CSR
        const uint32_t base = a.nn_ptr[row];
        uint32_t* nn = reinterpret_cast<uint32_t*>(a.indices);
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const uint32_t e = m * 32 + lane;
            if (e < need) {
                nn[(size_t)base + e] = key_idx(v[m]);
                if (a.sqr_dist) a.sqr_dist[(size_t)base + e] = key_d2(v[m]);
            }
        }
    }
}

template <int NSORT, int MODE>
int launch_search(const GridView& g, const SearchArgs& a, cudaStream_t stream)
{
    using Cfg = SearchCfg<NSORT, MODE>;
    const size_t smem = (size_t)kWarps * Cfg::CAP * sizeof(u64);
    auto kern = search_kernel<NSORT, MODE>;
    if (smem > 48 * 1024) PGEOF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned blocks = (a.n_query + kWarps - 1) / kWarps;
    {
        KernelTimer timer(MODE == SEARCH_KNN ? "knn_search" : "radius_search", stream);
        kern<<<blocks, kWarps * 32, smem, stream>>>(g, a);
    }
    PGEOF_LAUNCH_CHECK();
    return PGEOF_OK;
}

template <int MODE>
int dispatch_search(uint32_t k, const GridView& g, const SearchArgs& a, cudaStream_t stream)
{
    if (k <= 32) return launch_search<32, MODE>(g, a, stream);
    if (k <= 64) return launch_search<64, MODE>(g, a, stream);
    if (k <= 128) return launch_search<128, MODE>(g, a, stream);
    if (k <= 256) return launch_search<256, MODE>(g, a, stream);
    if (k <= 512) return launch_search<512, MODE>(g, a, stream);
    set_error("knn / max_knn = %u exceeds the supported maximum of 512 neighbours per query", k);
    return PGEOF_EINVAL;
}

float env_float(const char* name, float dflt)
{
    const char* e = std::getenv(name);
    return e ? (float)std::atof(e) : dflt;
}

}  // namespace

int search_run(SearchMode mode, const float* data, size_t n_data, const float* query, size_t n_query, uint32_t k,
               float radius, void* indices, float* sqr_dist, uint32_t* nn_ptr, cudaStream_t stream)
{
    if (n_query == 0 || n_data == 0) return PGEOF_OK;
    if (n_query > 0xfffffff0ull) { set_error("n_query too large"); return PGEOF_EINVAL; }
    Grid grid;
    float target = 0.f;
    if (mode == SEARCH_KNN) {
        // ball seeded to hold k + 2 sigma + 2 points; cell edge ~ that ball's radius
        target = (float)k + 2.f * std::sqrt((float)k) + 2.f;
        const float occ = std::max(2.f, target * env_float("PGEOF_KNN_CELL_OCC", 0.25f));
        PGEOF_TRY(grid_build(data, n_data, 0.f, occ, stream, &grid));
    } else {
        if (!(radius >= 0.f) || !std::isfinite(radius)) { set_error("search_radius must be finite and >= 0"); return PGEOF_EINVAL; }
        const float edge = radius * env_float("PGEOF_RADIUS_CELL_SCALE", 1.0f);
        PGEOF_TRY(grid_build(data, n_data, edge > 0.f ? edge : 1.f, 0.f, stream, &grid));
    }
    DeviceBuffer qsorted;
    const float4* qrec;
    if (query == data && n_query == n_data) qrec = grid.view.pts;   // self query: reuse the sorted cloud
    else { PGEOF_TRY(grid_sort_queries(grid, query, n_query, stream, &qsorted)); qrec = qsorted.as<float4>(); }
    SearchArgs a{qrec, (uint32_t)n_query, k, radius, target, indices, sqr_dist, nn_ptr};
    switch (mode) {
        case SEARCH_KNN: return dispatch_search<SEARCH_KNN>(k, grid.view, a, stream);
        case SEARCH_RADIUS: return dispatch_search<SEARCH_RADIUS>(k, grid.view, a, stream);
        case SEARCH_RADIUS_COUNT: return launch_search<32, SEARCH_RADIUS_COUNT>(grid.view, a, stream);
        case SEARCH_RADIUS_CSR: return dispatch_search<SEARCH_RADIUS_CSR>(k, grid.view, a, stream);
    }
    return PGEOF_EINVAL;
}

}  // namespace pgeof
