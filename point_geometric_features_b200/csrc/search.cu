// search.cu -- knn_search / radius_search on the uniform grid.
// Replaces nanoflann_knn_search / nanoflann_radius_search (include/nn_search.hpp:31-132).
//
// The kernels share the grid and the exactness argument (a point is only ever rejected by the
// 64-bit (d2, index) key threshold, and the cell coverage of a ball is computed with directed
// rounding, see search_core.cuh):
//
//  * knn_tile_kernel (kNN and padded radius search, k / max_knn <= 64): ONE THREAD PER QUERY, one warp per
//    32 consecutive queries of the cell-sorted order.  The 32 queries of a warp live in one (y, z) cell row, so they share one
//    candidate region: the cells that intersect the warp's query bounding box dilated by the
//    search radius R (kNN: seeded from the density and shape of the 3 x 3 block of cell rows; radius
//    search: given), staged in shared memory by one 1-D TMA bulk copy per cell row.  Every candidate is read ONCE per warp (a warp-uniform 128-bit
//    load) and tested by the 32 lanes against their own query -- no ballots, shuffles or partially
//    filled 32-candidate chunks in the inner loop.  Survivors (d2 <= fl(0.9999 R^2)) are appended
//    to a per-lane shared-memory list as 32-bit words (22 bits of d2 | staged slot); the scan uses the
//    FMA-contracted distance (6 instead of 8 FP32 instructions per candidate), every distance that is
//    returned or decides an order is recomputed with the defined, non-contracted formula.  Each lane
//    then loads its list into registers and runs a min/max sorting network that leaves the 64
//    smallest of up to 96 words in order; the exact (d2, index) pairs are rebuilt from the
//    staged candidates in that order, rows whose truncated keys collided are repaired by an
//    insertion sort on the exact order, and the rows are written coalesced through a
//    shared-memory transpose.  Lanes whose ball held fewer than k points, overflowed the network
//    or hit a truncated-distance tie between the k-th neighbour and a dropped key are queued for
//    knn_slow_kernel / search_list_kernel (the generic per-query routine, one warp per queued query);
//    half a warp failing the same way is re-queued as a group with a corrected radius instead.
//
//  * Clipped grids: when the queries of a call occupy a small box of the cloud, search_run indexes only
//    the points within a halo of that box; every ball radius is checked against GridView::rmax_safe and
//    the queries that exceed it are re-run on the full grid.
//
//  * search_big_kernel (k > 512): one CTA per query, keys in global scratch, block-wide bitonic sort.
//
//  * search_kernel (kNN with 64 < k <= 512, radius modes): one warp per query.  Per query:
//    (1) seed a radius from the local density of the 3x3x3 cell block, (2) scan the cells that
//    intersect the ball, keeping keys under the threshold in a per-warp shared-memory buffer,
//    (3) if the ball holds fewer than k points grow it, if it overflows the buffer shrink /
//    bisect the threshold, (4) pick the k smallest keys with a register-resident counting
//    bisection, (5) bitonic-sort them with warp shuffles and write the row.
//
// Roofline: algorithmic bytes/query = 12 (query) + 12 (data, once) + 8k (idx + d2) = 24 + 8k.
// Both kernels are issue bound (distance + select + sort), not HBM bound: DESIGN.md.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "eig3.cuh"
#include "ptx.cuh"
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
    uint32_t* nn_ptr;        // RADIUS_COUNT: per-row counts out; RADIUS_CSR: row offsets in; RADIUS with nn_ptr set: counts out AND
                             // the padded index rows hold their hits only (no -1 padding, no distances): the single-search CSR pair
    unsigned long long* stats;   // optional tile-kernel counters (PGEOF_KNN_STATS=1), else null
    uint32_t flags;          // debugging switches (PGEOF_KNN_FLAGS): 1 = volume-only radius seed, 2 = no radius retries
    uint2* slow_list;        // tile kernel: (query position, bits(radius hint)) of the queries left to knn_slow_kernel
    uint32_t* slow_count;
    uint2* unsafe_list;      // clipped grid: queries whose ball outgrew GridView::rmax_safe (re-run on the full grid)
    uint32_t* unsafe_count;
    // fused knn_features: the tile kernel turns its rows straight into features, the generic kernel parks its rows
    float* features;         // (n_query, 11) out
    uint32_t k_min;
    int eig_order;
    uint32_t* tmp_idx;       // [tmp_cap][k] neighbour rows of the queries the tile kernel queued
    uint32_t* tmp_rows;      // [tmp_cap] their row ids
    uint32_t tmp_cap;
    // queued queries whose ball spans many cells of this grid (sparse parts of a cloud whose cell edge was sized for its dense
    // surfaces) are deferred to a second run on a coarser grid: (query position, bits(start radius)) + counter, or null
    uint2* defer_list;
    uint32_t* defer_count;
    float r_split;           // start radius above which a queued query is deferred
};

// tile-kernel counters: why queries left the fast path, and how much work the fast path did
enum { ST_SHORT = 0, ST_OVER, ST_TIE, ST_REGION, ST_FIXED, ST_PASSES, ST_CANDS, ST_SURV, ST_N };

// ---------------------------------------------------------------------------------------
// generic per-query kNN (one warp): collect the keys under an adaptive threshold
// ---------------------------------------------------------------------------------------
// smallest geometric radius whose ball holds every point with key <= tau: tau_from_radius(radius_covering(tau)) >= tau
__device__ __forceinline__ float radius_covering(u64 tau)
{
    return __fmul_ru(__fsqrt_ru(__fdiv_ru(key_d2(tau), 0.9999f)), 1.00001f);
}

template <int CAP, bool PREFETCH = false>
__device__ __forceinline__ uint32_t knn_collect(const GridView& g, float qx, float qy, float qz, uint32_t k, float target,
                                                u64* keybuf, int lane, u64* tau_out, float r_hint = 0.f, bool* unsafe = nullptr)
{
    float R = r_hint;
    if (!(r_hint > 0.f)) {   // no hint (warp uniform): seed the ball from the density of the 3x3x3 block
        uint32_t n27, cells27;
        block27_count(g, qx, qy, qz, lane, &n27, &cells27);
        const float rho = fmaxf((float)n27, 1.f) / ((float)max(cells27, 1u) * g.hx * g.h * g.h);
        R = cbrtf(target / (4.18879f * rho)) + bbox_distance(g, qx, qy, qz);
    }
    float Rg = R, Rg_hi = 0.f;
    u64 tau_lo = 0, tau_hi = 0;
    bool have_lo = false, have_hi = false;
    int shrinks = 0;
    u64 tau = tau_from_radius(R);
    uint32_t c = 0;
    for (int it = 0; it < 512; ++it) {
        // a clipped grid (GridView::rmax_safe) only holds what balls up to that radius need: leave, the caller re-runs
        // the query on the full grid
        if (Rg > g.rmax_safe) { if (unsafe) *unsafe = true; *tau_out = 0; return 0; }
        c = scan_ball<CAP, true, PREFETCH>(g, qx, qy, qz, Rg, tau, keybuf, lane);
        if (c < k) {
            tau_lo = tau; have_lo = true;
            if (have_hi) { tau = tau_lo + (tau_hi - tau_lo) / 2; Rg = fminf(Rg_hi, radius_covering(tau)); }
            else {
                const float f = fminf(fmaxf(cbrtf(1.2f * target / fmaxf((float)c, 0.5f)), 1.2f), 2.5f);
                R *= f; Rg = R; tau = tau_from_radius(R);
            }
        } else if (c > (uint32_t)CAP) {
            tau_hi = tau; have_hi = true; Rg_hi = Rg;
            // dense spot: shrink the ball by the density estimate (twice at most) ...
            bool shrunk = false;
            if (!have_lo && shrinks < 2) {
                const float Rn = R * fminf(fmaxf(cbrtf(target / (float)c), 0.3f), 0.9f);
                const u64 tn = tau_from_radius(Rn);
                if (tn < tau_hi && (tn >> 32) != 0) { R = Rn; Rg = R; tau = tn; ++shrinks; shrunk = true; }
            }
            // ... else bisect the key space between tau_lo (0: nothing is below it) and tau_hi:
            // always converges because keys are distinct (many duplicates / exact ties land here)
            // (the rescan only has to cover the keys <= the new threshold: a smaller ball than the one that overflowed)
            if (!shrunk) { have_lo = true; tau = tau_lo + (tau_hi - tau_lo) / 2; Rg = fminf(Rg_hi, radius_covering(tau)); }
        } else break;
    }
    *tau_out = tau;
    return c;
}

// keybuf[0, c) holds keys <= tau, c <= CAP: pick the `need` smallest and return them sorted in v
// (element m * 32 + lane, padded with kKeyMax).
template <int NSORT, int CAP>
__device__ __forceinline__ void select_and_sort(u64* keybuf, uint32_t c, u64 tau, uint32_t need, int lane, u64 (&v)[NSORT / 32])
{
    constexpr int M = NSORT / 32, MC = CAP / 32;
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
    // Sort 32-bit keys (upper bits of d2 | position in keybuf) -- one SHFL and one min/max per element and stage where the
    // 64-bit keys need two SHFLs and a wide compare (the sort was 0.9 k of the 2.35 k instructions per query at k = 100) --
    // then fetch the exact 64-bit keys in that order.  Keys that share their truncated distance may come out swapped:
    // a few odd-even transposition rounds on the exact keys put isolated swaps right, anything longer (duplicates,
    // lattices) falls back to the 64-bit network.  The result is the exact (d2, index) order either way.
    constexpr uint32_t PB = NSORT <= 32 ? 5 : NSORT <= 64 ? 6 : NSORT <= 128 ? 7 : NSORT <= 256 ? 8 : 9, PM = (1u << PB) - 1u;
    uint32_t w[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const uint32_t e = m * 32 + lane;
        w[m] = e < c ? (((uint32_t)(keybuf[e] >> 32) & ~PM) | e) : 0xffffffffu;
    }
    warp_bitonic_sort32<M>(w, lane);
#pragma unroll
    for (int m = 0; m < M; ++m) v[m] = w[m] != 0xffffffffu ? keybuf[w[m] & PM] : kKeyMax;
    for (int it = 0; !warp_is_sorted<M>(v, lane); ++it) {
        if (it == 3) { warp_bitonic_sort<M>(v, lane); break; }
        warp_transpose_round<M>(v, lane);
    }
}

template <int M>
__device__ __forceinline__ void write_knn_row(const SearchArgs& a, uint32_t row, uint32_t k, const u64 (&v)[M], int lane)
{
    uint32_t* idx = reinterpret_cast<uint32_t*>(a.indices) + (size_t)row * k;
    float* d2 = a.sqr_dist + (size_t)row * k;
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const uint32_t e = m * 32 + lane;
        if (e < k) { idx[e] = key_idx(v[m]); d2[e] = key_d2(v[m]); }
    }
}

// ---------------------------------------------------------------------------------------
// one warp per query: kNN with k > 64 and the radius modes
// ---------------------------------------------------------------------------------------
template <int NSORT, int MODE>
__device__ __forceinline__ void search_one(const GridView& g, const SearchArgs& a, const float4 q4, u64* keybuf, int lane)
{
    using Cfg = SearchCfg<NSORT, MODE>;
    constexpr int M = Cfg::M, CAP = Cfg::CAP;
    const float qx = q4.x, qy = q4.y, qz = q4.z;
    const uint32_t row = __float_as_uint(q4.w);
    const uint32_t k = a.k;

    uint32_t c = 0;     // keys under `tau` (<= CAP once the loops below finish)
    u64 tau = 0;
    if (MODE == SEARCH_KNN) {
        c = knn_collect<CAP>(g, qx, qy, qz, k, a.target, keybuf, lane, &tau);
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
    select_and_sort<NSORT, CAP>(keybuf, c, tau, need, lane, v);

    // ---- write the row ----------------------------------------------------------------
    if (MODE == SEARCH_KNN) {
        write_knn_row<M>(a, row, k, v, lane);
    } else if (MODE == SEARCH_RADIUS) {
        int32_t* idx = reinterpret_cast<int32_t*>(a.indices) + (size_t)row * k;
        if (a.nn_ptr) {                                    // single-search CSR pair: the count + the hits, nothing else
            if (lane == 0) a.nn_ptr[row] = need;
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const uint32_t e = m * 32 + lane;
                if (e < need) idx[e] = (int32_t)key_idx(v[m]);
            }
            return;
        }
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
__global__ void __launch_bounds__(kWarps * 32) search_kernel(const GridView g, const SearchArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64* keybuf = reinterpret_cast<u64*>(smem_raw) + (size_t)warp * SearchCfg<NSORT, MODE>::CAP;
    const uint32_t w = blockIdx.x * kWarps + warp;
    if (w >= a.n_query) return;
    search_one<NSORT, MODE>(g, a, __ldg(a.queries + w), keybuf, lane);
}

// the same routine over the queries a tile kernel queued (radius mode)
template <int NSORT, int MODE>
__global__ void __launch_bounds__(kWarps * 32) search_list_kernel(const GridView g, const SearchArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64* keybuf = reinterpret_cast<u64*>(smem_raw) + (size_t)warp * SearchCfg<NSORT, MODE>::CAP;
    const uint32_t n = *a.slow_count;
    for (uint32_t w = blockIdx.x * kWarps + warp; w < n; w += gridDim.x * kWarps) {
        search_one<NSORT, MODE>(g, a, __ldg(a.queries + a.slow_list[w].x), keybuf, lane);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------
// one thread per query, one warp per 32 cell-sorted queries: kNN with k <= 64
// ---------------------------------------------------------------------------------------
// NOUT = sorted outputs kept per query (>= k), NEXTRA = further survivors the network can absorb.
template <int NOUT, int NEXTRA, int NWARPS = 2>
struct TileCfg {
    static constexpr int NLOAD = NOUT + NEXTRA;    // survivors a lane can sort
    static constexpr int SINK = 11;                // a batch of 8 appends may run 8 entries past the clamp (11: 16-B multiple)
    static constexpr int LCAP = NLOAD + 1 + SINK;  // list entries per lane
    static constexpr int STRIDE = 33;              // entry stride in words: lane-private walks AND column reads are conflict free
    static constexpr int WARPS = NWARPS;
    static constexpr int CMAX = NOUT > 64 ? 1016 : 872;   // candidates staged per pass (16 B each), multiple of 8
    static constexpr int PLANE_BYTES = NOUT * STRIDE * 4 + NOUT * 34 * 2;   // rolled rebuild: d2 plane + 16-bit staged-slot plane
    static constexpr int LIST_BYTES = ((LCAP * STRIDE * 4 > PLANE_BYTES ? LCAP * STRIDE * 4 : PLANE_BYTES) + 15) / 16 * 16;
    static constexpr int STAGE_BYTES = CMAX * 16;
    static constexpr int BAR_BYTES = 16 + 3 * 128;   // mbarrier + three per-lane words kept out of the register file (see below)
    static constexpr int SMEM_WARP_BYTES = STAGE_BYTES + LIST_BYTES + BAR_BYTES;
    static constexpr int MAX_PASSES = 12;          // passes ((y, z) rows straddled, halved regions, radius retries) before a warp falls back
    static constexpr int RETRY_MIN = 16;           // lanes that must fail the same way before they are re-queued with another radius
                                                   // (half a warp: uniform data never gets there, a surface or a cluster does at once)
    // a list entry / sort key is (bits(d2) & ~SLOT_MASK) | staged slot: 22 bits of distance order the
    // network, the slot finds the candidate again when the exact (d2, index) pair is rebuilt
    static constexpr uint32_t SLOT_BITS = 10;
    static constexpr uint32_t SLOT_MASK = (1u << SLOT_BITS) - 1;
    static constexpr uint32_t PAD = ~SLOT_MASK;    // key of an empty list position: above every real key (d2 bits of a NaN), slot 0
    static_assert(CMAX % 8 == 0 && CMAX <= (1 << SLOT_BITS), "slot field too small");
    static_assert(LIST_BYTES % 16 == 0 && STAGE_BYTES % 16 == 0, "alignment");
    static_assert(NOUT * STRIDE * 4 <= LIST_BYTES && (NOUT > 64 || NOUT * STRIDE * 4 <= STAGE_BYTES), "output planes must fit (NOUT > 64: rolled rebuild only, no index plane)");
    static_assert((1 << SLOT_BITS) * 16 <= SMEM_WARP_BYTES, "a padded slot must stay inside the warp's shared memory");
    static_assert(NOUT % 32 == 0 && NEXTRA % 32 == 0 && NEXTRA <= NOUT, "network shape");
};

// order preserving float <-> uint maps (for REDUX min / max)
__device__ __forceinline__ uint32_t f2o(float f) { const uint32_t b = __float_as_uint(f); return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u); }
__device__ __forceinline__ float o2f(uint32_t o) { return __uint_as_float(o ^ ((o >> 31) ? 0x80000000u : 0xffffffffu)); }

// Sorts the NLOAD keys of v and leaves the NOUT smallest, ascending, in v[0, NOUT).  Returns the
// smallest key that was dropped (0xffffffff if none): 32-key blocks by odd-even merge sort, blocks
// merged pairwise, then one half-cleaner against the (reversed) extra run and a bitonic merge.
template <int NOUT, int NEXTRA>
__device__ __forceinline__ uint32_t select_sort_network(uint32_t (&v)[NOUT + NEXTRA])
{
    constexpr int N = NOUT + NEXTRA;
    RegOddEvenSort<N, 0, NOUT>::run(v);
    RegOddEvenSort<N, NOUT, NEXTRA>::run(v);
    uint32_t dropped = 0xffffffffu;
#pragma unroll
    for (int i = 0; i < NEXTRA; ++i) {
        const uint32_t a = v[NOUT - 1 - i], b = v[NOUT + i];
        v[NOUT - 1 - i] = min(a, b);
        dropped = min(dropped, max(a, b));
    }
    RegBitonicMerge<N, 0, NOUT>::run(v);
    return dropped;
}

// LOCK: the warps of a CTA enter the straight-line sections (key load + sorting network, exact rebuild) together, so that
// one instruction-cache fill serves all of them (profiles/r2_summary.md: the GPC-level instruction cache runs at 93 % of
// its request rate when every warp streams the 62 KB of straight-line code on its own)
#define PGEOF_LOCK_BARRIER() do { if (LOCK) asm volatile("bar.sync 0;" ::: "memory"); } while (0)
// ROLLED: the exact rebuild is a rolled loop over sorted keys parked in shared memory and rows keep (d2, staged slot) pairs
// (the index is read from the staged candidate when the row is written): 26 KB less straight-line code per pass and 128
// instead of 255 registers, but the rolled loop has a quarter of the loads in flight: GCC requests 44 -> 33 M, kernel
// 7.5 -> 8.2 ms (profiles/r2_summary.md).  Off by default; PGEOF_KNN_ROLLED=1 selects it for 32 < k <= 64.
template <int NOUT, int NEXTRA, int NWARPS, int MODE, bool FUSED = false, bool LOCK = false, bool ROLLED = false>
__global__ void __launch_bounds__(NWARPS * 32) knn_tile_kernel(const GridView g, const SearchArgs a)
{
    static_assert(MODE == SEARCH_KNN || MODE == SEARCH_RADIUS, "tile kernel: kNN or padded radius search");
    static_assert(!FUSED || MODE == SEARCH_KNN, "fused features: kNN only");
    using Cfg = TileCfg<NOUT, NEXTRA, NWARPS>;
    constexpr int M = NOUT / 32, NLOAD = Cfg::NLOAD, S = Cfg::STRIDE;
    constexpr int PS = 34;                         // stride (in 16-bit entries) of the staged-slot plane: lane-private walks and column reads conflict free
    constexpr bool SLOTS = FUSED || ROLLED;        // rows are (d2, staged slot) pairs until they are written
    static_assert(NOUT * S * 4 + NOUT * PS * 2 <= Cfg::LIST_BYTES, "d2 plane + slot plane must fit in the list area");
    extern __shared__ __align__(128) unsigned char smem_tile[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* wsm = smem_tile + (size_t)warp * Cfg::SMEM_WARP_BYTES;
    float4* stage = reinterpret_cast<float4*>(wsm);
    uint32_t* list = reinterpret_cast<uint32_t*>(wsm + Cfg::STAGE_BYTES);
    uint64_t* bar = reinterpret_cast<uint64_t*>(wsm + Cfg::STAGE_BYTES + Cfg::LIST_BYTES);
    const uint32_t base = (blockIdx.x * Cfg::WARPS + warp) * 32u;
    if (!LOCK && base >= a.n_query) return;     // (LOCK: a warp without queries still meets the barriers)
    const uint32_t k = a.k;
    if (lane == 0) { ptx::mbarrier_init(bar, 1); ptx::fence_mbarrier_init(); }
    __syncwarp();
    uint32_t parity = 0;

    const bool valid = base + lane < a.n_query;
    const float4 q4 = valid ? __ldg(a.queries + base + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float qx = q4.x, qy = q4.y, qz = q4.z;
    const uint32_t row = __float_as_uint(q4.w);
    const int cqx = cell_coord(qx, g.lo[0], g.inv_hx, g.n[0]);
    const int cqy = cell_coord(qy, g.lo[1], g.inv_h, g.n[1]);
    const int cqz = cell_coord(qz, g.lo[2], g.inv_h, g.n[2]);
    const uint32_t rowid = (uint32_t)cqz * (uint32_t)g.n[1] + (uint32_t)cqy;

    unsigned remaining = __ballot_sync(kFull, valid);
    unsigned slow = 0;                       // lanes finished by the generic routine
    unsigned unsafe = 0;                     // lanes whose ball outgrew a clipped grid
    // per-lane state that lives across passes sits in shared memory: the sorting network leaves no register to spare
    float* const s_hint = reinterpret_cast<float*>(bar + 2) + lane;   // radius the generic routine starts from (0: local density)
    // feedback for data the uniform-density seed misjudges (surfaces, lines, clusters): lanes whose ball held too few /
    // too many points are re-queued ONCE OR TWICE as a group with a radius factor derived from the count they saw
    float* const s_rmul = s_hint + 32;       // this lane's factor on the density-seeded radius
    uint32_t* const s_retry = reinterpret_cast<uint32_t*>(s_hint + 64);   // bits 0-1: re-queues so far, bits 2-3: class (0 first try, 1 grow, 2 shrink)
    *s_hint = 0.f; *s_rmul = 1.f; *s_retry = 0u;
    // LOCK: one round of the CTA = every warp's pass.  The barrier in front of the key load doubles as the vote on another
    // round (does any warp still hold queries for a later pass?), so a round costs two barriers, not three.
    bool more = true;
#define PGEOF_LOCK_VOTE() do { if (LOCK) more = __syncthreads_or(remaining != 0) != 0; } while (0)
    for (int pass = 0; LOCK ? more : remaining != 0; ++pass) {
        if (LOCK && !remaining) { PGEOF_LOCK_VOTE(); PGEOF_LOCK_BARRIER(); continue; }
        // ---- the lanes of this pass: queries in the (y, z) cell row of the first remaining lane ----
        const int leader = __ffs(remaining) - 1;
        const uint32_t lrow = __shfl_sync(kFull, rowid, leader);
        const uint32_t cls = *s_retry >> 2;
        const uint32_t lcls = __shfl_sync(kFull, cls, leader);
        unsigned active = __ballot_sync(kFull, ((remaining >> lane) & 1u) && rowid == lrow && cls == lcls);
        remaining &= ~active;
        if (pass >= Cfg::MAX_PASSES) {
            slow |= active;
            if (a.stats && lane == 0) atomicAdd(a.stats + ST_REGION, (unsigned long long)__popc(active));
            PGEOF_LOCK_VOTE(); PGEOF_LOCK_BARRIER();
            continue;
        }

        // ---- region of the active queries; a region too large to stage is halved along x ---------
        float R = 0.f;
        uint32_t s = 0, len = 0, off = 0, C = 0, C8 = 0;
        bool mine = false;
        for (int split = 0;; ++split) {
            mine = (active >> lane) & 1u;
            // bounding box of the active queries
            const float xmin = o2f(__reduce_min_sync(kFull, mine ? f2o(qx) : 0xffffffffu));
            const float xmax = o2f(__reduce_max_sync(kFull, mine ? f2o(qx) : 0u));
            const float ymin = o2f(__reduce_min_sync(kFull, mine ? f2o(qy) : 0xffffffffu));
            const float ymax = o2f(__reduce_max_sync(kFull, mine ? f2o(qy) : 0u));
            const float zmin = o2f(__reduce_min_sync(kFull, mine ? f2o(qz) : 0xffffffffu));
            const float zmax = o2f(__reduce_max_sync(kFull, mine ? f2o(qz) : 0u));
            const int cxa = __reduce_min_sync(kFull, mine ? cqx : 0x7fffffff);
            const int cxb = __reduce_max_sync(kFull, mine ? cqx : -1);
            const int cy = __shfl_sync(kFull, cqy, leader), cz = __shfl_sync(kFull, cqz, leader);

            if (MODE == SEARCH_RADIUS) {
                // radius search: the ball is given (nn_search.hpp:98: r^2 = fl(r * r) decides, strictly)
                const float r2 = __fmul_rn(a.radius, a.radius);
                R = __fmul_ru(__fsqrt_ru(r2), 1.0001f);
            } else
            // density and SHAPE of the block around them -> search radius R.  The 3 x 3 (y, z) rows of the block hold
            // counts c_l; m = (sum c)^2 / sum c^2 is the number of rows the points effectively occupy: 9 for a filled
            // volume, 3 for a surface, 1 for a line along x.  The ball is sized for a structure of dimension
            // d = 1 + log3(m) whose density is that of the occupied rows (a uniform cloud gives d = 2.99; next to the
            // domain boundary the empty outside rows lower d and enlarge the ball, as they should).
            {
                const int bx0 = max(cxa - g.xf, 0), bx1 = min(cxb + g.xf, g.n[0] - 1);
                float c = 0.f;
                if (lane < 9) {
                    const int by = cy + lane % 3 - 1, bz = cz + lane / 3 - 1;
                    if (by >= 0 && by < g.n[1] && bz >= 0 && bz < g.n[2]) {
                        const uint32_t rb = ((uint32_t)bz * (uint32_t)g.n[1] + (uint32_t)by) * (uint32_t)g.n[0];
                        c = (float)(__ldg(g.cell_start + rb + bx1 + 1) - __ldg(g.cell_start + rb + bx0));
                    }
                }
                float s1 = c, s2 = c * c;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) { s1 += __shfl_xor_sync(kFull, s1, o); s2 += __shfl_xor_sync(kFull, s2, o); }
                s1 = fmaxf(__shfl_sync(kFull, s1, 0), 1.f);
                s2 = fmaxf(__shfl_sync(kFull, s2, 0), 1.f);
                const float m = fminf(fmaxf(s1 * s1 / s2, 1.f), 9.f);
                const float d = 1.f + log2f(m) * 0.6309298f;                               // 1 / log2(3)
                const float rho = s1 / (m * (float)(bx1 - bx0 + 1) * g.hx * g.h * g.h);     // density inside the occupied rows
                const float T = a.target;
                const float l3 = log2f(T / (4.18879f * rho)) * (1.f / 3.f);                 // volume:  4/3 pi R^3 rho = T
                const float l2 = log2f(T / (3.14159265f * rho * g.h)) * 0.5f;               // surface: pi R^2 (rho h) = T
                const float l1 = log2f(T / (2.f * rho * g.h * g.h));                        // line:    2 R (rho h^2) = T
                float lr = d >= 2.f ? (d - 2.f) * l3 + (3.f - d) * l2 : (d - 1.f) * l2 + (2.f - d) * l1;
                if (a.flags & 1u) lr = log2f(T / (4.18879f * s1 / (9.f * (float)(bx1 - bx0 + 1) * g.hx * g.h * g.h))) * (1.f / 3.f);
                // positive floats order like their bit patterns
                const float rm = __uint_as_float(__reduce_max_sync(kFull, mine ? __float_as_uint(*s_rmul) : 0u));
                R = exp2f(lr) * rm;
            }

            if (R > g.rmax_safe) { unsafe |= active; active = 0; break; }   // clipped grid: not every point of this ball is indexed
            // candidate region: cells meeting the dilated box; one contiguous span per (y, z) row
            const int cx0 = cell_coord(__fsub_rd(xmin, R), g.lo[0], g.inv_hx, g.n[0]);
            const int cx1 = cell_coord(__fadd_ru(xmax, R), g.lo[0], g.inv_hx, g.n[0]);
            const int cy0 = cell_coord(__fsub_rd(ymin, R), g.lo[1], g.inv_h, g.n[1]);
            const int cy1 = cell_coord(__fadd_ru(ymax, R), g.lo[1], g.inv_h, g.n[1]);
            const int cz0 = cell_coord(__fsub_rd(zmin, R), g.lo[2], g.inv_h, g.n[2]);
            const int cz1 = cell_coord(__fadd_ru(zmax, R), g.lo[2], g.inv_h, g.n[2]);
            const uint32_t nyr = (uint32_t)(cy1 - cy0 + 1);
            const uint32_t nrows = nyr * (uint32_t)(cz1 - cz0 + 1);
            bool fits = nrows <= 32u;
            if (fits) {
                s = 0; len = 0;
                if ((uint32_t)lane < nrows) {
                    const int rz = cz0 + (int)((uint32_t)lane / nyr), ry = cy0 + (int)((uint32_t)lane % nyr);
                    const uint32_t rb = ((uint32_t)rz * (uint32_t)g.n[1] + (uint32_t)ry) * (uint32_t)g.n[0];
                    s = __ldg(g.cell_start + rb + cx0);
                    len = __ldg(g.cell_start + rb + cx1 + 1) - s;
                }
                // exclusive prefix of the span lengths = where each row lands in the staging area
                off = len;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(kFull, off, o);
                    if (lane >= o) off += t;
                }
                C = __shfl_sync(kFull, off, 31);
                off -= len;
                C8 = (C + 7u) & ~7u;
                fits = C8 <= (uint32_t)Cfg::CMAX;
            }
            if (fits) break;
            const int na = __popc(active);
            if (na < 2 || split >= 3 || nrows > 32u) {   // dense spot: the generic routine adapts its ball
                slow |= active;
                if (a.stats && lane == 0) atomicAdd(a.stats + ST_REGION, (unsigned long long)na);
                active = 0;
                break;
            }
            // keep the lower half of the lanes (cell-sorted, so the lower half along x); the rest waits for a later pass
            const unsigned keep = __ballot_sync(kFull, mine && __popc(active & lanemask_lt()) < na / 2);
            remaining |= active & ~keep;
            active = keep;
        }
        if (!active) { PGEOF_LOCK_VOTE(); PGEOF_LOCK_BARRIER(); continue; }

        // ---- stage the region: one 1-D TMA bulk copy per row, all rows in flight at once ---------
        __syncwarp();
        if (lane == 0) { ptx::fence_proxy_async_smem(); ptx::mbarrier_arrive_expect_tx(bar, C * 16u); }
        __syncwarp();
        if (len) ptx::bulk_g2s(stage + off, g.pts + s, len * 16u, bar);
        // pad to a multiple of 8 with records no query accepts (d2 = +inf)
        if (C + (uint32_t)lane < C8) stage[C + lane] = make_float4(__int_as_float(0x7f800000), 0.f, 0.f, 0.f);
        ptx::mbarrier_wait(bar, parity);
        parity ^= 1u;
        __syncwarp();

        // ---- scan: every candidate is read once per warp (broadcast) and tested by all lanes ------
        // accept iff d2 <= t2 with t2 strictly inside R^2, so every accepted point is in the region
        // radius mode: everything whose DEFINED distance is < r^2 must pass the fused filter -> r^2 (1 + 2^-18), rounded up
        const float r2 = __fmul_rn(a.radius, a.radius);
        const float t2 = !mine ? -1.f : (MODE == SEARCH_RADIUS ? (r2 > 0.f ? __fmul_ru(r2, 1.0000038147f) : -1.f) : __fmul_rd(__fmul_rd(R, R), 0.9999f));
        const float t2_safe = __fmul_rd(t2, 0.99999618530273f);   // t2 (1 - 2^-18)
        uint32_t* const wbase = list + lane;
        // shared-memory byte address of the lane's next free entry: advances by one stride per survivor;
        // clamped once per batch of 8 to entry NLOAD + 1 (so the count saturates at NLOAD + 1 = overflow)
        const uint32_t waddr0 = ptx::smem_addr(wbase);
        const uint32_t wclamp = waddr0 + (NLOAD + 1) * S * 4;
        uint32_t waddr = waddr0;
        // branch-free append (the compiler turns the equivalent C++ into a divergent branch per candidate)
#define PGEOF_TILE_APPEND(D2, SLOT)                                                           \
        asm volatile("{\n\t.reg .pred p;\n\t"                                                 \
                     "setp.le.f32 p, %1, %2;\n\t"                                             \
                     "@p st.shared.u32 [%0], %3;\n\t"                                         \
                     "@p add.u32 %0, %0, %4;\n\t}"                                            \
                     : "+r"(waddr)                                                            \
                     : "f"(D2), "f"(t2), "r"((__float_as_uint(D2) & ~Cfg::SLOT_MASK) | (SLOT)), "n"(S * 4));
        for (uint32_t c = 0; c < C8; c += 8) {
            float d[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { const float4 p = stage[c + i]; d[i] = sqdist_fused(qx, qy, qz, p.x, p.y, p.z); }
#pragma unroll
            for (int i = 0; i < 8; ++i) PGEOF_TILE_APPEND(d[i], c + i)
            waddr = min(waddr, wclamp);
        }
#undef PGEOF_TILE_APPEND
        __syncwarp();
        const uint32_t cnt = (waddr - waddr0) / (S * 4);
        bool ok = mine && (MODE == SEARCH_RADIUS || cnt >= k) && cnt <= (uint32_t)NLOAD;
        // radius the generic routine starts from if this lane leaves the fast path: scaled by the count seen here
        if (mine) *s_hint = R * (cnt < k ? fminf(cbrtf(1.25f * a.target / fmaxf((float)cnt, 2.f)), 3.f) : (cnt > (uint32_t)NLOAD ? 0.88f : 1.f));
        // re-queue groups of lanes the seeded radius failed (>= RETRY_MIN of them: a pass costs as much as ~8 generic queries)
        unsigned requeued = 0;
        if (MODE == SEARCH_KNN && pass + 2 < Cfg::MAX_PASSES && !(a.flags & 2u)) {
            const uint32_t retry = *s_retry;
            const bool can = mine && (retry & 3u) < 2u;
            const unsigned sh = __ballot_sync(kFull, can && cnt < k), ov = __ballot_sync(kFull, can && cnt > (uint32_t)NLOAD);
            if (__popc(sh) >= Cfg::RETRY_MIN) {
                if ((sh >> lane) & 1u) {
                    // count ~ R^d with d between 2 (surface) and 3 (volume): exponent 1/2.5
                    *s_rmul *= fminf(fmaxf(exp2f(0.4f * log2f(1.3f * a.target / fmaxf((float)cnt, 1.f))), 1.15f), 3.f);
                    *s_retry = ((retry & 3u) + 1u) | (1u << 2);
                }
                requeued |= sh;
            }
            if (__popc(ov) >= Cfg::RETRY_MIN) {
                if ((ov >> lane) & 1u) { *s_rmul *= 0.7f; *s_retry = ((retry & 3u) + 1u) | (2u << 2); }
                requeued |= ov;
            }
            remaining |= requeued;
        }
        if (a.stats) {
            const unsigned sh = __ballot_sync(kFull, mine && cnt < k), ov = __ballot_sync(kFull, mine && cnt > (uint32_t)NLOAD);
            const uint32_t sv = __reduce_add_sync(kFull, mine ? cnt : 0u);
            if (lane == 0) {
                atomicAdd(a.stats + ST_SHORT, (unsigned long long)__popc(sh));
                atomicAdd(a.stats + ST_OVER, (unsigned long long)__popc(ov));
                atomicAdd(a.stats + ST_PASSES, 1ull);
                atomicAdd(a.stats + ST_CANDS, (unsigned long long)C);
                atomicAdd(a.stats + ST_SURV, (unsigned long long)sv);
            }
        }

        // ---- sort: one thread per row, 32-bit keys in registers ---------------------------------
        uint32_t need = k;                     // entries of the row that hold neighbours (radius mode: the rest is padding)
        uint32_t dropped;
        uint32_t bad[M] = {};                  // bit i: sorted entry i precedes entry i - 1 in the exact order
        uint32_t nvalid = 0;                   // radius mode: kept entries whose defined distance is < r^2
        {
            uint32_t v[NLOAD];
            PGEOF_LOCK_VOTE();
#pragma unroll
            for (int i = 0; i < NLOAD; ++i) {
                const uint32_t w = wbase[i * S];
                v[i] = (uint32_t)i < cnt ? w : Cfg::PAD;   // sorts behind every real key, and its slot field (0) is a staged candidate
            }
            __syncwarp();                      // the lists are dead: their memory becomes the d2 plane
            dropped = select_sort_network<NOUT, NEXTRA>(v);
            PGEOF_LOCK_BARRIER();
            // rebuild the exact (d2, index) pairs in sorted order.  Keys that share their truncated
            // distance may be out of order: remember it, the row is repaired in shared memory below.
            // plane[i * S + lane] = i-th neighbour of this lane's row
            uint32_t* plane_d = list + lane;
            uint32_t pd = 0, pi = 0;
            if (ROLLED) {
                uint16_t* pslot = reinterpret_cast<uint16_t*>(list + NOUT * S) + lane;
#pragma unroll
                for (int i = 0; i < NOUT; ++i) plane_d[i * S] = v[i];       // sorted keys: the registers are free from here on
#pragma unroll
                for (int r = 0; r < M; ++r) {
                    uint32_t b = 0;
#pragma unroll 16
                    for (int j = 0; j < 32; ++j) {
                        const int i = r * 32 + j;
                        const uint32_t slot = plane_d[i * S] & Cfg::SLOT_MASK;
                        const float4 p = stage[slot];
                        const uint32_t d = __float_as_uint(sqdist_f32(qx, qy, qz, p.x, p.y, p.z));
                        const uint32_t id = __float_as_uint(p.w);
                        b |= (i > 0 && (uint32_t)i < cnt && (d < pd || (d == pd && id < pi))) ? (1u << j) : 0u;
                        pd = d; pi = id;
                        if (MODE == SEARCH_RADIUS) nvalid += ((uint32_t)i < cnt && d < __float_as_uint(r2)) ? 1u : 0u;
                        plane_d[i * S] = d;
                        pslot[i * PS] = (uint16_t)slot;
                    }
                    bad[r] = b;
                }
            } else {
#pragma unroll
            for (int i = 0; i < NOUT; ++i) {
                const float4 p = stage[v[i] & Cfg::SLOT_MASK];
                const uint32_t d = __float_as_uint(sqdist_f32(qx, qy, qz, p.x, p.y, p.z));
                const uint32_t id = __float_as_uint(p.w);
                if (i > 0) bad[i / 32] |= ((uint32_t)i < cnt && (d < pd || (d == pd && id < pi))) ? (1u << (i % 32)) : 0u;
                pd = d; pi = id;
                if (MODE == SEARCH_RADIUS) nvalid += ((uint32_t)i < cnt && d < __float_as_uint(r2)) ? 1u : 0u;
                plane_d[i * S] = d;
                if (FUSED) reinterpret_cast<uint16_t*>(list + NOUT * S)[i * PS + lane] = (uint16_t)(v[i] & Cfg::SLOT_MASK);
                else v[i] = id;
            }
            if (!FUSED) {
                __syncwarp();                  // the staged candidates are dead: index plane
                uint32_t* plane_i = reinterpret_cast<uint32_t*>(stage) + lane;
#pragma unroll
                for (int i = 0; i < NOUT; ++i) plane_i[i * S] = v[i];
            }
            }
        }
        {
            uint32_t* plane_d = list + lane;
            uint32_t* plane_i = reinterpret_cast<uint32_t*>(stage) + lane;
            uint32_t any_bad = 0;
#pragma unroll
            for (int r = 0; r < M; ++r) any_bad |= bad[r];
            if (SLOTS && ok && any_bad) {
                // same repair on (d2, staged slot) pairs: the candidates stay staged, their index is read when two distances tie
                uint16_t* pslot = reinterpret_cast<uint16_t*>(list + NOUT * S) + lane;
                uint32_t first = 0, last = 0;
#pragma unroll
                for (int r = M - 1; r >= 0; --r) if (bad[r]) first = r * 32 + __ffs(bad[r]) - 1;
#pragma unroll
                for (int r = 0; r < M; ++r) if (bad[r]) last = r * 32 + 31 - __clz(bad[r]);
                const uint32_t nfix = min(cnt, (uint32_t)NOUT);
                for (uint32_t i = first; i < nfix; ++i) {
                    const uint32_t d = plane_d[i * S];
                    const uint16_t sl = pslot[i * PS];
                    const uint32_t id = __float_as_uint(stage[sl].w);
                    uint32_t j = i;
                    while (j > 0) {
                        const uint32_t qd = plane_d[(j - 1) * S];
                        const uint16_t qs = pslot[(j - 1) * PS];
                        if (qd < d || (qd == d && __float_as_uint(stage[qs].w) < id)) break;
                        plane_d[j * S] = qd; pslot[j * PS] = qs;
                        --j;
                    }
                    if (j != i) { plane_d[j * S] = d; pslot[j * PS] = sl; }
                    else if (i > last) break;
                }
            }
            if (!SLOTS && ok && any_bad) {
                // insertion sort on the exact (d2, index) order from the first inversion on; past the last
                // inversion the first entry found in place ends it (everything behind it is in order already)
                uint32_t first = 0, last = 0;
#pragma unroll
                for (int r = M - 1; r >= 0; --r) if (bad[r]) first = r * 32 + __ffs(bad[r]) - 1;
#pragma unroll
                for (int r = 0; r < M; ++r) if (bad[r]) last = r * 32 + 31 - __clz(bad[r]);
                const uint32_t nfix = min(cnt, (uint32_t)NOUT);
                for (uint32_t i = first; i < nfix; ++i) {
                    const uint32_t d = plane_d[i * S], id = plane_i[i * S];
                    uint32_t j = i;
                    while (j > 0) {
                        const uint32_t qd = plane_d[(j - 1) * S], qi = plane_i[(j - 1) * S];
                        if (qd < d || (qd == d && qi < id)) break;
                        plane_d[j * S] = qd; plane_i[j * S] = qi;
                        --j;
                    }
                    if (j != i) { plane_d[j * S] = d; plane_i[j * S] = id; }
                    else if (i > last) break;
                }
            }
            // The scan ordered and accepted candidates by the FUSED distance, which is within 2^-21 (relative) of
            // the defined one: at most one truncation bucket (2^-13) off.  The row is exact iff
            //  (a) no dropped key can precede the k-th neighbour: bucket(dropped) >= bucket(k-th) + 2, and
            //  (b) no rejected candidate can: d2(k-th) <= t2 (1 - 2^-18) < defined d2 of anything rejected.
            bool tie;
            if (MODE == SEARCH_RADIUS) {
                // rows: the min(nvalid, k) nearest.  Keys were only dropped if more than NOUT passed the filter; then the
                // row must be full (k valid entries) and the dropped keys two buckets above its last entry, as in (a)
                need = min(nvalid, k);
                const uint32_t kth = plane_d[(max(need, 1u) - 1) * S];
                tie = ok && dropped < Cfg::PAD && (need < k || (dropped >> Cfg::SLOT_BITS) < (kth >> Cfg::SLOT_BITS) + 2u);
            } else {
                const uint32_t kth = plane_d[(k - 1) * S];
                tie = ok && ((dropped < Cfg::PAD && (dropped >> Cfg::SLOT_BITS) < (kth >> Cfg::SLOT_BITS) + 2u) ||
                             !(__uint_as_float(kth) <= t2_safe));
            }
            if (tie) ok = false;
            if (a.stats) {
                const unsigned tm = __ballot_sync(kFull, tie), fm = __ballot_sync(kFull, ok && any_bad);
                if (lane == 0) {
                    atomicAdd(a.stats + ST_TIE, (unsigned long long)__popc(tm));
                    atomicAdd(a.stats + ST_FIXED, (unsigned long long)__popc(fm));
                }
            }
        }
        slow |= __ballot_sync(kFull, mine && !ok) & ~requeued;
        __syncwarp();
        if (FUSED) {
            // ---- fused knn_features: moments of the k nearest straight from the staged candidates, in the exact
            // order compute_features would walk them (same origin, same operations: the result is bit identical),
            // eigen solve + 11 formulas per lane, rows staged in shared memory and written as 44-byte runs
            float f[11];
#pragma unroll
            for (int i = 0; i < 11; ++i) f[i] = 0.f;
            if (ok && k >= a.k_min) {
                const uint16_t* pslot = reinterpret_cast<const uint16_t*>(list + NOUT * S) + lane;
                const float4 o = stage[pslot[0]];
                Moments m;
#pragma unroll 4
                for (uint32_t i = 1; i < k; ++i) {
                    const float4 p = stage[pslot[i * PS]];
                    m.add(p.x - o.x, p.y - o.y, p.z - o.z);
                }
                features11<float>(m.pca(k, a.eig_order), f);
            }
            __syncwarp();                      // the planes are dead
            float* s_feat = reinterpret_cast<float*>(list);
#pragma unroll
            for (int i = 0; i < 11; ++i) s_feat[lane * 11 + i] = f[i];
            __syncwarp();
            const unsigned okm = __ballot_sync(kFull, ok);
#pragma unroll 4
            for (int q = 0; q < 32; ++q) {
                const uint32_t rowq = __shfl_sync(kFull, row, q);
                if (((okm >> q) & 1u) && lane < 11) a.features[(size_t)rowq * 11 + lane] = s_feat[q * 11 + lane];
            }
        } else
        // ---- write the finished rows, one coalesced row at a time ---------------------------------
        {
            const uint32_t* plane_d = list;
            const uint32_t* plane_i = reinterpret_cast<const uint32_t*>(stage);
            const unsigned okm = __ballot_sync(kFull, ok);
            if (MODE == SEARCH_RADIUS && a.nn_ptr) {
                // single-search CSR pair (radius_search_csr): the row's count goes straight into the offsets array and only
                // the hits are written -- no -1 / 0 padding, no distances (the padded table is scratch, compacted afterwards)
                if (ok) a.nn_ptr[row] = need;
#pragma unroll 8
                for (int q = 0; q < 32; ++q) {
                    const uint32_t rowq = __shfl_sync(kFull, row, q);
                    const uint32_t needq = ((okm >> q) & 1u) ? __shfl_sync(kFull, need, q) : 0u;
                    uint32_t* idx = reinterpret_cast<uint32_t*>(a.indices) + (size_t)rowq * k + lane;
#pragma unroll
                    for (int r = 0; r < M; ++r) {
                        const uint32_t e = r * 32 + lane;
                        uint32_t vi;
                        if (SLOTS) vi = __float_as_uint(stage[reinterpret_cast<const uint16_t*>(list + NOUT * S)[e * PS + q]].w);
                        else vi = plane_i[e * S + q];
                        if (e < needq) idx[r * 32] = vi;
                    }
                }
            } else
            // branch free: predicated stores, one 64-bit row offset per row
#pragma unroll 8
            for (int q = 0; q < 32; ++q) {
                const uint32_t rowq = __shfl_sync(kFull, row, q);
                const uint32_t needq = MODE == SEARCH_RADIUS ? __shfl_sync(kFull, need, q) : k;
                const size_t o = (size_t)rowq * k + lane;
                uint32_t* idx = reinterpret_cast<uint32_t*>(a.indices) + o;
                float* d2 = a.sqr_dist + o;
#pragma unroll
                for (int r = 0; r < M; ++r) {
                    const uint32_t e = r * 32 + lane;
                    uint32_t vi, vd = plane_d[e * S + q];
                    if (SLOTS) vi = __float_as_uint(stage[reinterpret_cast<const uint16_t*>(list + NOUT * S)[e * PS + q]].w);
                    else vi = plane_i[e * S + q];
                    if (MODE == SEARCH_RADIUS && e >= needq) { vi = 0xffffffffu; vd = 0u; }   // pad: -1 / 0 (nn_search.hpp:104,108)
                    const uint32_t on = (((okm >> q) & 1u) && e < k) ? 1u : 0u;
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %4, 0;\n\t"
                                 "@p st.global.u32 [%0], %2;\n\t@p st.global.u32 [%1], %3;\n\t}"
                                 :: "l"(idx + r * 32), "l"(d2 + r * 32), "r"(vi), "r"(vd), "r"(on) : "memory");
                }
            }
        }
        __syncwarp();
    }

    // ---- queue the lanes the tile path could not finish for knn_slow_kernel --------------------
    // (kept out of this kernel: the generic routine is large, rarely needed and latency bound; inlined here
    // it evicted the sorting network from the instruction cache and ran at 1/8 occupancy)
    if (slow) {
        uint32_t at = 0;
        if (lane == 0) at = atomicAdd(a.slow_count, (uint32_t)__popc(slow));
        at = __shfl_sync(kFull, at, 0);
        if ((slow >> lane) & 1u) a.slow_list[at + __popc(slow & lanemask_lt())] = make_uint2(base + lane, __float_as_uint(*s_hint));
    }
    if (unsafe) {
        uint32_t at = 0;
        if (lane == 0) at = atomicAdd(a.unsafe_count, (uint32_t)__popc(unsafe));
        at = __shfl_sync(kFull, at, 0);
        if ((unsafe >> lane) & 1u) a.unsafe_list[at + __popc(unsafe & lanemask_lt())] = make_uint2(base + lane, 0u);
    }
}

// generic per-query routine (one warp per query) over the queries the tile kernel queued
template <int NOUT>
__global__ void __launch_bounds__(kWarps * 32) knn_slow_kernel(const GridView g, const SearchArgs a)
{
    constexpr int M = NOUT / 32, CAP = 256;
    __shared__ u64 keys[kWarps][CAP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64* keybuf = keys[warp];
    const uint32_t n = *a.slow_count, k = a.k;
    for (uint32_t w = blockIdx.x * kWarps + warp; w < n; w += gridDim.x * kWarps) {
        const uint2 rec = a.slow_list[w];
        const float4 q4 = __ldg(a.queries + rec.x);
        float r_start = __uint_as_float(rec.y);
        if (a.defer_list) {
            if (!(r_start > 0.f)) {   // no hint (the region did not fit the tile kernel at all): the density seed of knn_collect
                uint32_t n27, cells27;
                block27_count(g, q4.x, q4.y, q4.z, lane, &n27, &cells27);
                const float rho = fmaxf((float)n27, 1.f) / ((float)max(cells27, 1u) * g.hx * g.h * g.h);
                r_start = cbrtf(a.target / (4.18879f * rho)) + bbox_distance(g, q4.x, q4.y, q4.z);
            }
            if (r_start > a.r_split) {   // warp uniform: every row of cells this ball crosses holds a handful of points here
                if (lane == 0) a.defer_list[atomicAdd(a.defer_count, 1u)] = make_uint2(rec.x, __float_as_uint(r_start));
                continue;
            }
        }
        u64 tau;
        bool unsafe = false;
        const uint32_t c = knn_collect<CAP, true>(g, q4.x, q4.y, q4.z, k, a.target, keybuf, lane, &tau, r_start, &unsafe);
        if (unsafe) {   // warp uniform
            if (lane == 0) a.unsafe_list[atomicAdd(a.unsafe_count, 1u)] = make_uint2(rec.x, 0u);
            continue;
        }
        u64 v[M];
        select_and_sort<NOUT, CAP>(keybuf, c, tau, k, lane, v);
        if (a.tmp_idx) {   // fused knn_features: park the row (indices only) for the feature pass over the queued queries
            if (w < a.tmp_cap) {
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const uint32_t e = m * 32 + lane;
                    if (e < k) a.tmp_idx[(size_t)w * k + e] = key_idx(v[m]);
                }
                if (lane == 0) a.tmp_rows[w] = __float_as_uint(q4.w);
            }
        } else write_knn_row<M>(a, __float_as_uint(q4.w), k, v, lane);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------
// TWO LANES PER QUERY, one warp per 16 cell-sorted queries: kNN with k <= 128
// ---------------------------------------------------------------------------------------
// The thread-per-query tile kernel above is pinned at 8 warps per SM twice over: 255 registers for its 96-key sorting
// network and 28.6 KB of shared memory per warp.  Here a query owns the lane pair (q, q + 16):
//   * the region of 16 queries is a third smaller than that of 32 and each lane scans every other staged candidate
//     (one LDS.128 serves the two half warps), appending survivors to its own list;
//   * each lane sorts NL keys (what one list holds beyond NL is loaded by the partner).  The LOW lane works on the
//     bit-complemented keys, the high lane on the keys themselves, both ascending: one exchange v = max(v, ~partner's v)
//     then leaves the NL smallest of the 2 NL keys in the low lane and the NL largest in the high lane, each as a
//     "down then up" sequence in the lane's own domain -- which stays bitonic when padded with +inf to a power of two, so
//     a pruned bitonic merge finishes the sort (a padded "up then down" sequence would not be bitonic);
//   * ranks [HOUT, NL) move to the high lane with one shuffle each, so both lanes rebuild HOUT exact (d2, index) pairs;
//   * rows leave through the same shared-memory transpose.
// Exactness argument, fall-back lists and radius feedback are those of the tile kernel.
template <int NL_, int HOUT_, int NWARPS_, int CMAX_>
struct PairCfg {
    static constexpr int NL = NL_;                 // keys a lane sorts: a query absorbs 2 NL survivors
    static constexpr int HOUT = HOUT_;             // ranks a lane rebuilds: a row keeps KOUT = 2 HOUT >= k ranks
    static constexpr int KOUT = 2 * HOUT;
    static constexpr int NX = NL - HOUT;           // ranks [HOUT, NL) move from the low lane to the high lane
    static constexpr int LMAX = NL + NL / 3;       // entries ONE list may hold (the partner sorts what exceeds NL)
    static constexpr int SINK = 11;
    static constexpr int LCAP = LMAX + 1 + SINK;
    static constexpr int STRIDE = 33;
    static constexpr int WARPS = NWARPS_;
    static constexpr int CMAX = CMAX_;             // candidates staged per pass
    static constexpr int LIST_BYTES = LCAP * STRIDE * 4;
    static constexpr int STAGE_BYTES = CMAX * 16;
    static constexpr int BAR_BYTES = 16 + 3 * 128;
    static constexpr int SMEM_WARP_BYTES = STAGE_BYTES + LIST_BYTES + BAR_BYTES;
    static constexpr int MAX_PASSES = 12;
    static constexpr int RETRY_MIN = 8;            // queries (half of the warp's 16) that must fail the same way before a group retry
    static constexpr uint32_t SLOT_BITS = 10;
    static constexpr uint32_t SLOT_MASK = (1u << SLOT_BITS) - 1;
    static constexpr uint32_t PAD = ~SLOT_MASK;
    static constexpr int VN = NL <= 32 ? 32 : (NL <= 64 ? 64 : 128);   // virtual (power of two) length of the networks
    static_assert(CMAX % 16 == 0 && CMAX <= (1 << SLOT_BITS), "slot field too small");
    static_assert(LIST_BYTES % 16 == 0 && STAGE_BYTES % 16 == 0, "alignment");
    static_assert(HOUT * STRIDE * 4 <= LIST_BYTES && HOUT * STRIDE * 4 <= STAGE_BYTES, "output planes must fit");
    static_assert((1 << SLOT_BITS) * 16 <= SMEM_WARP_BYTES, "a padded slot must stay inside the warp's shared memory");
    static_assert(NX >= 0 && NX <= HOUT && 2 * HOUT - NL >= 0 && HOUT % 32 == 0 && (NX == 0 || HOUT <= 2 * NX), "rank split");
};

template <int NL, int HOUT, int NWARPS, int CMAX, int MINB>
__global__ void __launch_bounds__(NWARPS * 32, MINB) knn_pair_kernel(const GridView g, const SearchArgs a)
{
    using Cfg = PairCfg<NL, HOUT, NWARPS, CMAX>;
    constexpr int S = Cfg::STRIDE, KOUT = Cfg::KOUT, NX = Cfg::NX, MR = KOUT / 32;
    extern __shared__ __align__(128) unsigned char smem_tile[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ql = lane & 15;                      // query of this lane within the warp
    const uint32_t hi = (uint32_t)lane >> 4;       // 0: low lane (ranks [0, HOUT)), 1: high lane (ranks [HOUT, KOUT))
    unsigned char* wsm = smem_tile + (size_t)warp * Cfg::SMEM_WARP_BYTES;
    float4* stage = reinterpret_cast<float4*>(wsm);
    uint32_t* list = reinterpret_cast<uint32_t*>(wsm + Cfg::STAGE_BYTES);
    uint64_t* bar = reinterpret_cast<uint64_t*>(wsm + Cfg::STAGE_BYTES + Cfg::LIST_BYTES);
    const uint32_t base = (blockIdx.x * Cfg::WARPS + warp) * 16u;
    if (base >= a.n_query) return;
    const uint32_t k = a.k;
    if (lane == 0) { ptx::mbarrier_init(bar, 1); ptx::fence_mbarrier_init(); }
    __syncwarp();
    uint32_t parity = 0;

    const bool valid = base + ql < a.n_query;
    const float4 q4 = valid ? __ldg(a.queries + base + ql) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float qx = q4.x, qy = q4.y, qz = q4.z;
    const uint32_t row = __float_as_uint(q4.w);
    const int cqx = cell_coord(qx, g.lo[0], g.inv_hx, g.n[0]);
    const int cqy = cell_coord(qy, g.lo[1], g.inv_h, g.n[1]);
    const int cqz = cell_coord(qz, g.lo[2], g.inv_h, g.n[2]);
    const uint32_t rowid = (uint32_t)cqz * (uint32_t)g.n[1] + (uint32_t)cqy;
    const uint32_t cm = hi ? 0u : 0xffffffffu;     // the low lane sorts complemented keys

    // every mask below is symmetric: bit q and bit q + 16 describe the same query
    unsigned remaining = __ballot_sync(kFull, valid);
    unsigned slow = 0, unsafe = 0;
    float* const s_hint = reinterpret_cast<float*>(bar + 2) + lane;
    float* const s_rmul = s_hint + 32;
    uint32_t* const s_retry = reinterpret_cast<uint32_t*>(s_hint + 64);
    *s_hint = 0.f; *s_rmul = 1.f; *s_retry = 0u;
    for (int pass = 0; remaining; ++pass) {
        const int leader = __ffs(remaining) - 1;
        const uint32_t lrow = __shfl_sync(kFull, rowid, leader);
        const uint32_t cls = *s_retry >> 2;
        const uint32_t lcls = __shfl_sync(kFull, cls, leader);
        unsigned active = __ballot_sync(kFull, ((remaining >> lane) & 1u) && rowid == lrow && cls == lcls);
        remaining &= ~active;
        if (pass >= Cfg::MAX_PASSES) {
            slow |= active;
            if (a.stats && lane == 0) atomicAdd(a.stats + ST_REGION, (unsigned long long)__popc(active & 0xffffu));
            continue;
        }

        // ---- region of the active queries; a region too large to stage is halved along x ---------
        float R = 0.f;
        uint32_t s = 0, len = 0, off = 0, C = 0, C16 = 0;
        bool mine = false;
        for (int split = 0;; ++split) {
            mine = (active >> lane) & 1u;
            const float xmin = o2f(__reduce_min_sync(kFull, mine ? f2o(qx) : 0xffffffffu));
            const float xmax = o2f(__reduce_max_sync(kFull, mine ? f2o(qx) : 0u));
            const float ymin = o2f(__reduce_min_sync(kFull, mine ? f2o(qy) : 0xffffffffu));
            const float ymax = o2f(__reduce_max_sync(kFull, mine ? f2o(qy) : 0u));
            const float zmin = o2f(__reduce_min_sync(kFull, mine ? f2o(qz) : 0xffffffffu));
            const float zmax = o2f(__reduce_max_sync(kFull, mine ? f2o(qz) : 0u));
            const int cxa = __reduce_min_sync(kFull, mine ? cqx : 0x7fffffff);
            const int cxb = __reduce_max_sync(kFull, mine ? cqx : -1);
            const int cy = __shfl_sync(kFull, cqy, leader), cz = __shfl_sync(kFull, cqz, leader);
            {   // density and shape of the 3 x 3 block of (y, z) rows -> search radius (see knn_tile_kernel)
                const int bx0 = max(cxa - g.xf, 0), bx1 = min(cxb + g.xf, g.n[0] - 1);
                float c = 0.f;
                if (lane < 9) {
                    const int by = cy + lane % 3 - 1, bz = cz + lane / 3 - 1;
                    if (by >= 0 && by < g.n[1] && bz >= 0 && bz < g.n[2]) {
                        const uint32_t rb = ((uint32_t)bz * (uint32_t)g.n[1] + (uint32_t)by) * (uint32_t)g.n[0];
                        c = (float)(__ldg(g.cell_start + rb + bx1 + 1) - __ldg(g.cell_start + rb + bx0));
                    }
                }
                float s1 = c, s2 = c * c;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) { s1 += __shfl_xor_sync(kFull, s1, o); s2 += __shfl_xor_sync(kFull, s2, o); }
                s1 = fmaxf(__shfl_sync(kFull, s1, 0), 1.f);
                s2 = fmaxf(__shfl_sync(kFull, s2, 0), 1.f);
                const float m = fminf(fmaxf(s1 * s1 / s2, 1.f), 9.f);
                const float d = 1.f + log2f(m) * 0.6309298f;
                const float rho = s1 / (m * (float)(bx1 - bx0 + 1) * g.hx * g.h * g.h);
                const float T = a.target;
                const float l3 = log2f(T / (4.18879f * rho)) * (1.f / 3.f);
                const float l2 = log2f(T / (3.14159265f * rho * g.h)) * 0.5f;
                const float l1 = log2f(T / (2.f * rho * g.h * g.h));
                float lr = d >= 2.f ? (d - 2.f) * l3 + (3.f - d) * l2 : (d - 1.f) * l2 + (2.f - d) * l1;
                if (a.flags & 1u) lr = log2f(T / (4.18879f * s1 / (9.f * (float)(bx1 - bx0 + 1) * g.hx * g.h * g.h))) * (1.f / 3.f);
                const float rm = __uint_as_float(__reduce_max_sync(kFull, mine ? __float_as_uint(*s_rmul) : 0u));
                R = exp2f(lr) * rm;
            }
            if (R > g.rmax_safe) { unsafe |= active; active = 0; break; }
            const int cx0 = cell_coord(__fsub_rd(xmin, R), g.lo[0], g.inv_hx, g.n[0]);
            const int cx1 = cell_coord(__fadd_ru(xmax, R), g.lo[0], g.inv_hx, g.n[0]);
            const int cy0 = cell_coord(__fsub_rd(ymin, R), g.lo[1], g.inv_h, g.n[1]);
            const int cy1 = cell_coord(__fadd_ru(ymax, R), g.lo[1], g.inv_h, g.n[1]);
            const int cz0 = cell_coord(__fsub_rd(zmin, R), g.lo[2], g.inv_h, g.n[2]);
            const int cz1 = cell_coord(__fadd_ru(zmax, R), g.lo[2], g.inv_h, g.n[2]);
            const uint32_t nyr = (uint32_t)(cy1 - cy0 + 1);
            const uint32_t nrows = nyr * (uint32_t)(cz1 - cz0 + 1);
            bool fits = nrows <= 32u;
            if (fits) {
                s = 0; len = 0;
                if ((uint32_t)lane < nrows) {
                    const int rz = cz0 + (int)((uint32_t)lane / nyr), ry = cy0 + (int)((uint32_t)lane % nyr);
                    const uint32_t rb = ((uint32_t)rz * (uint32_t)g.n[1] + (uint32_t)ry) * (uint32_t)g.n[0];
                    s = __ldg(g.cell_start + rb + cx0);
                    len = __ldg(g.cell_start + rb + cx1 + 1) - s;
                }
                off = len;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(kFull, off, o);
                    if (lane >= o) off += t;
                }
                C = __shfl_sync(kFull, off, 31);
                off -= len;
                C16 = (C + 15u) & ~15u;
                fits = C16 <= (uint32_t)Cfg::CMAX;
            }
            if (fits) break;
            const int na = __popc(active & 0xffffu);
            if (na < 2 || split >= 3 || nrows > 32u) {
                slow |= active;
                if (a.stats && lane == 0) atomicAdd(a.stats + ST_REGION, (unsigned long long)na);
                active = 0;
                break;
            }
            // keep the lower half of the queries (cell-sorted: the lower half along x); the rest waits for a later pass
            const unsigned keep = __ballot_sync(kFull, mine && __popc(active & ((1u << ql) - 1u)) < na / 2);
            remaining |= active & ~keep;
            active = keep;
        }
        if (!active) continue;

        // ---- stage the region: one 1-D TMA bulk copy per row ------------------------------------
        __syncwarp();
        if (lane == 0) { ptx::fence_proxy_async_smem(); ptx::mbarrier_arrive_expect_tx(bar, C * 16u); }
        __syncwarp();
        if (len) ptx::bulk_g2s(stage + off, g.pts + s, len * 16u, bar);
        if (C + (uint32_t)lane < C16) stage[C + lane] = make_float4(__int_as_float(0x7f800000), 0.f, 0.f, 0.f);
        ptx::mbarrier_wait(bar, parity);
        parity ^= 1u;
        __syncwarp();

        // ---- scan: the low lanes test the even staged slots, the high lanes the odd ones ------------
        const float t2 = !mine ? -1.f : __fmul_rd(__fmul_rd(R, R), 0.9999f);
        const float t2_safe = __fmul_rd(t2, 0.99999618530273f);
        uint32_t* const wbase = list + lane;
        const uint32_t waddr0 = ptx::smem_addr(wbase);
        const uint32_t wclamp = waddr0 + (Cfg::LMAX + 1) * S * 4;
        uint32_t waddr = waddr0;
#define PGEOF_PAIR_APPEND(D2, SLOT)                                                           \
        asm volatile("{\n\t.reg .pred p;\n\t"                                                 \
                     "setp.le.f32 p, %1, %2;\n\t"                                             \
                     "@p st.shared.u32 [%0], %3;\n\t"                                         \
                     "@p add.u32 %0, %0, %4;\n\t}"                                            \
                     : "+r"(waddr)                                                            \
                     : "f"(D2), "f"(t2), "r"((__float_as_uint(D2) & ~Cfg::SLOT_MASK) | (SLOT)), "n"(S * 4));
        for (uint32_t c = hi; c < C16; c += 16) {
            float d[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { const float4 p = stage[c + 2 * i]; d[i] = sqdist_fused(qx, qy, qz, p.x, p.y, p.z); }
#pragma unroll
            for (int i = 0; i < 8; ++i) PGEOF_PAIR_APPEND(d[i], c + 2 * i)
            waddr = min(waddr, wclamp);
        }
#undef PGEOF_PAIR_APPEND
        __syncwarp();
        const uint32_t cnt_own = (waddr - waddr0) / (S * 4);
        const uint32_t cnt_par = __shfl_xor_sync(kFull, cnt_own, 16);
        const uint32_t cnt = cnt_own + cnt_par;                                   // survivors of the query
        const bool over = cnt_own > (uint32_t)Cfg::LMAX || cnt_par > (uint32_t)Cfg::LMAX || cnt > 2u * NL;
        bool ok = mine && cnt >= k && !over;
        if (mine) *s_hint = R * (cnt < k ? fminf(cbrtf(1.25f * a.target / fmaxf((float)cnt, 2.f)), 3.f) : (over ? 0.88f : 1.f));
        unsigned requeued = 0;
        if (pass + 2 < Cfg::MAX_PASSES && !(a.flags & 2u)) {
            const uint32_t retry = *s_retry;
            const bool can = mine && (retry & 3u) < 2u;
            const unsigned sh = __ballot_sync(kFull, can && cnt < k), ov = __ballot_sync(kFull, can && over);
            if (__popc(sh & 0xffffu) >= Cfg::RETRY_MIN) {
                if ((sh >> lane) & 1u) {
                    *s_rmul *= fminf(fmaxf(exp2f(0.4f * log2f(1.3f * a.target / fmaxf((float)cnt, 1.f))), 1.15f), 3.f);
                    *s_retry = ((retry & 3u) + 1u) | (1u << 2);
                }
                requeued |= sh;
            }
            if (__popc(ov & 0xffffu) >= Cfg::RETRY_MIN) {
                if ((ov >> lane) & 1u) { *s_rmul *= 0.7f; *s_retry = ((retry & 3u) + 1u) | (2u << 2); }
                requeued |= ov;
            }
            remaining |= requeued;
        }
        if (a.stats) {
            const unsigned sh = __ballot_sync(kFull, mine && cnt < k), ov = __ballot_sync(kFull, mine && over);
            const uint32_t sv = __reduce_add_sync(kFull, mine ? cnt_own : 0u);
            if (lane == 0) {
                atomicAdd(a.stats + ST_SHORT, (unsigned long long)__popc(sh & 0xffffu));
                atomicAdd(a.stats + ST_OVER, (unsigned long long)__popc(ov & 0xffffu));
                atomicAdd(a.stats + ST_PASSES, 1ull);
                atomicAdd(a.stats + ST_CANDS, (unsigned long long)C);
                atomicAdd(a.stats + ST_SURV, (unsigned long long)sv);
            }
        }

        // ---- sort: NL keys per lane in registers -------------------------------------------------
        uint32_t dropped;                      // smallest key outside the KOUT kept ranks (valid in both lanes after the shuffle)
        uint32_t bad[HOUT / 32] = {};          // bit j: this lane's step j precedes its step j - 1 in the exact order
        uint32_t fd = 0, fi = 0, ld = 0, li = 0;   // exact pair of this lane's first / last step
        {
            uint32_t v[NL];
            // register i <- own entry i, or (own list exhausted) the partner's entry NL + (i - cnt_own)
            const uint32_t own_a = waddr0;
            const uint32_t par_a = ptx::smem_addr(list + (lane ^ 16)) + (uint32_t)((NL - (int)min(cnt_own, (uint32_t)NL)) * S * 4);
            const uint32_t lim = min(cnt_own, (uint32_t)NL) + (cnt_par > (uint32_t)NL ? min(cnt_par, (uint32_t)Cfg::LMAX + 1u) - NL : 0u);
#pragma unroll
            for (int i = 0; i < NL; ++i) {
                const uint32_t ad = ((uint32_t)i < cnt_own ? own_a : par_a) + (uint32_t)(i * S * 4);
                uint32_t w;                    // predicated: an entry beyond `lim` may lie outside the list
                asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %2, %3;\n\tmov.u32 %0, %4;\n\t@p ld.shared.u32 %0, [%1];\n\t}"
                             : "=r"(w) : "r"(ad), "r"((uint32_t)i), "r"(lim), "r"(Cfg::PAD));
                v[i] = w ^ cm;
            }
            __syncwarp();                      // the lists are dead: their memory becomes the d2 plane
            RegOddEvenSort<NL, 0, Cfg::VN>::run(v);
            // low lane: the NL smallest of the pair's keys (complemented), high lane: the NL largest; both "down then up"
#pragma unroll
            for (int i = 0; i < NL; ++i) v[i] = max(v[i], ~__shfl_xor_sync(kFull, v[i], 16));
            RegBitonicMerge<NL, 0, Cfg::VN>::run(v);
            // low lane: v[i] = ~key of rank NL - 1 - i; high lane: v[i] = key of rank NL + i
            if constexpr (KOUT < 2 * NL) dropped = __shfl_sync(kFull, v[KOUT - NL], ql + 16);      // rank KOUT lives in the high lane
            else dropped = 0xffffffffu;
            // ranks [HOUT, NL) move to the high lane; afterwards step j of either lane reads v[NL - 1 - j]
            if (NX > 0) {
#pragma unroll
                for (int j = NX; j < HOUT; ++j) v[NL - 1 - j] = hi ? v[j - NX] : v[NL - 1 - j];
#pragma unroll
                for (int j = 0; j < NX; ++j) {          // low lane's rank HOUT + j (sources [0, NX) and targets [HOUT, NL) are disjoint)
                    const uint32_t t = ~__shfl_xor_sync(kFull, v[NX - 1 - j], 16);
                    v[NL - 1 - j] = hi ? t : v[NL - 1 - j];
                }
            } else {
#pragma unroll
                for (int j = 0; j < HOUT / 2; ++j) {   // NL == HOUT: the high lane only reverses its registers
                    const uint32_t x = v[j], y = v[NL - 1 - j];
                    v[j] = hi ? y : x; v[NL - 1 - j] = hi ? x : y;
                }
            }
            // rebuild the exact (d2, index) pairs: step j of this lane is rank hi * HOUT + j, plane[j * S + lane]
            uint32_t* plane_d = list + lane;
            uint32_t pd = 0, pi = 0;
            const uint32_t rank0 = hi * HOUT;
#pragma unroll
            for (int j = 0; j < HOUT; ++j) {
                const float4 p = stage[(v[NL - 1 - j] ^ cm) & Cfg::SLOT_MASK];
                const uint32_t d = __float_as_uint(sqdist_f32(qx, qy, qz, p.x, p.y, p.z));
                const uint32_t id = __float_as_uint(p.w);
                if (j > 0) bad[j / 32] |= (rank0 + (uint32_t)j < cnt && (d < pd || (d == pd && id < pi))) ? (1u << (j % 32)) : 0u;
                else { fd = d; fi = id; }
                pd = d; pi = id;
                plane_d[j * S] = d;
                v[NL - 1 - j] = id;
            }
            ld = pd; li = pi;
            __syncwarp();                      // the staged candidates are dead: index plane
            uint32_t* plane_i = reinterpret_cast<uint32_t*>(stage) + lane;
#pragma unroll
            for (int j = 0; j < HOUT; ++j) plane_i[j * S] = v[NL - 1 - j];
        }
        {
            // the high lane's first step follows the low lane's last one
            const uint32_t qd = __shfl_xor_sync(kFull, ld, 16), qi = __shfl_xor_sync(kFull, li, 16);
            if (hi && (uint32_t)HOUT < cnt && (fd < qd || (fd == qd && fi < qi))) bad[0] |= 1u;
            uint32_t any_bad = 0;
#pragma unroll
            for (int r = 0; r < HOUT / 32; ++r) any_bad |= bad[r];
            const uint32_t pair_bad = any_bad | __shfl_xor_sync(kFull, any_bad, 16);
            __syncwarp();
            uint32_t* const pd_q = list + ql;                                  // rank e of this query: [(e % HOUT) * S + (e / HOUT) * 16]
            uint32_t* const pi_q = reinterpret_cast<uint32_t*>(stage) + ql;
#define PGEOF_RANK_AT(e) (((e) % (uint32_t)HOUT) * S + ((e) / (uint32_t)HOUT) * 16u)
            // the pair's inversion flags, by rank
            uint32_t comb[MR];
#pragma unroll
            for (int r = 0; r < HOUT / 32; ++r) {
                const uint32_t o = __shfl_xor_sync(kFull, bad[r], 16);
                comb[r] = hi ? o : bad[r];
                comb[r + HOUT / 32] = hi ? bad[r] : o;
            }
            if (ok && pair_bad && !hi) {
                // insertion sort of the row on the exact (d2, index) order from the first inversion on (rare: truncated keys
                // that collided); past the last inversion the first entry found in place ends it
                uint32_t first = 0, last = 0;
#pragma unroll
                for (int r = MR - 1; r >= 0; --r) if (comb[r]) first = r * 32 + __ffs(comb[r]) - 1;
#pragma unroll
                for (int r = 0; r < MR; ++r) if (comb[r]) last = r * 32 + 31 - __clz(comb[r]);
                const uint32_t nfix = min(cnt, (uint32_t)KOUT);
                for (uint32_t i = first; i < nfix; ++i) {
                    const uint32_t d = pd_q[PGEOF_RANK_AT(i)], id = pi_q[PGEOF_RANK_AT(i)];
                    uint32_t j = i;
                    while (j > 0) {
                        const uint32_t qd2 = pd_q[PGEOF_RANK_AT(j - 1)], qi2 = pi_q[PGEOF_RANK_AT(j - 1)];
                        if (qd2 < d || (qd2 == d && qi2 < id)) break;
                        pd_q[PGEOF_RANK_AT(j)] = qd2; pi_q[PGEOF_RANK_AT(j)] = qi2;
                        --j;
                    }
                    if (j != i) { pd_q[PGEOF_RANK_AT(j)] = d; pi_q[PGEOF_RANK_AT(j)] = id; }
                    else if (i > last) break;
                }
            }
            __syncwarp();
            // exact iff (a) no dropped key can precede the k-th neighbour and (b) no rejected candidate can (see knn_tile_kernel)
            const uint32_t kth = pd_q[PGEOF_RANK_AT(k - 1)];
#undef PGEOF_RANK_AT
            const bool tie = ok && ((dropped < Cfg::PAD && (dropped >> Cfg::SLOT_BITS) < (kth >> Cfg::SLOT_BITS) + 2u) ||
                                    !(__uint_as_float(kth) <= t2_safe));
            if (tie) ok = false;
            if (a.stats) {
                const unsigned tm = __ballot_sync(kFull, tie), fm = __ballot_sync(kFull, ok && pair_bad);
                if (lane == 0) {
                    atomicAdd(a.stats + ST_TIE, (unsigned long long)__popc(tm & 0xffffu));
                    atomicAdd(a.stats + ST_FIXED, (unsigned long long)__popc(fm & 0xffffu));
                }
            }
        }
        slow |= __ballot_sync(kFull, mine && !ok) & ~requeued;
        __syncwarp();
        // ---- write the finished rows, one coalesced row at a time ---------------------------------
        {
            const uint32_t* plane_d = list;
            const uint32_t* plane_i = reinterpret_cast<const uint32_t*>(stage);
            const unsigned okm = __ballot_sync(kFull, ok);
#pragma unroll 8
            for (int q = 0; q < 16; ++q) {
                const uint32_t rowq = __shfl_sync(kFull, row, q);
                const size_t o = (size_t)rowq * k + lane;
                uint32_t* idx = reinterpret_cast<uint32_t*>(a.indices) + o;
                float* d2 = a.sqr_dist + o;
#pragma unroll
                for (int r = 0; r < MR; ++r) {
                    const uint32_t e = r * 32 + lane;
                    const uint32_t at = (e % (uint32_t)HOUT) * S + (e / (uint32_t)HOUT) * 16u + (uint32_t)q;
                    const uint32_t vi = plane_i[at], vd = plane_d[at];
                    const uint32_t on = (((okm >> q) & 1u) && e < k) ? 1u : 0u;
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %4, 0;\n\t"
                                 "@p st.global.u32 [%0], %2;\n\t@p st.global.u32 [%1], %3;\n\t}"
                                 :: "l"(idx + r * 32), "l"(d2 + r * 32), "r"(vi), "r"(vd), "r"(on) : "memory");
                }
            }
        }
        __syncwarp();
    }

    if (slow & 0xffffu) {
        const unsigned sl = slow & 0xffffu;
        uint32_t at = 0;
        if (lane == 0) at = atomicAdd(a.slow_count, (uint32_t)__popc(sl));
        at = __shfl_sync(kFull, at, 0);
        if ((sl >> lane) & 1u) a.slow_list[at + __popc(sl & lanemask_lt())] = make_uint2(base + lane, __float_as_uint(*s_hint));
    }
    if (unsafe & 0xffffu) {
        const unsigned us = unsafe & 0xffffu;
        uint32_t at = 0;
        if (lane == 0) at = atomicAdd(a.unsafe_count, (uint32_t)__popc(us));
        at = __shfl_sync(kFull, at, 0);
        if ((us >> lane) & 1u) a.unsafe_list[at + __popc(us & lanemask_lt())] = make_uint2(base + lane, 0u);
    }
}

// ---------------------------------------------------------------------------------------
// any k: one CTA per query, keys in GLOBAL scratch (knn / max_knn beyond the 512 the register-resident select + sort of
// search_kernel covers; the reference accepts any knn <= len(data), nn_search.hpp:37,92).  Same exactness argument
// as the warp routine: the ball is grown until it holds k points, the 64-bit key threshold is bisected while more than
// `cap` keys pass it, the survivors are sorted by a block-wide bitonic network over global memory.  Built for
// completeness, not speed: every query costs O(cap log^2 cap) global-memory compare-exchanges.
// ---------------------------------------------------------------------------------------
constexpr int kBigThreads = 256;

// keys <= tau of ball(q, R): returns the count (block uniform); the first `cap` are appended to buf when STORE
template <bool STORE>
__device__ __forceinline__ uint32_t cta_scan_ball(const GridView& g, float qx, float qy, float qz, float R, u64 tau,
                                                  u64* __restrict__ buf, uint32_t cap, uint32_t* s_count)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) *s_count = 0;
    __syncthreads();
    const int cx0 = cell_coord(__fsub_rd(qx, R), g.lo[0], g.inv_hx, g.n[0]);
    const int cx1 = cell_coord(__fadd_ru(qx, R), g.lo[0], g.inv_hx, g.n[0]);
    const int cy0 = cell_coord(__fsub_rd(qy, R), g.lo[1], g.inv_h, g.n[1]);
    const int cy1 = cell_coord(__fadd_ru(qy, R), g.lo[1], g.inv_h, g.n[1]);
    const int cz0 = cell_coord(__fsub_rd(qz, R), g.lo[2], g.inv_h, g.n[2]);
    const int cz1 = cell_coord(__fadd_ru(qz, R), g.lo[2], g.inv_h, g.n[2]);
    const int cqy = cell_coord(qy, g.lo[1], g.inv_h, g.n[1]);
    const int cqz = cell_coord(qz, g.lo[2], g.inv_h, g.n[2]);
    const uint32_t nyr = (uint32_t)(cy1 - cy0 + 1);
    const uint32_t nrows = nyr * (uint32_t)(cz1 - cz0 + 1);
    const float R2u = __fmul_ru(R, R);
    const unsigned lt = lanemask_lt();
    for (uint32_t r = warp; r < nrows; r += kBigThreads / 32) {       // one (y, z) cell row per warp
        const int cz = cz0 + (int)(r / nyr), cy = cy0 + (int)(r % nyr);
        const float gy = axis_gap(qy, cy, cqy, g.lo[1], g.h, g.slack);
        const float gz = axis_gap(qz, cz, cqz, g.lo[2], g.h, g.slack);
        const float rem = __fsub_ru(__fsub_ru(R2u, __fmul_rd(gy, gy)), __fmul_rd(gz, gz));
        if (!(rem >= 0.f)) continue;
        const float xr = __fsqrt_ru(rem);
        const int x0 = max(cx0, cell_coord(__fsub_rd(qx, xr), g.lo[0], g.inv_hx, g.n[0]));
        const int x1 = min(cx1, cell_coord(__fadd_ru(qx, xr), g.lo[0], g.inv_hx, g.n[0]));
        const uint32_t row = ((uint32_t)cz * (uint32_t)g.n[1] + (uint32_t)cy) * (uint32_t)g.n[0];
        const uint32_t s = __ldg(g.cell_start + row + x0), e = __ldg(g.cell_start + row + x1 + 1);
        for (uint32_t base = s; base < e; base += 32) {
            const uint32_t j = base + lane;
            bool acc = false;
            u64 key = 0;
            if (j < e) {
                const float4 p = __ldg(g.pts + j);
                key = make_key(sqdist_f32(qx, qy, qz, p.x, p.y, p.z), __float_as_uint(p.w));
                acc = key <= tau;
            }
            const unsigned m = __ballot_sync(kFull, acc);
            if (m) {
                uint32_t at = 0;
                if (lane == 0) at = atomicAdd(s_count, (uint32_t)__popc(m));
                at = __shfl_sync(kFull, at, 0);
                if (STORE && acc) {
                    const uint32_t pos = at + __popc(m & lt);
                    if (pos < cap) buf[pos] = key;
                }
            }
        }
    }
    __syncthreads();
    const uint32_t c = *s_count;
    __syncthreads();
    return c;
}

// ascending bitonic sort of buf[0, n), n a power of two, by the whole block
__device__ __forceinline__ void cta_bitonic_sort(u64* __restrict__ buf, uint32_t n)
{
    for (uint32_t size = 2; size <= n; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = threadIdx.x; t < n / 2; t += kBigThreads) {
                const uint32_t lo = 2 * t - (t & (stride - 1));
                const uint32_t hi = lo + stride;
                const bool asc = (lo & size) == 0;
                const u64 a = buf[lo], b = buf[hi];
                if ((a > b) == asc) { buf[lo] = b; buf[hi] = a; }
            }
            __syncthreads();
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(kBigThreads) search_big_kernel(const GridView g, const SearchArgs a, u64* __restrict__ scratch, uint32_t cap)
{
    __shared__ uint32_t s_count;
    u64* buf = scratch + (size_t)blockIdx.x * cap;
    const uint32_t k = a.k;
    for (uint32_t w = blockIdx.x; w < a.n_query; w += gridDim.x) {
        const float4 q4 = __ldg(a.queries + w);
        const float qx = q4.x, qy = q4.y, qz = q4.z;
        const uint32_t row = __float_as_uint(q4.w);
        uint32_t c = 0;          // keys <= tau, stored in buf (c <= cap when the loops end)
        uint32_t need = k;
        u64 tau = 0;
        if (MODE == SEARCH_KNN) {
            // ball seeded from the global density, grown until it holds k points; then the threshold is bisected while it overflows
            float R = fmaxf(a.target, 1e-30f) + bbox_distance(g, qx, qy, qz);
            u64 tau_lo = 0, tau_hi = 0;
            bool have_hi = false;
            float Rg = R;
            tau = tau_from_radius(R);
            for (int it = 0; it < 600; ++it) {
                c = cta_scan_ball<true>(g, qx, qy, qz, Rg, tau, buf, cap, &s_count);
                if (c < k) {
                    tau_lo = tau;
                    if (have_hi) tau = tau_lo + (tau_hi - tau_lo) / 2;
                    else { R *= 1.35f; Rg = R; tau = tau_from_radius(R); }
                } else if (c > cap) {
                    tau_hi = tau; have_hi = true;
                    tau = tau_lo + (tau_hi - tau_lo) / 2;                  // Rg stays: it covers every key <= tau_hi
                } else break;
            }
        } else {
            const float r2 = __fmul_rn(a.radius, a.radius);                   // nn_search.hpp:98
            if (r2 > 0.f) {
                tau = ((u64)__float_as_uint(r2) << 32) - 1;                   // d2 < r2, strict
                const float Rg = __fmul_ru(__fsqrt_ru(r2), 1.0001f);
                c = cta_scan_ball<true>(g, qx, qy, qz, Rg, tau, buf, cap, &s_count);
                need = min(c, k);
                if (c > cap) {                                                // more than cap (>= 2k) inside: keep between k and cap of the nearest
                    u64 tau_lo = 0, tau_hi = tau;
                    for (int it = 0; it < 600; ++it) {
                        tau = tau_lo + (tau_hi - tau_lo) / 2;
                        c = cta_scan_ball<true>(g, qx, qy, qz, Rg, tau, buf, cap, &s_count);
                        if (c < k) tau_lo = tau; else if (c > cap) tau_hi = tau; else break;
                    }
                }
            } else need = 0;
        }
        // sort what was kept (padded to a power of two)
        uint32_t n = 1;
        while (n < c) n <<= 1;
        for (uint32_t t = c + threadIdx.x; t < n; t += kBigThreads) buf[t] = kKeyMax;
        __syncthreads();
        cta_bitonic_sort(buf, n);
        if (MODE == SEARCH_KNN) {
            uint32_t* idx = reinterpret_cast<uint32_t*>(a.indices) + (size_t)row * k;
            float* d2 = a.sqr_dist + (size_t)row * k;
            for (uint32_t e = threadIdx.x; e < k; e += kBigThreads) { idx[e] = key_idx(buf[e]); d2[e] = key_d2(buf[e]); }
        } else if (MODE == SEARCH_RADIUS && a.nn_ptr) {                    // single-search CSR pair: the count + the hits
            int32_t* idx = reinterpret_cast<int32_t*>(a.indices) + (size_t)row * k;
            if (threadIdx.x == 0) a.nn_ptr[row] = need;
            for (uint32_t e = threadIdx.x; e < need; e += kBigThreads) idx[e] = (int32_t)key_idx(buf[e]);
        } else if (MODE == SEARCH_RADIUS) {
            int32_t* idx = reinterpret_cast<int32_t*>(a.indices) + (size_t)row * k;
            float* d2 = a.sqr_dist + (size_t)row * k;
            for (uint32_t e = threadIdx.x; e < k; e += kBigThreads) {      // pad: -1 / 0 (nn_search.hpp:104,108)
                const bool hit = e < need;
                idx[e] = hit ? (int32_t)key_idx(buf[e]) : -1;
                d2[e] = hit ? key_d2(buf[e]) : 0.f;
            }
        } else {   // SEARCH_RADIUS_CSR
            const uint32_t base = a.nn_ptr[row];
            uint32_t* nn = reinterpret_cast<uint32_t*>(a.indices);
            for (uint32_t e = threadIdx.x; e < need; e += kBigThreads) {
                nn[(size_t)base + e] = key_idx(buf[e]);
                if (a.sqr_dist) a.sqr_dist[(size_t)base + e] = key_d2(buf[e]);
            }
        }
        __syncthreads();
    }
}

template <int MODE>
int launch_big(const GridView& g, SearchArgs a, size_t n_data, cudaStream_t stream)
{
    // capacity: a power of two >= 2k (room for the ball to overshoot), at least 2048
    uint64_t cap = 2048;
    while (cap < 2ull * a.k) cap <<= 1;
    if (cap > 0x80000000ull) { set_error("knn / max_knn = %u is too large", a.k); return PGEOF_EINVAL; }
    const int sms = sm_count();
    const uint64_t budget = 1ull << 30;                                     // bytes of key scratch
    const uint64_t blocks = std::max<uint64_t>(1, std::min<uint64_t>({(uint64_t)a.n_query, (uint64_t)sms * 8, budget / (cap * sizeof(u64))}));
    DeviceBuffer scratch;
    PGEOF_TRY(scratch.alloc((size_t)(blocks * cap * sizeof(u64)), stream));
    if (MODE == SEARCH_KNN) {
        // radius of a ball that holds ~1.2 k points at the mean density of the indexed box
        double vol = 1;
        int dims = 0;
        for (int d = 0; d < 3; ++d) { const double e = (double)g.n[d] * (d == 0 ? g.hx : g.h); if (e > 0) { vol *= e; ++dims; } }
        const double per = vol / (double)std::max<size_t>(n_data, 1);
        a.target = (float)std::cbrt(1.2 * a.k * per / 4.18879);
    }
    {
        KernelTimer timer(MODE == SEARCH_KNN ? "knn_search" : "radius_search", stream);
        search_big_kernel<MODE><<<(unsigned)blocks, kBigThreads, 0, stream>>>(g, a, scratch.as<u64>(), (uint32_t)cap);
    }
    PGEOF_LAUNCH_CHECK();
    return PGEOF_OK;
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

template <int NOUT, int NEXTRA, int NWARPS = 2, int MODE = SEARCH_KNN, bool FUSED = false, bool LOCK = true, bool ROLLED = false>
int launch_tile(const GridView& g, const SearchArgs& a, cudaStream_t stream, bool run_slow = true)
{
    using Cfg = TileCfg<NOUT, NEXTRA, NWARPS>;
    const size_t smem = (size_t)Cfg::WARPS * Cfg::SMEM_WARP_BYTES;
    auto kern = knn_tile_kernel<NOUT, NEXTRA, NWARPS, MODE, FUSED, LOCK, ROLLED>;
    PGEOF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned blocks = (a.n_query + Cfg::WARPS * 32 - 1) / (Cfg::WARPS * 32);
    {
        KernelTimer timer(MODE == SEARCH_KNN ? "knn_search" : "radius_search", stream);
        kern<<<blocks, Cfg::WARPS * 32, smem, stream>>>(g, a);
        if (MODE == SEARCH_KNN) { if (run_slow) knn_slow_kernel<NOUT><<<sm_count() * 4, kWarps * 32, 0, stream>>>(g, a); }
        else search_list_kernel<NOUT, SEARCH_RADIUS><<<sm_count() * 4, kWarps * 32, (size_t)kWarps * SearchCfg<NOUT, SEARCH_RADIUS>::CAP * sizeof(u64), stream>>>(g, a);
    }
    PGEOF_LAUNCH_CHECK();
    PGEOF_LAUNCH_CHECK();
    return PGEOF_OK;
}

template <int NL, int HOUT, int NWARPS, int CMAX, int MINB>
int launch_pair(const GridView& g, const SearchArgs& a, cudaStream_t stream)
{
    using Cfg = PairCfg<NL, HOUT, NWARPS, CMAX>;
    const size_t smem = (size_t)Cfg::WARPS * Cfg::SMEM_WARP_BYTES;
    auto kern = knn_pair_kernel<NL, HOUT, NWARPS, CMAX, MINB>;
    PGEOF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned blocks = (a.n_query + Cfg::WARPS * 16 - 1) / (Cfg::WARPS * 16);
    {
        KernelTimer timer("knn_search", stream);
        kern<<<blocks, Cfg::WARPS * 32, smem, stream>>>(g, a);
        knn_slow_kernel<Cfg::KOUT><<<sm_count() * 4, kWarps * 32, 0, stream>>>(g, a);
    }
    PGEOF_LAUNCH_CHECK();
    PGEOF_LAUNCH_CHECK();
    return PGEOF_OK;
}

// survivors the tile kernel's sorting network absorbs for a given k
float env_float(const char* name, float dflt);

// survivors the tile kernel's sorting network absorbs for a given k (a 128-key network for 52 < k <= 64 was
// measured slower than this one at every k: 6 instead of 8 resident warps, 1342 instead of 985 comparators)
inline int tile_nload(uint32_t k) { return k <= 32 ? 64 : (k <= 64 ? 96 : 160); }

template <int MODE>
int dispatch_search(uint32_t k, const GridView& g, const SearchArgs& a, size_t n_data, cudaStream_t stream)
{
    if (k <= 32) return launch_search<32, MODE>(g, a, stream);
    if (k <= 64) return launch_search<64, MODE>(g, a, stream);
    if (k <= 128) return launch_search<128, MODE>(g, a, stream);
    if (k <= 256) return launch_search<256, MODE>(g, a, stream);
    if (k <= 512) return launch_search<512, MODE>(g, a, stream);
    return launch_big<MODE>(g, a, n_data, stream);   // any knn <= len(data), as the reference (nn_search.hpp:37,92)
}

__global__ void iota_scale_u32(uint32_t* out, size_t n, uint32_t k)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)(i * k);
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
    // PGEOF_KNN_PAIR = 0 keeps the thread-per-query tile kernel for 32 < k <= 64 (and the warp-per-query routine above 64)
    const bool pair = mode == SEARCH_KNN && k > 32 && k <= 64 && env_float("PGEOF_KNN_PAIR", 0.f) != 0.f;
    // 64 < k <= 128: PGEOF_KNN_TILE128 = 1 runs the tile kernel with a 160-key network and the rolled rebuild (254 registers,
    // 42 KB of shared memory per warp -> 5 warps per SM).  Bit-exact, measured at 10 M x k = 100: 31.4 ms against 25.0 ms
    // for the warp-per-query routine (887 candidates per pass, 20 % of the rows need the tie repair) -> off by default.
    const bool tile128 = mode == SEARCH_KNN && k > 64 && k <= 128 && env_float("PGEOF_KNN_TILE128", 0.f) != 0.f;
    const bool tile = (mode == SEARCH_KNN ? env_float("PGEOF_KNN_TILE", 1.f) != 0.f : mode == SEARCH_RADIUS && env_float("PGEOF_RADIUS_TILE", 1.f) != 0.f) &&
                      (k <= 64 || tile128);
    // the reference only ever uses r * r (nn_search.hpp:98): a negative radius searches the ball of |r|
    if (mode != SEARCH_KNN && !std::isfinite(radius)) { set_error("search_radius must be finite"); return PGEOF_EINVAL; }
    radius = std::fabs(radius);
    float occ = 0.f, edge = 0.f;
    int xf = 1;
    if (mode == SEARCH_KNN) {
        // ball seeded to hold k + z sigma + 2 points.  Tile path: z balances the two ways a query leaves the
        // fast path (fewer than k survivors / more than the sorting network absorbs); cell edge h slightly
        // above that ball's radius so that a warp's region is 3 x 3 (y, z) rows, cells 8x finer along x.
        target = (float)k + (tile ? env_float("PGEOF_KNN_Z", 2.6f) : 2.f) * std::sqrt((float)k) + 2.f;
        if (tile) target = std::min(target, 0.5f * (float)(k + tile_nload(k)));
        occ = std::max(2.f, target * env_float("PGEOF_KNN_CELL_OCC", tile ? 0.28f : 0.25f));
        // (the warp-per-query routine also trims its x-spans at the finer granularity: 21.5 -> 18.9 ms at 10 M x k = 100)
        xf = tile ? (int)env_float("PGEOF_KNN_XF", 8.f) : std::max(1, (int)env_float("PGEOF_GENERIC_XF", 8.f));
    } else {
        // tile path: cell edge just above the radius, so that the ball of a query reaches one cell row to either side
        edge = radius * env_float("PGEOF_RADIUS_CELL_SCALE", tile ? 1.002f : 1.0f);
        if (!(edge > 0.f)) edge = 1.f;
        xf = tile ? (int)env_float("PGEOF_KNN_XF", 8.f) : std::max(1, (int)env_float("PGEOF_GENERIC_XF", 8.f));
    }
    // Queries that live in a small part of the cloud (a spatial shard of a multi-GPU run, a region of interest): index only
    // the points within `halo` of their bounding box.  Exact as long as no ball outgrows the halo; the kernels check
    // that (GridView::rmax_safe) and hand the few queries that do to a second run on the full grid.
    GridClip clip{};
    bool clipped = false;
    const bool sub_query = !(query == data && n_query == n_data);
    if (sub_query && n_query * 4 <= n_data * 3 && (mode != SEARCH_KNN || tile) && env_float("PGEOF_GRID_CLIP", 1.f) != 0.f) {
        float dlo[3], dhi[3], qlo[3], qhi[3];
        PGEOF_TRY(bbox_host(data, n_data, query, n_query, stream, dlo, dhi, qlo, qhi));
        double vol = 1, h_est = edge;
        int dims = 0;
        for (int d = 0; d < 3; ++d) if (dhi[d] > dlo[d]) { vol *= (double)dhi[d] - (double)dlo[d]; ++dims; }
        if (mode == SEARCH_KNN) h_est = dims ? std::pow(vol * occ / (double)n_data, 1.0 / dims) : 1.0;
        const float halo = mode == SEARCH_KNN ? 3.f * (float)h_est : 2.2f * radius + 1e-30f;
        double cvol = 1;
        bool any_cut = false;
        for (int d = 0; d < 3; ++d) {
            clip.lo[d] = std::max(dlo[d], std::nextafter(qlo[d] - halo, -INFINITY));
            clip.hi[d] = std::min(dhi[d], std::nextafter(qhi[d] + halo, INFINITY));
            if (clip.lo[d] > clip.hi[d]) { clip.lo[d] = dlo[d]; clip.hi[d] = dhi[d]; }   // queries outside the cloud on this axis
            any_cut = any_cut || clip.lo[d] > dlo[d] || clip.hi[d] < dhi[d];
            if (dhi[d] > dlo[d]) cvol *= std::max((double)clip.hi[d] - (double)clip.lo[d], 0.0) / ((double)dhi[d] - (double)dlo[d]);
        }
        bool inside = true;    // the exactness argument needs every query inside the undilated box: true by construction,
        for (int d = 0; d < 3; ++d) inside = inside && qlo[d] >= dlo[d] - halo && qhi[d] <= dhi[d] + halo;   // except far outside the cloud
        if (any_cut && inside && cvol < 0.6) {
            clip.cell_edge = (float)h_est;
            clip.rmax_safe = halo;
            clipped = true;
        }
    }
    PGEOF_TRY(grid_build(data, n_data, mode == SEARCH_KNN ? 0.f : edge, occ, xf, stream, &grid, clipped ? &clip : nullptr));
    DeviceBuffer qsorted;
    const float4* qrec;
    if (query == data && n_query == n_data) qrec = grid.view.pts;   // self query: reuse the sorted cloud
    else { PGEOF_TRY(grid_sort_queries(grid, query, n_query, stream, &qsorted)); qrec = qsorted.as<float4>(); }
    SearchArgs a{qrec, (uint32_t)n_query, k, radius, target, indices, sqr_dist, nn_ptr, nullptr, (uint32_t)env_float("PGEOF_KNN_FLAGS", 0.f),
                 nullptr, nullptr, nullptr, nullptr, nullptr, 1u, 0, nullptr, nullptr, 0u};
    if (tile) {
        DeviceBuffer stats, slow;
        PGEOF_TRY(slow.alloc(16 + (clipped ? 2 : 1) * n_query * sizeof(uint2), stream));
        a.slow_count = slow.as<uint32_t>();
        a.slow_list = reinterpret_cast<uint2*>(slow.as<unsigned char>() + 16);
        a.unsafe_count = a.slow_count + 1;
        a.unsafe_list = a.slow_list + (clipped ? n_query : 0);   // never written on a full grid (rmax_safe = +inf)
        PGEOF_CUDA(cudaMemsetAsync(a.slow_count, 0, 16, stream));
        const bool want_stats = env_float("PGEOF_KNN_STATS", 0.f) != 0.f;
        if (want_stats) {
            PGEOF_TRY(stats.alloc(ST_N * sizeof(unsigned long long), stream));
            PGEOF_CUDA(cudaMemsetAsync(stats.ptr, 0, ST_N * sizeof(unsigned long long), stream));
            a.stats = stats.as<unsigned long long>();
        }
        // kNN on the full grid through the default tile kernels: the queued queries run from here (two levels), not from launch_tile
        const bool two_level = mode == SEARCH_KNN && !clipped && k <= 64 && !pair && env_float("PGEOF_KNN_TWO_LEVEL", 1.f) != 0.f &&
                               env_float("PGEOF_KNN_ROLLED", 0.f) == 0.f && env_float("PGEOF_KNN_LOCK", 1.f) != 0.f && env_float("PGEOF_KNN_WARPS", 4.f) == 4.f;
        const bool two_level_now = two_level;
        int st;
        if (mode == SEARCH_RADIUS) st = k <= 32 ? launch_tile<32, 32, 2, SEARCH_RADIUS>(grid.view, a, stream) : launch_tile<64, 32, 4, SEARCH_RADIUS>(grid.view, a, stream);
        else if (k <= 32) st = launch_tile<32, 32>(grid.view, a, stream, !two_level);
        else if (pair && k <= 64) {
            const int variant = (int)env_float("PGEOF_PAIR_VARIANT", 0.f);
            st = variant == 1 ? launch_pair<48, 32, 3, 640, 3>(grid.view, a, stream)
               : variant == 2 ? launch_pair<48, 32, 5, 640, 2>(grid.view, a, stream)
                              : launch_pair<48, 32, 4, 512, 3>(grid.view, a, stream);
        }
        else if (k > 64) st = launch_tile<128, 32, 5, SEARCH_KNN, false, true, true>(grid.view, a, stream);
        else if (env_float("PGEOF_KNN_ROLLED", 0.f) != 0.f) st = launch_tile<64, 32, 4, SEARCH_KNN, false, true, true>(grid.view, a, stream);   // A/B switches
        else if (env_float("PGEOF_KNN_LOCK", 1.f) == 0.f) st = launch_tile<64, 32, 4, SEARCH_KNN, false, false>(grid.view, a, stream);
        else st = env_float("PGEOF_KNN_WARPS", 4.f) == 4.f ? launch_tile<64, 32, 4>(grid.view, a, stream, !two_level) : launch_tile<64, 32>(grid.view, a, stream);
        // Two-level handling of the queued queries (non-uniform clouds): the cell edge is sized for where most points live, so in
        // the sparse parts a ball crosses hundreds of cell rows that hold a handful of points each and the warp-per-query routine
        // spends its time walking rows.  Those queries (start radius > 1.5 cell edges) are deferred by the first run and, if there
        // are enough of them to pay for a second index, answered on a grid with 3x the cell edge (sweep: profiles/r2_summary.md).  Exact either way (the routine
        // is exact on any grid); uniform clouds defer nothing and pay one 4-byte read-back.
        if (st == PGEOF_OK && two_level_now) {
            DeviceBuffer defer;
            PGEOF_TRY(defer.alloc(16 + n_query * sizeof(uint2), stream));
            a.defer_count = defer.as<uint32_t>();
            a.defer_list = reinterpret_cast<uint2*>(defer.as<unsigned char>() + 16);
            a.r_split = env_float("PGEOF_KNN_SPLIT", 1.5f) * grid.view.h;
            PGEOF_CUDA(cudaMemsetAsync(a.defer_count, 0, 16, stream));
            auto run_list = [&](const GridView& gv, const SearchArgs& args) -> int {
                KernelTimer timer("knn_search", stream);
                if (k <= 32) knn_slow_kernel<32><<<sm_count() * 4, kWarps * 32, 0, stream>>>(gv, args);
                else knn_slow_kernel<64><<<sm_count() * 4, kWarps * 32, 0, stream>>>(gv, args);
                PGEOF_LAUNCH_CHECK();
                return PGEOF_OK;
            };
            PGEOF_TRY(run_list(grid.view, a));
            uint32_t n_defer = 0;
            PGEOF_CUDA(cudaMemcpyAsync(&n_defer, a.defer_count, sizeof(n_defer), cudaMemcpyDeviceToHost, stream));
            PGEOF_CUDA(cudaStreamSynchronize(stream));
            if (n_defer) {
                SearchArgs b = a;
                b.slow_list = a.defer_list;
                b.slow_count = a.defer_count;
                b.defer_list = nullptr;
                b.defer_count = nullptr;
                Grid coarse;
                const bool build = n_defer >= (uint32_t)env_float("PGEOF_KNN_COARSE_MIN", 50000.f);
                if (build) PGEOF_TRY(grid_build(data, n_data, env_float("PGEOF_KNN_COARSE", 3.f) * grid.view.h, 0.f, xf, stream, &coarse));
                PGEOF_TRY(run_list(build ? coarse.view : grid.view, b));
                if (env_float("PGEOF_KNN_STATS", 0.f) != 0.f)
                    std::fprintf(stderr, "[pgeof knn tile] deferred %u queued queries to a %s grid (cell edge %.3f)\n", n_defer, build ? "coarser" : "the same", build ? coarse.view.h : grid.view.h);
            }
        }
        uint32_t n_unsafe = 0;
        if (st == PGEOF_OK && clipped && mode == SEARCH_KNN) {
            PGEOF_CUDA(cudaMemcpyAsync(&n_unsafe, a.unsafe_count, sizeof(n_unsafe), cudaMemcpyDeviceToHost, stream));
            PGEOF_CUDA(cudaStreamSynchronize(stream));
            if (n_unsafe) {   // balls that outgrew the halo: the same generic routine over those queries on the FULL grid
                Grid full;
                PGEOF_TRY(grid_build(data, n_data, 0.f, occ, xf, stream, &full));
                SearchArgs b = a;
                b.slow_list = a.unsafe_list;
                b.slow_count = a.unsafe_count;
                if (k <= 32) knn_slow_kernel<32><<<sm_count() * 4, kWarps * 32, 0, stream>>>(full.view, b);
                else if (k <= 64) knn_slow_kernel<64><<<sm_count() * 4, kWarps * 32, 0, stream>>>(full.view, b);
                else knn_slow_kernel<128><<<sm_count() * 4, kWarps * 32, 0, stream>>>(full.view, b);
                PGEOF_LAUNCH_CHECK();
            }
        }
        if (st == PGEOF_OK && want_stats) {
            unsigned long long h[ST_N];
            PGEOF_CUDA(cudaMemcpyAsync(h, stats.ptr, sizeof(h), cudaMemcpyDeviceToHost, stream));
            PGEOF_CUDA(cudaStreamSynchronize(stream));
            uint32_t n_slow = 0;
            PGEOF_CUDA(cudaMemcpyAsync(&n_slow, a.slow_count, sizeof(n_slow), cudaMemcpyDeviceToHost, stream));
            PGEOF_CUDA(cudaStreamSynchronize(stream));
            std::fprintf(stderr, "[pgeof knn tile] clipped grid=%d (halo %.3g) re-run on the full grid=%u\n", (int)clipped, clipped ? clip.rmax_safe : 0.f, n_unsafe);
            std::fprintf(stderr, "[pgeof knn tile] n=%zu k=%u target=%.1f h=%.3f: generic queries=%u | per-pass events: short=%llu over=%llu tie=%llu region=%llu "
                         "fixed rows=%llu passes=%llu candidates/pass=%.1f survivors/query=%.1f\n", n_query, k, target, grid.view.h, n_slow,
                         h[ST_SHORT], h[ST_OVER], h[ST_TIE], h[ST_REGION], h[ST_FIXED], h[ST_PASSES],
                         (double)h[ST_CANDS] / (double)std::max(h[ST_PASSES], 1ull), (double)h[ST_SURV] / (double)n_query);
        }
        return st;
    }
    switch (mode) {
        case SEARCH_KNN: return dispatch_search<SEARCH_KNN>(k, grid.view, a, n_data, stream);
        case SEARCH_RADIUS: return dispatch_search<SEARCH_RADIUS>(k, grid.view, a, n_data, stream);
        case SEARCH_RADIUS_COUNT: return launch_search<32, SEARCH_RADIUS_COUNT>(grid.view, a, stream);
        case SEARCH_RADIUS_CSR: return dispatch_search<SEARCH_RADIUS_CSR>(k, grid.view, a, n_data, stream);
    }
    return PGEOF_EINVAL;
}

// Fused knn_features: knn_search(xyz, xyz, k) + compute_features without ever writing the (idx, d2) rows
// or reading them back: the tile kernel accumulates the moments of the k nearest from its staged candidates.  The ~2 %
// of the queries it queues go through the generic search into a compact scratch CSR and the ordinary feature kernel.
int knn_features_fused_run(const float* xyz, size_t n, uint32_t k, uint32_t k_min, int eig_order, float* features, cudaStream_t stream, int* done)
{
    *done = 0;
    if (n == 0 || k == 0 || k > 64 || n > 0xfffffff0ull || env_float("PGEOF_KNN_FUSED", 1.f) == 0.f) return PGEOF_OK;
    float target = (float)k + env_float("PGEOF_KNN_Z", 2.6f) * std::sqrt((float)k) + 2.f;
    target = std::min(target, 0.5f * (float)(k + tile_nload(k)));
    const float occ = std::max(2.f, target * env_float("PGEOF_KNN_CELL_OCC", 0.28f));
    Grid grid;
    PGEOF_TRY(grid_build(xyz, n, 0.f, occ, (int)env_float("PGEOF_KNN_XF", 8.f), stream, &grid));
    const uint32_t cap = (uint32_t)std::max<size_t>(n / 8, 4096);
    DeviceBuffer slow, tmp_idx, tmp_rows;
    PGEOF_TRY(slow.alloc(16 + n * sizeof(uint2), stream));
    PGEOF_TRY(tmp_idx.alloc((size_t)cap * k * sizeof(uint32_t), stream));
    PGEOF_TRY(tmp_rows.alloc((size_t)cap * sizeof(uint32_t), stream));
    SearchArgs a{grid.view.pts, (uint32_t)n, k, 0.f, target, nullptr, nullptr, nullptr, nullptr, (uint32_t)env_float("PGEOF_KNN_FLAGS", 0.f),
                 nullptr, nullptr, nullptr, nullptr, features, k_min, eig_order, tmp_idx.as<uint32_t>(), tmp_rows.as<uint32_t>(), cap};
    a.slow_count = slow.as<uint32_t>();
    a.slow_list = reinterpret_cast<uint2*>(slow.as<unsigned char>() + 16);
    a.unsafe_count = a.slow_count + 1;
    a.unsafe_list = a.slow_list;
    PGEOF_CUDA(cudaMemsetAsync(a.slow_count, 0, 16, stream));
    if (k <= 32) PGEOF_TRY((launch_tile<32, 32, 2, SEARCH_KNN, true>(grid.view, a, stream)));
    else PGEOF_TRY((launch_tile<64, 32, 4, SEARCH_KNN, true>(grid.view, a, stream)));
    uint32_t n_slow = 0;
    PGEOF_CUDA(cudaMemcpyAsync(&n_slow, a.slow_count, sizeof(n_slow), cudaMemcpyDeviceToHost, stream));
    PGEOF_CUDA(cudaStreamSynchronize(stream));
    if (n_slow > cap) return PGEOF_OK;              // unusually many queued queries (very non-uniform data): the caller runs the two kernels
    if (n_slow) {
        DeviceBuffer ptr;
        PGEOF_TRY(ptr.alloc(((size_t)n_slow + 1) * sizeof(uint32_t), stream));
        iota_scale_u32<<<(n_slow + 1 + 255) / 256, 256, 0, stream>>>(ptr.as<uint32_t>(), (size_t)n_slow + 1, k);
        PGEOF_LAUNCH_CHECK();
        PGEOF_TRY(features_run(xyz, n, tmp_idx.as<uint32_t>(), (size_t)n_slow * k, ptr.as<uint32_t>(), n_slow, k_min, eig_order, features, stream,
                               tmp_rows.as<uint32_t>()));
    }
    *done = 1;
    return PGEOF_OK;
}

}  // namespace pgeof
