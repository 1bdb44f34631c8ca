// search_core.cuh -- warp-per-query neighbour search primitives.
//
// A candidate is the 64-bit key (bits(d2) << 32) | original_index.  d2 >= 0, so the
// float bit pattern is order preserving and one unsigned 64-bit compare realises the
// (d2, index) order that makes neighbour sets a pure function of the input
// (nanoflann leaves ties to traversal order).
//
// d2 is the defined float32 metric of nanoflann's L2_Simple (nn_search.hpp:35):
//   d2 = fl(fl(fl(dx*dx) + fl(dy*dy)) + fl(dz*dz)),  no FMA contraction.
#pragma once
#include "grid.cuh"

namespace pgeof {

typedef unsigned long long u64;
constexpr u64 kKeyMax = ~0ull;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float sqdist_f32(float qx, float qy, float qz, float px, float py, float pz)
{
    const float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// FMA-contracted variant for filtering only (within 2^-21 relative of sqdist_f32; never returned to the caller)
__device__ __forceinline__ float sqdist_fused(float qx, float qy, float qz, float px, float py, float pz)
{
    const float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

__device__ __forceinline__ u64 make_key(float d2, uint32_t idx) { return ((u64)__float_as_uint(d2) << 32) | idx; }
__device__ __forceinline__ float key_d2(u64 k) { return __uint_as_float((uint32_t)(k >> 32)); }
__device__ __forceinline__ uint32_t key_idx(u64 k) { return (uint32_t)k; }

// Acceptance threshold that is strictly inside the geometric radius R used for cell
// coverage: d2_f32 <= T2 implies the exact distance is < R (float error of d2 is ~2^-22
// relative, the margin is 1e-4).
__device__ __forceinline__ u64 tau_from_radius(float R)
{
    const float t2 = __fmul_rd(__fmul_rd(R, R), 0.9999f);
    return ((u64)__float_as_uint(t2) << 32) | 0xffffffffull;
}

// Lower bound of |p - q| along one axis for any point p stored in cell `c`.
__device__ __forceinline__ float axis_gap(float q, int c, int cq, float lo, float h, float slack)
{
    if (c == cq) return 0.f;
    if (c > cq) {
        const float b = __fsub_rd(__fmaf_rd((float)c, h, lo), slack);
        return fmaxf(0.f, __fsub_rd(b, q));
    }
    const float b = __fadd_ru(__fmaf_ru((float)(c + 1), h, lo), slack);
    return fmaxf(0.f, __fsub_rd(q, b));
}

// Visits every point whose cell intersects ball(q, R) and appends the keys <= tau to
// keybuf (first CAP only).  Returns how many keys were <= tau (may exceed CAP).
// Rows of cells (fixed y, z) are contiguous spans of `pts`; 32 rows are resolved at a
// time (one per lane) so the cell_start loads of a chunk are a single round trip.
template <int CAP, bool STORE, bool PREFETCH = false>
__device__ __forceinline__ uint32_t scan_ball(const GridView& g, float qx, float qy, float qz, float R, u64 tau,
                                              u64* __restrict__ keybuf, int lane)
{
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
    uint32_t count = 0;

    for (uint32_t rbase = 0; rbase < nrows; rbase += 32) {
        const uint32_t r = rbase + lane;
        uint32_t s = 0, e = 0;
        if (r < nrows) {
            const int cz = cz0 + (int)(r / nyr), cy = cy0 + (int)(r % nyr);
            const float gy = axis_gap(qy, cy, cqy, g.lo[1], g.h, g.slack);
            const float gz = axis_gap(qz, cz, cqz, g.lo[2], g.h, g.slack);
            const float rem = __fsub_ru(__fsub_ru(R2u, __fmul_rd(gy, gy)), __fmul_rd(gz, gz));
            if (rem >= 0.f) {
                const float xr = __fsqrt_ru(rem);
                const int x0 = max(cx0, cell_coord(__fsub_rd(qx, xr), g.lo[0], g.inv_hx, g.n[0]));
                const int x1 = min(cx1, cell_coord(__fadd_ru(qx, xr), g.lo[0], g.inv_hx, g.n[0]));
                const uint32_t row = ((uint32_t)cz * (uint32_t)g.n[1] + (uint32_t)cy) * (uint32_t)g.n[0];
                s = __ldg(g.cell_start + row + x0);
                e = __ldg(g.cell_start + row + x1 + 1);
            }
        }
        unsigned nonempty = __ballot_sync(kFull, e > s);
        // PREFETCH (the queued queries of a tile kernel: latency bound, 30 % of the stall samples wait for these loads on
        // non-uniform clouds): the candidate loads are software pipelined -- the first 32 points of the NEXT row and the next
        // 32 of the current one are requested before the current 32 are evaluated.  The all-queries kernels are issue bound
        // (82 % issue-active at k = 100) and lose 10 % to the extra instructions, so they keep the plain loop.
        auto visit = [&](uint32_t j, uint32_t e_r, const float4& p) {
            bool acc = false;
            u64 key = 0;
            if (j < e_r) {
                key = make_key(sqdist_f32(qx, qy, qz, p.x, p.y, p.z), __float_as_uint(p.w));
                acc = key <= tau;
            }
            const unsigned m = __ballot_sync(kFull, acc);
            if (STORE && acc) {
                const uint32_t pos = count + __popc(m & lt);
                if (pos < CAP) keybuf[pos] = key;
            }
            count += __popc(m);
        };
        if constexpr (!PREFETCH) {
            while (nonempty) {
                const int rr = __ffs(nonempty) - 1;
                nonempty &= nonempty - 1;
                const uint32_t s_r = __shfl_sync(kFull, s, rr), e_r = __shfl_sync(kFull, e, rr);
                for (uint32_t base = s_r; base < e_r; base += 32) {
                    const uint32_t j = base + lane;
                    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (j < e_r) p = __ldg(g.pts + j);
                    visit(j, e_r, p);
                }
            }
        } else {
            float4 pf = make_float4(0.f, 0.f, 0.f, 0.f);      // prefetched head of the row about to be scanned
            if (nonempty) {
                const int r0 = __ffs(nonempty) - 1;
                const uint32_t s0 = __shfl_sync(kFull, s, r0), e0 = __shfl_sync(kFull, e, r0);
                if (s0 + lane < e0) pf = __ldg(g.pts + s0 + lane);
            }
            while (nonempty) {
                const int rr = __ffs(nonempty) - 1;
                nonempty &= nonempty - 1;
                const uint32_t s_r = __shfl_sync(kFull, s, rr), e_r = __shfl_sync(kFull, e, rr);
                float4 p = pf;
                if (nonempty) {
                    const int r2 = __ffs(nonempty) - 1;
                    const uint32_t s2 = __shfl_sync(kFull, s, r2), e2 = __shfl_sync(kFull, e, r2);
                    if (s2 + lane < e2) pf = __ldg(g.pts + s2 + lane);
                }
                for (uint32_t base = s_r; base < e_r; base += 32) {
                    const uint32_t j = base + lane;
                    float4 pn = p;
                    if (j + 32 < e_r) pn = __ldg(g.pts + j + 32);
                    visit(j, e_r, p);
                    p = pn;
                }
            }
        }
    }
    __syncwarp();
    return count;
}

// Points and cells of the (clipped) 3x3x3 block around the query's cell: the local
// density estimate that seeds the search radius.
__device__ __forceinline__ void block27_count(const GridView& g, float qx, float qy, float qz, int lane, uint32_t* pts, uint32_t* cells)
{
    const int cqx = cell_coord(qx, g.lo[0], g.inv_hx, g.n[0]);
    const int cqy = cell_coord(qy, g.lo[1], g.inv_h, g.n[1]);
    const int cqz = cell_coord(qz, g.lo[2], g.inv_h, g.n[2]);
    const int bx0 = max(cqx - g.xf, 0), bx1 = min(cqx + g.xf, g.n[0] - 1);
    uint32_t c = 0, nc = 0;
    if (lane < 9) {
        const int cy = cqy + lane % 3 - 1, cz = cqz + lane / 3 - 1;
        if (cy >= 0 && cy < g.n[1] && cz >= 0 && cz < g.n[2]) {
            const uint32_t row = ((uint32_t)cz * (uint32_t)g.n[1] + (uint32_t)cy) * (uint32_t)g.n[0];
            c = __ldg(g.cell_start + row + bx1 + 1) - __ldg(g.cell_start + row + bx0);
            nc = (uint32_t)(bx1 - bx0 + 1);
        }
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        c += __shfl_xor_sync(kFull, c, o);
        nc += __shfl_xor_sync(kFull, nc, o);
    }
    *pts = __shfl_sync(kFull, c, 0);
    *cells = __shfl_sync(kFull, nc, 0);
}

// distance from q to the grid's bounding box (0 inside)
__device__ __forceinline__ float bbox_distance(const GridView& g, float qx, float qy, float qz)
{
    const float q[3] = {qx, qy, qz};
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float hi = fmaf((float)g.n[d], d == 0 ? g.hx : g.h, g.lo[d]);
        const float gap = fmaxf(fmaxf(g.lo[d] - q[d], q[d] - hi), 0.f);
        s += gap * gap;
    }
    return sqrtf(s);
}

// count of keys <= t among the lane-distributed registers key[0..MC)
template <int MC>
__device__ __forceinline__ uint32_t count_le(const u64 (&key)[MC], u64 t, uint32_t c)
{
    uint32_t n = 0;
#pragma unroll
    for (int r = 0; r < MC; ++r)
        if ((uint32_t)(r * 32) < c) n += __popc(__ballot_sync(kFull, key[r] <= t));
    return n;
}

// Finds t with need <= #{key <= t} <= cap_hi, given #{key <= hi} = c > cap_hi.
// Density interpolation (count ~ d2^(3/2)) for the first tries, then plain bisection
// of the 64-bit key space, which always terminates because keys are distinct.
template <int MC>
__device__ __forceinline__ u64 select_threshold(const u64 (&key)[MC], uint32_t c, u64 hi, uint32_t need, uint32_t cap_hi)
{
    u64 lo = 0;
    float cur_cnt = (float)c, cur_d2 = key_d2(hi);
    const float target = 0.5f * (float)(need + cap_hi);
    for (int it = 0; it < 96; ++it) {   // 64-bit bisection needs <= 64 steps; the cap only guards against a hang
        u64 mid = lo + (hi - lo) / 2;
        if (it < 3) {
            const float gd2 = cur_d2 * exp2f(0.6666667f * log2f(target / cur_cnt));
            const u64 guess = ((u64)__float_as_uint(gd2) << 32) | 0xffffffffull;
            if (guess > lo && guess < hi) mid = guess;
        }
        const uint32_t n = count_le<MC>(key, mid, c);
        if (n < need) lo = mid;
        else if (n > cap_hi) hi = mid;
        else return mid;
        cur_cnt = fmaxf((float)n, 0.5f);
        cur_d2 = key_d2(mid);
    }
    return hi;
}

// Bitonic sort of 32*M keys held as v[m] at element index m*32 + lane, ascending.
template <int M>
__device__ __forceinline__ void warp_bitonic_sort(u64 (&v)[M], int lane)
{
#pragma unroll
    for (int size = 2; size <= 32 * M; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= 32) {
                const int ms = stride >> 5;
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    if ((m & ms) == 0) {
                        const bool asc = (((m << 5) & size) == 0);
                        const u64 a = v[m], b = v[m | ms];
                        const bool sw = (a > b) == asc;
                        v[m] = sw ? b : a;
                        v[m | ms] = sw ? a : b;
                    }
                }
            } else {
                const bool lower = (lane & stride) == 0;
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const u64 o = __shfl_xor_sync(kFull, v[m], stride);
                    const bool asc = ((((m << 5) | lane) & size) == 0);
                    const bool keep_min = (lower == asc);
                    const bool o_less = o < v[m];
                    v[m] = (o_less == keep_min) ? o : v[m];
                }
            }
        }
    }
}

// The same network on 32-bit keys (one SHFL + one min/max per element and stage instead of two SHFLs and a 64-bit compare).
template <int M>
__device__ __forceinline__ void warp_bitonic_sort32(uint32_t (&v)[M], int lane)
{
#pragma unroll
    for (int size = 2; size <= 32 * M; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= 32) {
                const int ms = stride >> 5;
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    if ((m & ms) == 0) {
                        const bool asc = (((m << 5) & size) == 0);
                        const uint32_t a = v[m], b = v[m | ms];
                        v[m] = asc ? min(a, b) : max(a, b);
                        v[m | ms] = asc ? max(a, b) : min(a, b);
                    }
                }
            } else {
                const bool lower = (lane & stride) == 0;
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const uint32_t o = __shfl_xor_sync(kFull, v[m], stride);
                    const bool asc = ((((m << 5) | lane) & size) == 0);
                    v[m] = (lower == asc) ? min(v[m], o) : max(v[m], o);
                }
            }
        }
    }
}

// v holds 32 * M keys at element index m * 32 + lane: true iff they ascend
template <int M>
__device__ __forceinline__ bool warp_is_sorted(const u64 (&v)[M], int lane)
{
    bool inv = false;
#pragma unroll
    for (int m = 0; m < M; ++m) {
        u64 nxt = __shfl_down_sync(kFull, v[m], 1);
        if (m + 1 < M) { const u64 head = __shfl_sync(kFull, v[m + 1], 0); if (lane == 31) nxt = head; }
        else if (lane == 31) nxt = kKeyMax;
        inv = inv || nxt < v[m];
    }
    return __ballot_sync(kFull, inv) == 0u;
}

// one odd-even transposition round (both parities) on the same layout: fixes isolated adjacent swaps
template <int M>
__device__ __forceinline__ void warp_transpose_round(u64 (&v)[M], int lane)
{
    // pairs (e, e + 1), e even: lanes (2i, 2i + 1) of the same register
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const u64 o = __shfl_xor_sync(kFull, v[m], 1);
        const bool low = (lane & 1) == 0;
        v[m] = (low == (o < v[m])) ? o : v[m];
    }
    // pairs (e, e + 1), e odd: lanes (2i + 1, 2i + 2); lane 31 pairs with lane 0 of the next register
    u64 w[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        u64 o = __shfl_sync(kFull, v[m], (lane & 1) ? lane + 1 : lane - 1);      // (lanes 0 and 31 are overridden below)
        bool have = true;
        if (lane == 31) { have = m + 1 < M; if (m + 1 < M) o = 0; }
        if (lane == 0) { have = m > 0; }
        u64 head = 0, tail = 0;
        if (m + 1 < M) head = __shfl_sync(kFull, v[m + 1], 0);
        if (m > 0) tail = __shfl_sync(kFull, v[m - 1], 31);
        if (lane == 31 && m + 1 < M) o = head;
        if (lane == 0 && m > 0) o = tail;
        const bool low = (lane & 1) == 1;                                         // the odd lane holds the lower element of the pair
        w[m] = (have && (low == (o < v[m]))) ? o : v[m];
    }
#pragma unroll
    for (int m = 0; m < M; ++m) v[m] = w[m];
}

// Thread-local sorting networks on N 32-bit keys held in registers.  Every index is a compile-time
// constant after unrolling, so v[] never leaves the register file; a compare-exchange is one
// unsigned min / max pair (no predicates, no shuffles) and the 32 lanes of a warp sort 32
// independent lists at once.  Batcher odd-even merge sort: 191 compare-exchanges for 32 keys,
// 161 more to merge two sorted runs of 32.
template <int N>
__device__ __forceinline__ void reg_cmpswap(uint32_t (&v)[N], int i, int j)
{
    // a network laid over a VIRTUAL array longer than N (positions >= N hold +inf and never move: every comparator is
    // (low index <- min, high index <- max)) is pruned to the comparators between real registers
    if (i < N && j < N) {
        const uint32_t a = v[i], b = v[j];
        v[i] = min(a, b);
        v[j] = max(a, b);
    }
}

template <int N, int LO, int CNT, int R>
struct RegOddEvenMerge {
    static __device__ __forceinline__ void run(uint32_t (&v)[N])
    {
        constexpr int STEP = R * 2;
        if constexpr (STEP < CNT) {
            RegOddEvenMerge<N, LO, CNT, STEP>::run(v);
            RegOddEvenMerge<N, LO + R, CNT, STEP>::run(v);
#pragma unroll
            for (int i = LO + R; i + R < LO + CNT; i += STEP) reg_cmpswap<N>(v, i, i + R);
        } else {
            reg_cmpswap<N>(v, LO, LO + R);
        }
    }
};

template <int N, int LO, int CNT>
struct RegOddEvenSort {
    static __device__ __forceinline__ void run(uint32_t (&v)[N])
    {
        if constexpr (CNT > 1) {
            RegOddEvenSort<N, LO, CNT / 2>::run(v);
            RegOddEvenSort<N, LO + CNT / 2, CNT / 2>::run(v);
            RegOddEvenMerge<N, LO, CNT, 1>::run(v);
        }
    }
};

// v[LO, LO+CNT) is bitonic -> ascending (CNT a power of two)
template <int N, int LO, int CNT>
struct RegBitonicMerge {
    static __device__ __forceinline__ void run(uint32_t (&v)[N])
    {
        if constexpr (CNT > 1) {
#pragma unroll
            for (int i = 0; i < CNT / 2; ++i) reg_cmpswap<N>(v, LO + i, LO + i + CNT / 2);
            RegBitonicMerge<N, LO, CNT / 2>::run(v);
            RegBitonicMerge<N, LO + CNT / 2, CNT / 2>::run(v);
        }
    }
};

}  // namespace pgeof
