// selected.cu -- compute_features_selected: fused radius search + PCA + chosen features,
// float and double.  Replaces compute_geometric_features_selected (include/pgeof.hpp:325-375)
// and compute_selected_features (include/pca.hpp:212-295).
//
// One warp per point.  Pass 1 scans the cells intersecting the ball, counts the points with
// d2 < r*r (strict, pgeof.hpp:348-352) and accumulates their origin-shifted moments on the
// fly (origin = the query point itself).  If the ball holds more than max_knn points the
// max_knn nearest by (d2, index) are isolated with a counting bisection on the distance bit
// pattern (and on the index inside an exact-distance tie), then one more pass accumulates
// the moments of exactly that set.  No neighbour list is ever materialised.
//
// The double flavour keeps the reference's double metric: the grid is built over the
// float-rounded cloud, the ball radius is inflated by the rounding slack so coverage stays
// exact, and distances / moments / eigen solve run in double on the original coordinates.
#include <cmath>

#include "eig3.cuh"
#include "search_core.cuh"

namespace pgeof {

namespace {

constexpr int kWarps = 8;

template <typename T> struct Bits;
template <> struct Bits<float> {
    typedef uint32_t U;
    static __device__ __forceinline__ U of(float v) { return __float_as_uint(v); }
    static __device__ __forceinline__ float from(U u) { return __uint_as_float(u); }
    static constexpr U kMax = 0xffffffffu;
};
template <> struct Bits<double> {
    typedef unsigned long long U;
    static __device__ __forceinline__ U of(double v) { return (U)__double_as_longlong(v); }
    static __device__ __forceinline__ double from(U u) { return __longlong_as_double((long long)u); }
    static constexpr U kMax = ~0ull;
};

__device__ __forceinline__ double sqdist_f64(double qx, double qy, double qz, double px, double py, double pz)
{
    const double dx = __dsub_rn(qx, px), dy = __dsub_rn(qy, py), dz = __dsub_rn(qz, pz);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

template <typename T>
struct Mom {
    T sx = 0, sy = 0, sz = 0, sxx = 0, sxy = 0, sxz = 0, syy = 0, syz = 0, szz = 0;
    __device__ __forceinline__ void add(T dx, T dy, T dz)
    {
        sx += dx; sy += dy; sz += dz;
        sxx += dx * dx; sxy += dx * dy; sxz += dx * dz; syy += dy * dy; syz += dy * dz; szz += dz * dz;
    }
    __device__ __forceinline__ void warp_reduce()
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sx += __shfl_xor_sync(kFull, sx, o); sy += __shfl_xor_sync(kFull, sy, o); sz += __shfl_xor_sync(kFull, sz, o);
            sxx += __shfl_xor_sync(kFull, sxx, o); sxy += __shfl_xor_sync(kFull, sxy, o); sxz += __shfl_xor_sync(kFull, sxz, o);
            syy += __shfl_xor_sync(kFull, syy, o); syz += __shfl_xor_sync(kFull, syz, o); szz += __shfl_xor_sync(kFull, szz, o);
        }
    }
};

// Threshold in the (d2 bits, index) lexicographic order: accept iff key <= (d2, idx).
template <typename T>
struct Thr {
    typename Bits<T>::U d2;
    uint32_t idx;
    __device__ __forceinline__ bool accepts(typename Bits<T>::U b, uint32_t i) const { return b < d2 || (b == d2 && i <= idx); }
};

// One pass over the cells intersecting ball(qf, Rg): counts the points under `thr` and, when
// ACC, accumulates their moments relative to q.
template <typename T, bool ACC>
__device__ __forceinline__ uint32_t sel_scan(const GridView& g, const T* __restrict__ xyz, float qfx, float qfy, float qfz, T qx, T qy,
                                             T qz, float Rg, Thr<T> thr, Mom<T>* mom, int lane)
{
    const int cx0 = cell_coord(__fsub_rd(qfx, Rg), g.lo[0], g.inv_hx, g.n[0]);
    const int cx1 = cell_coord(__fadd_ru(qfx, Rg), g.lo[0], g.inv_hx, g.n[0]);
    const int cy0 = cell_coord(__fsub_rd(qfy, Rg), g.lo[1], g.inv_h, g.n[1]);
    const int cy1 = cell_coord(__fadd_ru(qfy, Rg), g.lo[1], g.inv_h, g.n[1]);
    const int cz0 = cell_coord(__fsub_rd(qfz, Rg), g.lo[2], g.inv_h, g.n[2]);
    const int cz1 = cell_coord(__fadd_ru(qfz, Rg), g.lo[2], g.inv_h, g.n[2]);
    const int cqy = cell_coord(qfy, g.lo[1], g.inv_h, g.n[1]);
    const int cqz = cell_coord(qfz, g.lo[2], g.inv_h, g.n[2]);
    const uint32_t nyr = (uint32_t)(cy1 - cy0 + 1);
    const uint32_t nrows = nyr * (uint32_t)(cz1 - cz0 + 1);
    const float R2u = __fmul_ru(Rg, Rg);
    uint32_t count = 0;
    for (uint32_t rbase = 0; rbase < nrows; rbase += 32) {
        const uint32_t r = rbase + lane;
        uint32_t s = 0, e = 0;
        if (r < nrows) {
            const int cz = cz0 + (int)(r / nyr), cy = cy0 + (int)(r % nyr);
            const float gy = axis_gap(qfy, cy, cqy, g.lo[1], g.h, g.slack);
            const float gz = axis_gap(qfz, cz, cqz, g.lo[2], g.h, g.slack);
            const float rem = __fsub_ru(__fsub_ru(R2u, __fmul_rd(gy, gy)), __fmul_rd(gz, gz));
            if (rem >= 0.f) {
                const float xr = __fsqrt_ru(rem);
                const int x0 = max(cx0, cell_coord(__fsub_rd(qfx, xr), g.lo[0], g.inv_hx, g.n[0]));
                const int x1 = min(cx1, cell_coord(__fadd_ru(qfx, xr), g.lo[0], g.inv_hx, g.n[0]));
                const uint32_t row = ((uint32_t)cz * (uint32_t)g.n[1] + (uint32_t)cy) * (uint32_t)g.n[0];
                s = __ldg(g.cell_start + row + x0);
                e = __ldg(g.cell_start + row + x1 + 1);
            }
        }
        unsigned nonempty = __ballot_sync(kFull, e > s);
        while (nonempty) {
            const int rr = __ffs(nonempty) - 1;
            nonempty &= nonempty - 1;
            const uint32_t s_r = __shfl_sync(kFull, s, rr), e_r = __shfl_sync(kFull, e, rr);
            for (uint32_t base = s_r; base < e_r; base += 32) {
                const uint32_t j = base + lane;
                bool acc = false;
                if (j < e_r) {
                    const float4 p = __ldg(g.pts + j);
                    const uint32_t idx = __float_as_uint(p.w);
                    T px, py, pz, d2;
                    if constexpr (sizeof(T) == 4) {
                        px = p.x; py = p.y; pz = p.z;
                        d2 = sqdist_f32(qx, qy, qz, px, py, pz);
                    } else {
                        px = __ldg(xyz + 3 * (size_t)idx); py = __ldg(xyz + 3 * (size_t)idx + 1); pz = __ldg(xyz + 3 * (size_t)idx + 2);
                        d2 = sqdist_f64(qx, qy, qz, px, py, pz);
                    }
                    acc = thr.accepts(Bits<T>::of(d2), idx);
                    if (ACC && acc) mom->add(px - qx, py - qy, pz - qz);
                }
                count += __popc(__ballot_sync(kFull, acc));
            }
        }
    }
    return count;
}

template <typename T>
struct SelArgs {
    const T* xyz; uint32_t n; T radius; uint32_t max_knn;
    const int32_t* ids; uint32_t n_ids; int eig_order; T* out;
};

template <typename T>
__global__ void __launch_bounds__(kWarps * 32) selected_kernel(const GridView g, const SelArgs<T> a)
{
    typedef typename Bits<T>::U U;
    const int lane = threadIdx.x & 31;
    const uint32_t w = blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (w >= a.n) return;
    const float4 q4 = __ldg(g.pts + w);
    const uint32_t row = __float_as_uint(q4.w);
    T qx, qy, qz;
    if constexpr (sizeof(T) == 4) { qx = q4.x; qy = q4.y; qz = q4.z; }
    else { qx = a.xyz[3 * (size_t)row]; qy = a.xyz[3 * (size_t)row + 1]; qz = a.xyz[3 * (size_t)row + 2]; }
    T* out = a.out + (size_t)row * a.n_ids;

    const T r2 = a.radius * a.radius;                      // pgeof.hpp:336
    uint32_t c = 0, found = 0;   // found = points in the ball, c = points kept (<= max_knn)
    Mom<T> mom;
    float Rg = 0.f;
    if (r2 > T(0)) {
        // geometric radius covering every stored (float-rounded) point of the exact ball
        if constexpr (sizeof(T) == 4) Rg = __fmul_ru(__fsqrt_ru(r2), 1.0001f);
        else Rg = __fadd_ru(__fmul_ru(__double2float_ru(sqrt(r2)), 1.0001f), g.slack);
        Thr<T> thr{(U)(Bits<T>::of(r2) - 1), 0xffffffffu};  // d2 < r2
        c = sel_scan<T, true>(g, a.xyz, q4.x, q4.y, q4.z, qx, qy, qz, Rg, thr, &mom, lane);
        found = c;
        if (c > a.max_knn && a.max_knn > 0) {
            // Isolate the max_knn nearest by (d2, index): the threshold (D, I) with exactly
            // `need` keys <= (D, I) exists because indices are distinct.
            const uint32_t need = a.max_knn;
            auto count_le = [&](U d, uint32_t i) {
                return sel_scan<T, false>(g, a.xyz, q4.x, q4.y, q4.z, qx, qy, qz, Rg, Thr<T>{d, i}, nullptr, lane);
            };
            // phase 1: smallest D with #{d2 bits <= D} >= need (density-interpolated lower-bound search)
            U lo = 0, hi = thr.d2;
            bool exact = false;
            float cur_cnt = (float)c;
            T cur_d2 = r2;
            for (int it = 0; lo < hi; ++it) {
                U mid = lo + (hi - lo) / 2;
                if (it < 4) {
                    const float ratio = exp2f(0.6666667f * log2f(((float)need + 0.5f) / cur_cnt));
                    const U guess = Bits<T>::of(cur_d2 * (T)ratio);
                    if (guess >= lo && guess < hi) mid = guess;
                }
                const uint32_t n = count_le(mid, 0xffffffffu);
                if (n == need) { thr = Thr<T>{mid, 0xffffffffu}; exact = true; break; }
                if (n > need) hi = mid; else lo = mid + 1;
                cur_cnt = fmaxf((float)n, 0.5f);
                cur_d2 = Bits<T>::from(mid);
            }
            if (!exact) {
                // phase 2: an exact-distance tie straddles the boundary at D = lo; smallest index bound reaching `need`
                const U D = lo;
                uint32_t ilo = 0, ihi = 0xffffffffu;
                while (ilo < ihi) {
                    const uint32_t imid = ilo + (ihi - ilo) / 2;
                    const uint32_t n = count_le(D, imid);
                    if (n == need) { ilo = imid; break; }
                    if (n > need) ihi = imid; else ilo = imid + 1;
                }
                thr = Thr<T>{D, ilo};
            }
            mom = Mom<T>();
            c = sel_scan<T, true>(g, a.xyz, q4.x, q4.y, q4.z, qx, qy, qz, Rg, thr, &mom, lane);
        }
    }
    if (found < 2 || c == 0 || a.max_knn == 0) {            // pgeof.hpp:355 -> zeros (calloc)
        for (uint32_t f = lane; f < a.n_ids; f += 32) out[f] = T(0);
        return;
    }
    mom.warp_reduce();
    const T inv = T(1) / (T)c;
    const T mx = mom.sx * inv, my = mom.sy * inv, mz = mom.sz * inv;
    const Pca<T> p = pca_from_cov<T>(mom.sxx * inv - mx * mx, mom.sxy * inv - mx * my, mom.sxz * inv - mx * mz,
                                     mom.syy * inv - my * my, mom.syz * inv - my * mz, mom.szz * inv - mz * mz, a.eig_order);
    for (uint32_t f = lane; f < a.n_ids; f += 32) out[f] = feature_selected<T>(p, a.ids[f]);
}

__global__ void f64_to_f32_kernel(const double* __restrict__ in, float* __restrict__ out, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __double2float_rn(in[i]);
}

template <typename T>
int selected_run(const T* xyz, const float* xyz_f32, size_t n, T radius, uint32_t max_knn, const int32_t* ids_host, size_t n_ids,
                 int eig_order, T* out, cudaStream_t stream)
{
    if (n == 0 || n_ids == 0) return PGEOF_OK;
    // the reference only ever uses r * r (pgeof.hpp:336): a negative radius searches the ball of |r|
    if (!std::isfinite((double)radius)) { set_error("search_radius must be finite"); return PGEOF_EINVAL; }
    radius = radius < T(0) ? -radius : radius;
    if (eig_order != PGEOF_EIG_LITERAL && eig_order != PGEOF_EIG_DOCUMENTED) { set_error("bad eig_order %d", eig_order); return PGEOF_EINVAL; }
    Grid grid;
    const float edge = (float)radius > 0.f ? (float)radius : 1.f;
    PGEOF_TRY(grid_build(xyz_f32, n, edge, 0.f, 1, stream, &grid));
    DeviceBuffer ids;
    PGEOF_TRY(ids.alloc(n_ids * sizeof(int32_t), stream));
    PGEOF_CUDA(cudaMemcpyAsync(ids.ptr, ids_host, n_ids * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
    SelArgs<T> a{xyz, (uint32_t)n, radius, max_knn, ids.as<int32_t>(), (uint32_t)n_ids, eig_order, out};
    const unsigned blocks = (unsigned)((n + kWarps - 1) / kWarps);
    {
        KernelTimer timer("selected", stream);
        selected_kernel<T><<<blocks, kWarps * 32, 0, stream>>>(grid.view, a);
    }
    PGEOF_LAUNCH_CHECK();
    // ids_host may be a temporary of the caller: the pageable H2D copy above is staged
    // synchronously by the runtime, so no extra synchronisation is needed here.
    return PGEOF_OK;
}

}  // namespace

int selected_run_f32(const float* xyz, size_t n, float radius, uint32_t max_knn, const int32_t* ids_host, size_t n_ids,
                     int eig_order, float* out, cudaStream_t stream)
{
    return selected_run<float>(xyz, xyz, n, radius, max_knn, ids_host, n_ids, eig_order, out, stream);
}

int selected_run_f64(const double* xyz, size_t n, double radius, uint32_t max_knn, const int32_t* ids_host, size_t n_ids,
                     int eig_order, double* out, cudaStream_t stream)
{
    if (n == 0 || n_ids == 0) return PGEOF_OK;
    DeviceBuffer f32;
    PGEOF_TRY(f32.alloc(3 * n * sizeof(float), stream));
    f64_to_f32_kernel<<<(unsigned)((3 * n + 255) / 256), 256, 0, stream>>>(xyz, f32.as<float>(), 3 * n);
    PGEOF_LAUNCH_CHECK();
    return selected_run<double>(xyz, f32.as<float>(), n, radius, max_knn, ids_host, n_ids, eig_order, out, stream);
}

}  // namespace pgeof
