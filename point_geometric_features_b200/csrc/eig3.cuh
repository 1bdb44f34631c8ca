// eig3.cuh -- per-thread symmetric 3x3 eigen solve and the pgeof feature formulas.
//
// Replaces Eigen::SelfAdjointEigenSolver<Matrix3> (include/pca.hpp:79) by a register
// resident cyclic Jacobi iteration (quadratically convergent, orthonormal vectors to
// round-off, high relative accuracy on the small eigenvalue of a PSD matrix), and
// compute_features / compute_selected_features / compute_eigentropy
// (include/pca.hpp:140-295) by __device__ functions with the same constants.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/pgeof_b200.h"

namespace pgeof {

template <typename T> struct Real;
template <> struct Real<float> {
    static __device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
    static __device__ __forceinline__ float rsqrt_(float x) { return rsqrtf(x); }
    static __device__ __forceinline__ float abs_(float x) { return fabsf(x); }
    static __device__ __forceinline__ float max_(float a, float b) { return fmaxf(a, b); }
    static __device__ __forceinline__ float log_(float x) { return logf(x); }
    static __device__ __forceinline__ float cbrt_(float x) { return cbrtf(x); }
    static constexpr float kOffTol = 1e-9f;
    static constexpr int kSweeps = 8;
};
template <> struct Real<double> {
    static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
    static __device__ __forceinline__ double rsqrt_(double x) { return 1.0 / sqrt(x); }
    static __device__ __forceinline__ double abs_(double x) { return fabs(x); }
    static __device__ __forceinline__ double max_(double a, double b) { return fmax(a, b); }
    static __device__ __forceinline__ double log_(double x) { return log(x); }
    static __device__ __forceinline__ double cbrt_(double x) { return cbrt(x); }
    static constexpr double kOffTol = 1e-18;
    static constexpr int kSweeps = 12;
};

// PCAResult of include/pca.hpp:37-44
template <typename T>
struct Pca {
    T val[3];
    T v0[3], v1[3], v2[3];
};

template <typename T>
__device__ __forceinline__ void jacobi_rotate(T& app, T& aqq, T& apq, T& arp, T& arq, T (&V)[3][3], int p, int q)
{
    if (apq == T(0)) return;
    const T theta = (aqq - app) / (T(2) * apq);
    const T at = Real<T>::abs_(theta);
    T t = T(1) / (at + Real<T>::sqrt_(theta * theta + T(1)));
    if (theta < T(0)) t = -t;
    const T c = Real<T>::rsqrt_(t * t + T(1));
    const T s = t * c;
    app -= t * apq;
    aqq += t * apq;
    apq = T(0);
    const T rp = arp, rq = arq;
    arp = c * rp - s * rq;
    arq = s * rp + c * rq;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const T vp = V[i][p], vq = V[i][q];
        V[i][p] = c * vp - s * vq;
        V[i][q] = s * vp + c * vq;
    }
}

// Eigen-decomposition of the symmetric matrix [[a00 a01 a02] [a01 a11 a12] [a02 a12 a22]].
// Output ordered per `eig_order` (stable sort: increasing = PGEOF_EIG_LITERAL, what
// pca.hpp:79-89 yields with Eigen 3.4; decreasing = PGEOF_EIG_DOCUMENTED), eigenvalues
// clamped at 0 (pca.hpp:85), v2 flipped to z >= 0 (pca.hpp:96).
template <typename T>
__device__ __forceinline__ Pca<T> pca_from_cov(T a00, T a01, T a02, T a11, T a12, T a22, int eig_order)
{
    T scale = Real<T>::max_(Real<T>::max_(Real<T>::abs_(a00), Real<T>::abs_(a11)), Real<T>::abs_(a22));
    scale = Real<T>::max_(scale, Real<T>::max_(Real<T>::max_(Real<T>::abs_(a01), Real<T>::abs_(a02)), Real<T>::abs_(a12)));
    if (!(scale > T(0))) scale = T(1);
    const T inv = T(1) / scale;
    a00 *= inv; a01 *= inv; a02 *= inv; a11 *= inv; a12 *= inv; a22 *= inv;
    T V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
#pragma unroll 1
    for (int sweep = 0; sweep < Real<T>::kSweeps; ++sweep) {
        const T off = Real<T>::abs_(a01) + Real<T>::abs_(a02) + Real<T>::abs_(a12);
        const T dia = Real<T>::abs_(a00) + Real<T>::abs_(a11) + Real<T>::abs_(a22);
        if (off <= Real<T>::kOffTol * dia) break;
        jacobi_rotate(a00, a11, a01, a02, a12, V, 0, 1);
        jacobi_rotate(a00, a22, a02, a01, a12, V, 0, 2);
        jacobi_rotate(a11, a22, a12, a01, a02, V, 1, 2);
    }
    // stable 3-element sort of (eigenvalue, eigenvector) pairs with register swaps only
    T w0 = a00 * scale, w1 = a11 * scale, w2 = a22 * scale;
    T c0[3] = {V[0][0], V[1][0], V[2][0]}, c1[3] = {V[0][1], V[1][1], V[2][1]}, c2[3] = {V[0][2], V[1][2], V[2][2]};
    const bool inc = (eig_order == PGEOF_EIG_LITERAL);
#define PGEOF_CSWAP(wa, ca, wb, cb)                                               \
    {                                                                             \
        const bool sw = inc ? (wb < wa) : (wb > wa);                              \
        const T tw = sw ? wb : wa; wb = sw ? wa : wb; wa = tw;                    \
        _Pragma("unroll") for (int i = 0; i < 3; ++i) { const T tc = sw ? cb[i] : ca[i]; cb[i] = sw ? ca[i] : cb[i]; ca[i] = tc; } \
    }
    PGEOF_CSWAP(w0, c0, w1, c1)
    PGEOF_CSWAP(w1, c1, w2, c2)
    PGEOF_CSWAP(w0, c0, w1, c1)
#undef PGEOF_CSWAP
    Pca<T> r;
    r.val[0] = Real<T>::max_(w0, T(0));
    r.val[1] = Real<T>::max_(w1, T(0));
    r.val[2] = Real<T>::max_(w2, T(0));
#pragma unroll
    for (int i = 0; i < 3; ++i) { r.v0[i] = c0[i]; r.v1[i] = c1[i]; r.v2[i] = c2[i]; }
    if (r.v2[2] < T(0)) { r.v2[0] = -r.v2[0]; r.v2[1] = -r.v2[1]; r.v2[2] = -r.v2[2]; }
    return r;
}

// Eigenvalues only, float, cyclic Jacobi without the eigenvector accumulation: backward stable, every eigenvalue is
// within a few ulp of the matrix NORM (also when eigenvalues coincide or vanish, where the closed-form trigonometric
// solution loses half the digits).  Used as the FILTER of the optimal-k scan; unsorted, clamped at 0.
__device__ __forceinline__ void jacobi_eigvals_f32(float a00, float a01, float a02, float a11, float a12, float a22, float (&w)[3])
{
    float scale = fmaxf(fmaxf(fabsf(a00), fabsf(a11)), fabsf(a22));
    scale = fmaxf(scale, fmaxf(fmaxf(fabsf(a01), fabsf(a02)), fabsf(a12)));
    if (!(scale > 0.f)) { w[0] = w[1] = w[2] = 0.f; return; }
    const float inv = 1.f / scale;
    a00 *= inv; a01 *= inv; a02 *= inv; a11 *= inv; a12 *= inv; a22 *= inv;
#define PGEOF_ROT(app, aqq, apq, arp, arq)                                            \
    if (apq != 0.f) {                                                                 \
        const float theta = (aqq - app) / (2.f * apq);                                \
        float t = 1.f / (fabsf(theta) + sqrtf(theta * theta + 1.f));                  \
        if (theta < 0.f) t = -t;                                                      \
        const float c = rsqrtf(t * t + 1.f), sn = t * c;                              \
        app -= t * apq; aqq += t * apq; apq = 0.f;                                    \
        const float rp = arp, rq = arq;                                               \
        arp = c * rp - sn * rq; arq = sn * rp + c * rq;                               \
    }
#pragma unroll 1
    for (int sweep = 0; sweep < 6; ++sweep) {
        const float off = fabsf(a01) + fabsf(a02) + fabsf(a12);
        if (off <= 1e-8f * (fabsf(a00) + fabsf(a11) + fabsf(a22))) break;
        PGEOF_ROT(a00, a11, a01, a02, a12)
        PGEOF_ROT(a00, a22, a02, a01, a12)
        PGEOF_ROT(a11, a22, a12, a01, a02)
    }
#undef PGEOF_ROT
    w[0] = fmaxf(a00 * scale, 0.f); w[1] = fmaxf(a11 * scale, 0.f); w[2] = fmaxf(a22 * scale, 0.f);
}

// origin-shifted first and second moments of a neighbourhood (shared by the feature kernels and the fused knn_features path)
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

// include/pca.hpp:140-150
template <typename T>
__device__ __forceinline__ T eigentropy_of(T l0, T l1, T l2)
{
    const T eps = T(1e-3);
    const T s = l0 + l1 + l2 + eps;
    const T e0 = l0 / s, e1 = l1 / s, e2 = l2 / s;
    return -e0 * Real<T>::log_(e0 + eps) - e1 * Real<T>::log_(e1 + eps) - e2 * Real<T>::log_(e2 + eps);
}

template <typename T>
__device__ __forceinline__ T verticality_pgeof(const Pca<T>& p)
{
    T u[3];
#pragma unroll
    for (int d = 0; d < 3; ++d)
        u[d] = p.val[0] * Real<T>::abs_(p.v0[d]) + p.val[1] * Real<T>::abs_(p.v1[d]) + p.val[2] * Real<T>::abs_(p.v2[d]);
    return u[2] / Real<T>::sqrt_(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
}

// The 11 features of include/pca.hpp:160-200 in EFeatureID order.
template <typename T>
__device__ __forceinline__ void features11(const Pca<T>& p, T (&f)[11])
{
    const T eps = T(1e-3);
    const T s0 = Real<T>::sqrt_(p.val[0]), s1 = Real<T>::sqrt_(p.val[1]), s2 = Real<T>::sqrt_(p.val[2]);
    const T fact = T(1) / (s0 + eps);
    f[PGEOF_LINEARITY] = (s0 - s1) * fact;
    f[PGEOF_PLANARITY] = (s1 - s2) * fact;
    f[PGEOF_SCATTERING] = s2 * fact;
    f[PGEOF_VERTICALITY_PGEOF] = s0 > T(0) ? verticality_pgeof(p) : T(0);
    f[PGEOF_NORMAL_X] = p.v2[0];
    f[PGEOF_NORMAL_Y] = p.v2[1];
    f[PGEOF_NORMAL_Z] = p.v2[2];
    f[PGEOF_LENGTH] = s0;
    f[PGEOF_SURFACE] = Real<T>::sqrt_(s0 * s1 + T(1e-6));
    f[PGEOF_VOLUME] = Real<T>::cbrt_(s0 * s1 * s2 + T(1e-9));
    f[PGEOF_CURVATURE] = s2 / (s0 + s1 + s2 + eps);
}

// One feature by id, include/pca.hpp:212-295 (K_optimal has no case -> 0).
template <typename T>
__device__ __forceinline__ T feature_selected(const Pca<T>& p, int id)
{
    const T eps = T(1e-3);
    const T s0 = Real<T>::sqrt_(p.val[0]), s1 = Real<T>::sqrt_(p.val[1]), s2 = Real<T>::sqrt_(p.val[2]);
    const T fact = T(1) / (s0 + eps);
    switch (id) {
        case PGEOF_LINEARITY: return (s0 - s1) * fact;
        case PGEOF_PLANARITY: return (s1 - s2) * fact;
        case PGEOF_SCATTERING: return s2 * fact;
        case PGEOF_VERTICALITY_PGEOF: return s0 > T(0) ? verticality_pgeof(p) : T(0);
        case PGEOF_NORMAL_X: return p.v2[0];
        case PGEOF_NORMAL_Y: return p.v2[1];
        case PGEOF_NORMAL_Z: return p.v2[2];
        case PGEOF_LENGTH: return s0;
        case PGEOF_SURFACE: return Real<T>::sqrt_(s0 * s1 + T(1e-6f));   // float literal even for double, pca.hpp:252
        case PGEOF_VOLUME: return Real<T>::cbrt_(s0 * s1 * s2 + T(1e-9));
        case PGEOF_CURVATURE: return s2 / (s0 + s1 + s2 + eps);
        case PGEOF_VERTICALITY: return T(1) - Real<T>::abs_(p.v2[2]);
        case PGEOF_EIGENTROPY: return eigentropy_of(p.val[0], p.val[1], p.val[2]);
        default: return T(0);
    }
}

// Eigenvalues only (closed form, Smith 1961), clamped at 0: used by the optimal-k scan
// where up to ~100 neighbourhood sizes are evaluated per point.
__device__ __forceinline__ void eigvals3_f64(double a00, double a01, double a02, double a11, double a12, double a22, double (&w)[3])
{
    const double q = (a00 + a11 + a22) / 3.0;
    const double b00 = a00 - q, b11 = a11 - q, b22 = a22 - q;
    const double p2 = b00 * b00 + b11 * b11 + b22 * b22 + 2.0 * (a01 * a01 + a02 * a02 + a12 * a12);
    if (!(p2 > 0.0)) { w[0] = w[1] = w[2] = fmax(q, 0.0); return; }
    const double p = sqrt(p2 / 6.0);
    const double ip = 1.0 / p;
    const double c00 = b00 * ip, c11 = b11 * ip, c22 = b22 * ip, c01 = a01 * ip, c02 = a02 * ip, c12 = a12 * ip;
    double r = 0.5 * (c00 * (c11 * c22 - c12 * c12) - c01 * (c01 * c22 - c12 * c02) + c02 * (c01 * c12 - c11 * c02));
    r = fmin(1.0, fmax(-1.0, r));
    const double phi = acos(r) / 3.0;
    const double hi = q + 2.0 * p * cos(phi);
    const double lo = q + 2.0 * p * cos(phi + 2.0943951023931954923);   // + 2 pi / 3
    const double mid = 3.0 * q - hi - lo;
    w[0] = fmax(lo, 0.0); w[1] = fmax(mid, 0.0); w[2] = fmax(hi, 0.0);
}

}  // namespace pgeof
