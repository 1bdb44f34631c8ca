// cpu_ref.cpp -- C++ restatement of the pgeof hot path.  TEST INFRASTRUCTURE ONLY.
//
// This is the CPU oracle / CPU baseline ("port") for the CUDA path.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load it; the product (point_geometric_features_b200/) never does.
//
// Why a port and not the reference binary: the reference needs Eigen 3.4.0,
// nanoflann (@9c930ba4) and Taskflow (@9d9cea03) whose submodule directories are
// empty in /root/reference (.SUBMODULES.json:8-30) and nanobind is absent, so it
// is unbuildable here.  Parity status: neighbour indices are pinned against
// scipy.spatial.KDTree (the reference's own test oracle, tests/test_pgeof.py:8-27);
// every feature value is "parity unpinned" (the reference holds no golden values).
//
// What follows which reference lines (all paths relative to /root/reference):
//   kd-tree search, leaf size 10, sorted-k result set  include/nn_search.hpp:31-67
//   radius search, strict d2 < r*r, -1 / 0 padding     include/nn_search.hpp:85-132
//   mean / centred covariance / eigen / clamp / flip   include/pca.hpp:71-98
//   eigenentropy                                       include/pca.hpp:140-150
//   11 features                                        include/pca.hpp:160-200
//   selected features                                  include/pca.hpp:212-295
//   drivers (k_min gate, multiscale, optimal, selected) include/pgeof.hpp:75-375
// Third-party arithmetic restated from the published algorithms:
//   Eigen 3.4.0 SelfAdjointEigenSolver::compute -> scale by max|a_ij|, Householder
//   tridiagonalisation, implicit symmetric QR with Wilkinson shift (Golub & Van Loan
//   alg. 8.3.2/8.3.3), eigenvalues sorted increasing.
//   nanoflann metric_L2_Simple -> sequential float accumulation of diff*diff;
//   KNNResultSet / RKNNResultSet -> ascending insertion; here ties are ordered by
//   index so that results are a pure function of the input.
//   Taskflow for_each_index + StaticPartitioner -> std::thread, contiguous chunks.
//
// Build: g++ -O3 -std=c++17 -ffp-contract=off -fPIC -shared -pthread (see oracle/Makefile)

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

template <typename F>
void parallel_for(size_t n, int nthreads, F&& fn)
{
    if (nthreads <= 0) nthreads = (int)std::max(1u, std::thread::hardware_concurrency());
    if ((size_t)nthreads > n) nthreads = (int)std::max<size_t>(1, n);
    if (nthreads == 1) { fn(size_t(0), n); return; }
    std::vector<std::thread> pool;
    const size_t chunk = (n + nthreads - 1) / nthreads;  // StaticPartitioner(0): one contiguous chunk per worker
    for (int t = 0; t < nthreads; ++t) {
        const size_t a = std::min(n, chunk * t), b = std::min(n, a + chunk);
        if (a < b) pool.emplace_back([=, &fn] { fn(a, b); });
    }
    for (auto& th : pool) th.join();
}

// ---------------------------------------------------------------------------
// Defined float32 metric (nanoflann L2_Simple, nn_search.hpp:35).
// ---------------------------------------------------------------------------
template <typename T>
inline T sqdist(const T* a, const T* b)
{
    const T dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    T r = dx * dx;
    r = r + dy * dy;
    r = r + dz * dz;
    return r;
}

template <typename T>
struct Hit { T d2; uint32_t idx; };

template <typename T>
inline bool hit_less(const Hit<T>& a, const Hit<T>& b)
{
    return a.d2 < b.d2 || (a.d2 == b.d2 && a.idx < b.idx);
}

// Fixed-capacity ascending result set (KNNResultSet / RKNNResultSet semantics,
// ordered by (d2, idx)).
template <typename T>
struct ResultSet {
    Hit<T>* items; uint32_t cap; uint32_t count = 0; T radius2; bool strict_radius;
    ResultSet(Hit<T>* buf, uint32_t capacity, T r2, bool strict) : items(buf), cap(capacity), radius2(r2), strict_radius(strict) {}
    inline bool full() const { return count == cap; }
    // bound a subtree must not exceed to be worth visiting
    inline T worst() const { return full() ? items[cap - 1].d2 : radius2; }
    inline void add(T d2, uint32_t idx)
    {
        if (strict_radius && !(d2 < radius2)) return;
        Hit<T> h{d2, idx};
        if (full() && !hit_less(h, items[cap - 1])) return;
        uint32_t i = full() ? cap - 1 : count++;
        while (i > 0 && hit_less(h, items[i - 1])) { items[i] = items[i - 1]; --i; }
        items[i] = h;
    }
};

// ---------------------------------------------------------------------------
// KD-tree, leaf size 10 (nn_search.hpp:39).
// ---------------------------------------------------------------------------
template <typename T>
struct KDTree {
    struct Node { int32_t left, right; uint32_t lo, hi; int dim; T split_lo, split_hi; };
    const T* pts; size_t n; std::vector<uint32_t> order; std::vector<Node> nodes; T bb_lo[3], bb_hi[3];
    static constexpr uint32_t kLeaf = 10;

    KDTree(const T* p, size_t n_) : pts(p), n(n_), order(n_)
    {
        for (size_t i = 0; i < n; ++i) order[i] = (uint32_t)i;
        for (int d = 0; d < 3; ++d) { bb_lo[d] = std::numeric_limits<T>::max(); bb_hi[d] = std::numeric_limits<T>::lowest(); }
        for (size_t i = 0; i < n; ++i)
            for (int d = 0; d < 3; ++d) { bb_lo[d] = std::min(bb_lo[d], pts[3 * i + d]); bb_hi[d] = std::max(bb_hi[d], pts[3 * i + d]); }
        nodes.reserve(2 * (n / kLeaf + 1));
        if (n > 0) build(0, (uint32_t)n);
    }

    int32_t build(uint32_t lo, uint32_t hi)
    {
        const int32_t id = (int32_t)nodes.size();
        nodes.push_back(Node{-1, -1, lo, hi, 0, 0, 0});
        if (hi - lo <= kLeaf) return id;
        T mn[3], mx[3];
        for (int d = 0; d < 3; ++d) { mn[d] = std::numeric_limits<T>::max(); mx[d] = std::numeric_limits<T>::lowest(); }
        for (uint32_t i = lo; i < hi; ++i)
            for (int d = 0; d < 3; ++d) { const T v = pts[3 * (size_t)order[i] + d]; mn[d] = std::min(mn[d], v); mx[d] = std::max(mx[d], v); }
        int dim = 0;
        for (int d = 1; d < 3; ++d) if (mx[d] - mn[d] > mx[dim] - mn[dim]) dim = d;
        if (!(mx[dim] > mn[dim])) return id;  // all coincident: keep as one (big) leaf
        const uint32_t mid = lo + (hi - lo) / 2;
        std::nth_element(order.begin() + lo, order.begin() + mid, order.begin() + hi,
                         [&](uint32_t a, uint32_t b) { return pts[3 * (size_t)a + dim] < pts[3 * (size_t)b + dim]; });
        T left_max = std::numeric_limits<T>::lowest(), right_min = std::numeric_limits<T>::max();
        for (uint32_t i = lo; i < mid; ++i) left_max = std::max(left_max, pts[3 * (size_t)order[i] + dim]);
        for (uint32_t i = mid; i < hi; ++i) right_min = std::min(right_min, pts[3 * (size_t)order[i] + dim]);
        const int32_t l = build(lo, mid);
        const int32_t r = build(mid, hi);
        Node& nd = nodes[id];
        nd.left = l; nd.right = r; nd.dim = dim; nd.split_lo = left_max; nd.split_hi = right_min;
        return id;
    }

    // Lower bounds are evaluated in double and deflated by 1e-6 so that float
    // rounding of the metric can never prune a point that belongs to the result.
    void search(int32_t id, const T* q, double mind2, double* off, ResultSet<T>& rs) const
    {
        const Node& nd = nodes[id];
        if (nd.left < 0) {
            for (uint32_t i = nd.lo; i < nd.hi; ++i) { const uint32_t j = order[i]; rs.add(sqdist(q, pts + 3 * (size_t)j), j); }
            return;
        }
        const int d = nd.dim;
        const double qd = (double)q[d];
        const double dl = qd - (double)nd.split_lo;   // > 0 when q is right of the left child's extent
        const double dr = (double)nd.split_hi - qd;   // > 0 when q is left of the right child's extent
        int32_t near, far; double cut;
        if (dl <= dr) { near = nd.left; far = nd.right; cut = std::max(0.0, dr); }
        else { near = nd.right; far = nd.left; cut = std::max(0.0, dl); }
        search(near, q, mind2, off, rs);
        const double saved = off[d];
        const double far2 = mind2 - saved * saved + cut * cut;
        if (far2 * (1.0 - 1e-6) <= (double)rs.worst()) {
            off[d] = cut;
            search(far, q, far2, off, rs);
            off[d] = saved;
        }
    }

    void query(const T* q, ResultSet<T>& rs) const
    {
        if (n == 0) return;
        double off[3], mind2 = 0;
        for (int d = 0; d < 3; ++d) {
            off[d] = 0;
            if (q[d] < bb_lo[d]) off[d] = (double)bb_lo[d] - (double)q[d];
            if (q[d] > bb_hi[d]) off[d] = (double)q[d] - (double)bb_hi[d];
            mind2 += off[d] * off[d];
        }
        search(0, q, mind2, off, rs);
    }
};

// ---------------------------------------------------------------------------
// Symmetric 3x3 eigen-decomposition, Eigen::SelfAdjointEigenSolver::compute
// restated (pca.hpp:79): scale, tridiagonalise, implicit QR, sort increasing.
// ---------------------------------------------------------------------------
template <typename T>
void eig3_sym(const T a_in[3][3], T w[3], T V[3][3])
{
    T a[3][3];
    T scale = 0;
    for (int i = 0; i < 3; ++i) for (int j = 0; j <= i; ++j) scale = std::max(scale, std::abs(a_in[i][j]));
    if (scale == T(0)) scale = T(1);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) a[i][j] = a_in[std::max(i, j)][std::min(i, j)] / scale;

    // Householder: reflect (a10, a20) onto (beta, 0).  Q = diag(1, H), H = [[c, s], [s, -c]].
    T d[3], e[2];
    T Q[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    d[0] = a[0][0];
    const T tiny = std::numeric_limits<T>::min();
    if (a[2][0] * a[2][0] <= tiny) {
        d[1] = a[1][1]; d[2] = a[2][2]; e[0] = a[1][0]; e[1] = a[2][1];
    } else {
        const T beta = std::sqrt(a[1][0] * a[1][0] + a[2][0] * a[2][0]);
        const T c = a[1][0] / beta, s = a[2][0] / beta;
        // H * A22 * H with A22 = [[a11, a21], [a21, a22]]
        const T t = T(2) * c * a[2][1] + s * (a[2][2] - a[1][1]);
        d[1] = a[1][1] + s * t;
        d[2] = a[2][2] - s * t;
        e[0] = beta;
        e[1] = a[2][1] - c * t;
        Q[1][1] = c; Q[1][2] = s; Q[2][1] = s; Q[2][2] = -c;
    }

    // implicit symmetric QR iterations on the unreduced trailing block
    const T eps = std::numeric_limits<T>::epsilon();
    int end = 2, iter = 0;
    while (end > 0) {
        for (int i = 0; i < end; ++i) {
            if (std::abs(e[i]) < tiny) e[i] = 0;
            else { const T se = e[i] / eps; if (se * se <= std::abs(d[i]) + std::abs(d[i + 1])) e[i] = 0; }
        }
        while (end > 0 && e[end - 1] == T(0)) --end;
        if (end <= 0) break;
        if (++iter > 30 * 3) break;
        int start = end - 1;
        while (start > 0 && e[start - 1] != T(0)) --start;
        // Wilkinson shift from the trailing 2x2
        const T td = (d[end - 1] - d[end]) * T(0.5);
        const T ee = e[end - 1];
        T mu = d[end];
        if (td == T(0)) mu -= std::abs(ee);
        else if (ee != T(0)) {
            const T h = std::hypot(td, ee);
            const T e2 = ee * ee;
            const T den = td + (td > T(0) ? h : -h);
            mu -= (e2 == T(0)) ? ee / (den / ee) : e2 / den;
        }
        T x = d[start] - mu, z = e[start];
        for (int k = start; k < end && z != T(0); ++k) {
            // Givens rotation G with G^T [x z]^T = [r 0]^T
            T c, s;
            {
                const T r = std::hypot(x, z);
                c = x / r; s = -z / r;
            }
            const T sdk = s * d[k] + c * e[k];
            const T dkp1 = s * e[k] + c * d[k + 1];
            d[k] = c * (c * d[k] - s * e[k]) - s * (c * e[k] - s * d[k + 1]);
            d[k + 1] = s * sdk + c * dkp1;
            e[k] = c * sdk - s * dkp1;
            if (k > start) e[k - 1] = c * e[k - 1] - s * z;
            x = e[k];
            if (k < end - 1) { z = -s * e[k + 1]; e[k + 1] = c * e[k + 1]; }
            for (int r = 0; r < 3; ++r) {   // Q <- Q * G on columns k, k+1
                const T qk = Q[r][k], qk1 = Q[r][k + 1];
                Q[r][k] = c * qk - s * qk1;
                Q[r][k + 1] = s * qk + c * qk1;
            }
        }
    }
    // increasing order (selection sort, as Eigen does)
    for (int i = 0; i < 2; ++i) {
        int k = i;
        for (int j = i + 1; j < 3; ++j) if (d[j] < d[k]) k = j;
        if (k != i) { std::swap(d[i], d[k]); for (int r = 0; r < 3; ++r) std::swap(Q[r][i], Q[r][k]); }
    }
    for (int i = 0; i < 3; ++i) { w[i] = d[i] * scale; for (int r = 0; r < 3; ++r) V[r][i] = Q[r][i]; }
}

template <typename T>
struct PCA { T val[3]; T v0[3], v1[3], v2[3]; };

// pca.hpp:71-98 on a gathered (k,3) cloud.  eig_order: 0 literal (increasing,
// what the snapshot + Eigen 3.4.0 yields), 1 documented (decreasing).
template <typename T>
PCA<T> pca_from_cloud(const T* cloud, size_t k, int eig_order)
{
    T mean[3] = {0, 0, 0};
    for (size_t i = 0; i < k; ++i) for (int d = 0; d < 3; ++d) mean[d] += cloud[3 * i + d];
    for (int d = 0; d < 3; ++d) mean[d] /= T(k);
    T cov[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (size_t i = 0; i < k; ++i) {
        const T c0 = cloud[3 * i] - mean[0], c1 = cloud[3 * i + 1] - mean[1], c2 = cloud[3 * i + 2] - mean[2];
        cov[0][0] += c0 * c0; cov[1][0] += c1 * c0; cov[1][1] += c1 * c1;
        cov[2][0] += c2 * c0; cov[2][1] += c2 * c1; cov[2][2] += c2 * c2;
    }
    for (int i = 0; i < 3; ++i) for (int j = 0; j <= i; ++j) { cov[i][j] /= T(k); cov[j][i] = cov[i][j]; }
    T w[3], V[3][3];
    eig3_sym(cov, w, V);
    PCA<T> p;
    int perm[3] = {0, 1, 2};
    if (eig_order != 0)   // documented: decreasing, stable (ties keep the solver's column order)
        std::stable_sort(perm, perm + 3, [&](int a, int b) { return w[a] > w[b]; });
    for (int s = 0; s < 3; ++s) {
        const int c = perm[s];
        p.val[s] = std::max(w[c], T(0));
        T* dst = s == 0 ? p.v0 : (s == 1 ? p.v1 : p.v2);
        for (int r = 0; r < 3; ++r) dst[r] = V[r][c];
    }
    if (p.v2[2] < T(0)) for (int r = 0; r < 3; ++r) p.v2[r] = -p.v2[r];
    return p;
}

template <typename T, typename S>
PCA<T> pca_from_row(const S* xyz, const uint32_t* nn, size_t k, int eig_order, std::vector<T>& scratch)
{
    scratch.resize(3 * k);
    for (size_t j = 0; j < k; ++j) for (int d = 0; d < 3; ++d) scratch[3 * j + d] = (T)xyz[3 * (size_t)nn[j] + d];
    return pca_from_cloud<T>(scratch.data(), k, eig_order);
}

template <typename T>
inline T eigentropy(const PCA<T>& p)
{
    const T eps = T(1e-3);
    const T s = p.val[0] + p.val[1] + p.val[2] + eps;
    const T e0 = p.val[0] / s, e1 = p.val[1] / s, e2 = p.val[2] / s;
    return -e0 * std::log(e0 + eps) - e1 * std::log(e1 + eps) - e2 * std::log(e2 + eps);
}

template <typename T>
inline T verticality_pgeof(const PCA<T>& p)
{
    T u[3];
    for (int d = 0; d < 3; ++d) u[d] = p.val[0] * std::abs(p.v0[d]) + p.val[1] * std::abs(p.v1[d]) + p.val[2] * std::abs(p.v2[d]);
    return u[2] / std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
}

// One feature by EFeatureID (pca.hpp:47-63, :160-200, :212-295).  `surface_eps`
// is 1e-6 in compute_features and the float literal 1e-6f in the selected path.
template <typename T>
inline T feature_by_id(const PCA<T>& p, int id, bool selected_path)
{
    const T eps = T(1e-3);
    const T s0 = std::sqrt(p.val[0]), s1 = std::sqrt(p.val[1]), s2 = std::sqrt(p.val[2]);
    const T fact = T(1) / (s0 + eps);
    switch (id) {
        case 0: return (s0 - s1) * fact;
        case 1: return (s1 - s2) * fact;
        case 2: return s2 * fact;
        case 3: return s0 > T(0) ? verticality_pgeof(p) : T(0);
        case 4: return p.v2[0];
        case 5: return p.v2[1];
        case 6: return p.v2[2];
        case 7: return s0;
        case 8: return selected_path ? (T)std::sqrt(s0 * s1 + 1e-6f) : std::sqrt(s0 * s1 + T(1e-6));
        case 9: return std::pow(s0 * s1 * s2 + T(1e-9), T(1) / T(3));
        case 10: return s2 / (s0 + s1 + s2 + eps);
        case 12: return T(1) - std::abs(p.v2[2]);
        case 13: return eigentropy(p);
        default: return T(0);   // K_optimal (11): no case in the reference
    }
}

template <typename T, typename O>
inline void write_features11(const PCA<T>& p, O* out)
{
    for (int f = 0; f < 11; ++f) out[f] = (O)feature_by_id(p, f, false);
}

// --------------------------- drivers ---------------------------------------
template <typename T, typename O>
void features_driver(const float* xyz, const uint32_t* nn, const uint32_t* nn_ptr, size_t n_rows, uint32_t k_min,
                     int eig_order, O* out, int nthreads)
{
    std::memset(out, 0, sizeof(O) * n_rows * 11);
    parallel_for(n_rows, nthreads, [&](size_t a, size_t b) {
        std::vector<T> scratch;
        for (size_t i = a; i < b; ++i) {
            const size_t k = nn_ptr[i + 1] - nn_ptr[i];
            if (k >= k_min) write_features11(pca_from_row<T>(xyz, nn + nn_ptr[i], k, eig_order, scratch), out + 11 * i);
        }
    });
}

template <typename T, typename O>
void multiscale_driver(const float* xyz, const uint32_t* nn, const uint32_t* nn_ptr, size_t n_rows, const uint32_t* scales,
                       size_t n_scales, int eig_order, O* out, int nthreads)
{
    std::memset(out, 0, sizeof(O) * n_rows * n_scales * 11);
    parallel_for(n_rows, nthreads, [&](size_t a, size_t b) {
        std::vector<T> scratch;
        for (size_t i = a; i < b; ++i) {
            const size_t k = nn_ptr[i + 1] - nn_ptr[i];
            for (size_t s = 0; s < n_scales; ++s) {
                if (k < scales[s]) break;
                write_features11(pca_from_row<T>(xyz, nn + nn_ptr[i], scales[s], eig_order, scratch), out + (i * n_scales + s) * 11);
            }
        }
    });
}

template <typename T, typename O>
void optimal_driver(const float* xyz, const uint32_t* nn, const uint32_t* nn_ptr, size_t n_rows, uint32_t k_min, uint32_t k_step,
                    uint32_t k_min_search, int eig_order, O* out, O* margin, int nthreads)
{
    std::memset(out, 0, sizeof(O) * n_rows * 12);
    parallel_for(n_rows, nthreads, [&](size_t a, size_t b) {
        std::vector<T> scratch;
        for (size_t i = a; i < b; ++i) {
            const size_t k_nn = nn_ptr[i + 1] - nn_ptr[i];
            if (margin) margin[i] = std::numeric_limits<O>::infinity();
            if (k_nn >= k_min && k_nn >= k_min_search) {
                const size_t k0 = std::min<size_t>(std::max<size_t>(k_min, k_min_search), k_nn);
                PCA<T> best{}; T best_h = T(1), second = std::numeric_limits<T>::infinity(); size_t best_k = k_nn;
                for (size_t k = k0; k <= k_nn; ++k) {
                    if (k > k0 && (k % k_step != 0) && k != k_nn) continue;
                    const PCA<T> p = pca_from_row<T>(xyz, nn + nn_ptr[i], k, eig_order, scratch);
                    const T h = eigentropy(p);
                    if (k == k0 || h < best_h) { if (k != k0) second = best_h; best_h = h; best_k = k; best = p; }
                    else second = std::min(second, h);
                }
                write_features11(best, out + 12 * i);
                out[12 * i + 11] = (O)best_k;
                if (margin) margin[i] = (O)(second - best_h);
            }
        }
    });
}

template <typename T>
void knn_driver(const T* data, size_t nd, const T* query, size_t nq, uint32_t k, uint32_t* idx, T* d2, int nthreads)
{
    KDTree<T> tree(data, nd);
    parallel_for(nq, nthreads, [&](size_t a, size_t b) {
        std::vector<Hit<T>> buf(k);
        for (size_t i = a; i < b; ++i) {
            ResultSet<T> rs(buf.data(), k, std::numeric_limits<T>::infinity(), false);
            tree.query(query + 3 * i, rs);
            for (uint32_t j = 0; j < k; ++j) { idx[i * k + j] = buf[j].idx; d2[i * k + j] = buf[j].d2; }
        }
    });
}

template <typename T>
void radius_driver(const T* data, size_t nd, const T* query, size_t nq, T radius, uint32_t max_knn, int32_t* idx, T* d2, int nthreads)
{
    KDTree<T> tree(data, nd);
    const T r2 = radius * radius;   // nn_search.hpp:98
    parallel_for(nq, nthreads, [&](size_t a, size_t b) {
        std::vector<Hit<T>> buf(std::max<uint32_t>(max_knn, 1));
        for (size_t i = a; i < b; ++i) {
            ResultSet<T> rs(buf.data(), max_knn, r2, true);
            if (max_knn > 0) tree.query(query + 3 * i, rs);
            for (uint32_t j = 0; j < max_knn; ++j) {
                idx[i * max_knn + j] = j < rs.count ? (int32_t)buf[j].idx : -1;
                d2[i * max_knn + j] = j < rs.count ? buf[j].d2 : T(0);
            }
        }
    });
}

template <typename T>
void selected_driver(const T* xyz, size_t n, T radius, uint32_t max_knn, const int32_t* ids, size_t n_ids, int eig_order, T* out,
                     int nthreads)
{
    KDTree<T> tree(xyz, n);
    const T r2 = radius * radius;
    std::memset(out, 0, sizeof(T) * n * n_ids);
    // The reference collects every point with d2 < r2 and partial-sorts to the max_knn
    // nearest (pgeof.hpp:348-364): the same set as a bounded ascending result set.
    parallel_for(n, nthreads, [&](size_t a, size_t b) {
        std::vector<Hit<T>> buf(std::max<uint32_t>(max_knn, 1));
        std::vector<T> cloud;
        for (size_t i = a; i < b; ++i) {
            ResultSet<T> rs(buf.data(), max_knn, r2, true);
            if (max_knn > 0) tree.query(xyz + 3 * i, rs);
            // `num_found < 2` (pgeof.hpp:355) counts all points in the ball; with max_knn >= 2
            // that equals rs.count < 2.  max_knn < 2 keeps <= 1 point: PCA of one point.
            size_t found = rs.count;
            if (max_knn < 2) {   // count the ball without the cap
                Hit<T> two[2]; ResultSet<T> probe(two, 2, r2, true); tree.query(xyz + 3 * i, probe); found = probe.count;
            }
            if (found < 2) continue;
            const size_t k = rs.count;
            if (k == 0) continue;   // max_knn == 0: Eigen would divide by zero rows; leave zeros
            cloud.resize(3 * k);
            for (size_t j = 0; j < k; ++j) for (int d = 0; d < 3; ++d) cloud[3 * j + d] = xyz[3 * (size_t)buf[j].idx + d];
            const PCA<T> p = pca_from_cloud<T>(cloud.data(), k, eig_order);
            for (size_t f = 0; f < n_ids; ++f) out[i * n_ids + f] = feature_by_id(p, ids[f], true);
        }
    });
}

}  // namespace

// ---------------------------------------------------------------------------
// C entry points (ctypes).  `_f32` = the reference's arithmetic (float), the timed
// CPU baseline; `_f64` = same algorithm evaluated in double from the float inputs,
// the high-precision side of the feature tolerance check.
// ---------------------------------------------------------------------------
extern "C" {

int oracle_hardware_threads(void) { return (int)std::max(1u, std::thread::hardware_concurrency()); }

void oracle_knn_f32(const float* data, size_t nd, const float* query, size_t nq, uint32_t k, uint32_t* idx, float* d2, int nthreads)
{ knn_driver<float>(data, nd, query, nq, k, idx, d2, nthreads); }

void oracle_knn_brute_f32(const float* data, size_t nd, const float* query, size_t nq, uint32_t k, uint32_t* idx, float* d2, int nthreads)
{
    parallel_for(nq, nthreads, [&](size_t a, size_t b) {
        std::vector<Hit<float>> buf(k);
        for (size_t i = a; i < b; ++i) {
            ResultSet<float> rs(buf.data(), k, std::numeric_limits<float>::infinity(), false);
            for (size_t j = 0; j < nd; ++j) rs.add(sqdist(query + 3 * i, data + 3 * j), (uint32_t)j);
            for (uint32_t j = 0; j < k; ++j) { idx[i * k + j] = buf[j].idx; d2[i * k + j] = buf[j].d2; }
        }
    });
}

void oracle_radius_f32(const float* data, size_t nd, const float* query, size_t nq, float r, uint32_t max_knn, int32_t* idx, float* d2, int nthreads)
{ radius_driver<float>(data, nd, query, nq, r, max_knn, idx, d2, nthreads); }

void oracle_features_f32(const float* xyz, const uint32_t* nn, const uint32_t* nn_ptr, size_t n_rows, uint32_t k_min, int eig_order, float* out, int nthreads)
{ features_driver<float, float>(xyz, nn, nn_ptr, n_rows, k_min, eig_order, out, nthreads); }
void oracle_features_f64(const float* xyz, const uint32_t* nn, const uint32_t* nn_ptr, size_t n_rows, uint32_t k_min, int eig_order, double* out, int nthreads)
{ features_driver<double, double>(xyz, nn, nn_ptr, n_rows, k_min, eig_order, out, nthreads); }

void oracle_multiscale_f32(const float* xyz, const uint32_t* nn, const uint32_t* nn_ptr, size_t n_rows, const uint32_t* scales, size_t n_scales, int eig_order, float* out, int nthreads)
{ multiscale_driver<float, float>(xyz, nn, nn_ptr, n_rows, scales, n_scales, eig_order, out, nthreads); }
void oracle_multiscale_f64(const float* xyz, const uint32_t* nn, const uint32_t* nn_ptr, size_t n_rows, const uint32_t* scales, size_t n_scales, int eig_order, double* out, int nthreads)
{ multiscale_driver<double, double>(xyz, nn, nn_ptr, n_rows, scales, n_scales, eig_order, out, nthreads); }

void oracle_optimal_f32(const float* xyz, const uint32_t* nn, const uint32_t* nn_ptr, size_t n_rows, uint32_t k_min, uint32_t k_step, uint32_t k_min_search, int eig_order, float* out, float* margin, int nthreads)
{ optimal_driver<float, float>(xyz, nn, nn_ptr, n_rows, k_min, k_step, k_min_search, eig_order, out, margin, nthreads); }
void oracle_optimal_f64(const float* xyz, const uint32_t* nn, const uint32_t* nn_ptr, size_t n_rows, uint32_t k_min, uint32_t k_step, uint32_t k_min_search, int eig_order, double* out, double* margin, int nthreads)
{ optimal_driver<double, double>(xyz, nn, nn_ptr, n_rows, k_min, k_step, k_min_search, eig_order, out, margin, nthreads); }

void oracle_selected_f32(const float* xyz, size_t n, float r, uint32_t max_knn, const int32_t* ids, size_t n_ids, int eig_order, float* out, int nthreads)
{ selected_driver<float>(xyz, n, r, max_knn, ids, n_ids, eig_order, out, nthreads); }
void oracle_selected_f64(const double* xyz, size_t n, double r, uint32_t max_knn, const int32_t* ids, size_t n_ids, int eig_order, double* out, int nthreads)
{ selected_driver<double>(xyz, n, r, max_knn, ids, n_ids, eig_order, out, nthreads); }

// symmetric 3x3 eigen solve exposed for unit tests (row-major a[9] -> w[3], V[9] column eigenvectors)
void oracle_eig3_f32(const float* a, float* w, float* V)
{ float A[3][3], VV[3][3]; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[i][j] = a[3 * i + j]; eig3_sym(A, w, VV); for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) V[3 * i + j] = VV[i][j]; }
void oracle_eig3_f64(const double* a, double* w, double* V)
{ double A[3][3], VV[3][3]; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[i][j] = a[3 * i + j]; eig3_sym(A, w, VV); for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) V[3 * i + j] = VV[i][j]; }

}  // extern "C"
