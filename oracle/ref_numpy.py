"""numpy restatement of the pgeof hot path -- TEST INFRASTRUCTURE ONLY.

This file is the CPU oracle the CUDA path is checked against.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it; nothing under ``point_geometric_features_b200/``
does (the product has no CPU fallback).

Parity pinning
-------------------------------------
* neighbour *indices* (``knn_search`` / ``radius_search``): PINNED.  The
  reference's own tests compare against ``scipy.spatial.KDTree``
  (``tests/test_pgeof.py:8-27``); ``tests/test_oracle.py`` checks this oracle
  against scipy on the same (seeded) inputs.
* squared distances, all feature values, ``compute_features_optimal`` and
  ``compute_features_selected``: **parity unpinned** -- the reference holds no
  golden vector or external check for them, and its third-party arithmetic
  (Eigen 3.4.0 ``SelfAdjointEigenSolver``, nanoflann @9c930ba4) is absent from
  ``/root/reference`` (empty submodules), so the reference cannot be executed
  here.  The restatement follows the reference source line by line (citations
  below, all relative to ``/root/reference``) and is cross-checked against
  LAPACK ``eigh`` and analytic known answers.

Arithmetic.  Features are evaluated in float64 from the float32 inputs (the
tolerance target of BASELINE.json is 1e-4 abs / 1e-3 rel against this).
Neighbour search uses the *defined float32 metric*
``d2 = fl(fl(fl(dx*dx) + fl(dy*dy)) + fl(dz*dz))`` (nanoflann ``L2_Simple``
accumulates ``diff*diff`` sequentially in ``float``, ``nn_search.hpp:35``) with
no FMA contraction, ordered by ``(d2, index)``.

Eigen order.  ``include/pca.hpp:79-89`` takes ``es.eigenvalues()`` of Eigen's
``SelfAdjointEigenSolver`` *in returned (increasing) order* with no re-sort, so
read literally ``val(0)`` is the smallest eigenvalue and ``v2`` ("the normal")
belongs to the largest.  ``eig_order="literal"`` reproduces that;
``eig_order="documented"`` gives the decreasing order every comment and the
README assume.
"""
from __future__ import annotations

import numpy as np

EPS = 1e-3  # include/pca.hpp:30-35  epsilon<float> == epsilon<double> == 1e-3

# include/pca.hpp:47-63
FEATURE_IDS = {
    "Linearity": 0, "Planarity": 1, "Scattering": 2, "VerticalityPGEOF": 3,
    "Normal_x": 4, "Normal_y": 5, "Normal_z": 6, "Length": 7, "Surface": 8,
    "Volume": 9, "Curvature": 10, "K_optimal": 11, "Verticality": 12,
    "Eigentropy": 13,
}


# --------------------------------------------------------------------------
# PCA (include/pca.hpp:71-98)
# --------------------------------------------------------------------------
def pca_from_pointcloud(cloud, eig_order="literal"):
    """``cloud`` (k,3) -> (val[3], v0, v1, v2).  include/pca.hpp:71-98."""
    cloud = np.asarray(cloud, dtype=np.float64)
    k = cloud.shape[0]
    centered = cloud - cloud.mean(axis=0)                 # :75
    cov = centered.T @ centered / float(k)                # :76 (population cov)
    w, v = np.linalg.eigh(cov)                            # :79 ascending, like Eigen
    if eig_order == "documented":
        o = np.argsort(-w, kind="stable")                 # decreasing; ties keep Eigen's column order
        w, v = w[o], v[:, o]
    elif eig_order != "literal":
        raise ValueError("eig_order must be 'literal' or 'documented'")
    val = np.maximum(w, 0.0)                              # :85 clamp
    v0, v1, v2 = v[:, 0].copy(), v[:, 1].copy(), v[:, 2].copy()   # :87-89
    if v2[2] < 0.0:                                       # :96
        v2 = -v2
    return val, v0, v1, v2


def compute_eigentropy(val):
    """include/pca.hpp:140-150."""
    s = val.sum() + EPS
    e = val / s
    return float(-(e * np.log(e + EPS)).sum())


def features_from_pca(val, v0, v1, v2):
    """The 11 features of include/pca.hpp:160-200 in EFeatureID order."""
    f = np.zeros(11, dtype=np.float64)
    s0, s1, s2 = np.sqrt(val[0]), np.sqrt(val[1]), np.sqrt(val[2])
    fact = 1.0 / (s0 + EPS)
    f[4:7] = v2
    f[0] = (s0 - s1) * fact
    f[1] = (s1 - s2) * fact
    f[2] = s2 * fact
    f[7] = s0
    f[8] = np.sqrt(s0 * s1 + 1e-6)
    f[9] = (s0 * s1 * s2 + 1e-9) ** (1.0 / 3.0)
    f[10] = s2 / (s0 + s1 + s2 + EPS)
    if s0 > 0.0:                                          # :187
        u = val[0] * np.abs(v0) + val[1] * np.abs(v1) + val[2] * np.abs(v2)
        f[3] = u[2] / np.linalg.norm(u)
    return f


def selected_from_pca(val, v0, v1, v2, ids):
    """include/pca.hpp:212-295.  ``K_optimal`` has no case -> stays 0."""
    full = features_from_pca(val, v0, v1, v2)
    out = np.zeros(len(ids), dtype=np.float64)
    for j, fid in enumerate(ids):
        fid = int(fid)
        if 0 <= fid <= 10:
            out[j] = full[fid]
        elif fid == 12:
            out[j] = 1.0 - abs(v2[2])                     # :280-285
        elif fid == 13:
            out[j] = compute_eigentropy(val)              # :286-288
    return out


# --------------------------------------------------------------------------
# Drivers (include/pgeof.hpp)
# --------------------------------------------------------------------------
def compute_features(xyz, nn, nn_ptr, k_min=1, eig_order="literal"):
    """include/pgeof.hpp:75-117."""
    if k_min < 1:
        raise ValueError("k_min should be > 1")
    n = len(nn_ptr) - 1
    out = np.zeros((n, 11), dtype=np.float64)
    for i in range(n):
        a, b = int(nn_ptr[i]), int(nn_ptr[i + 1])
        if b - a >= k_min:                                # :103
            out[i] = features_from_pca(*pca_from_pointcloud(xyz[nn[a:b]], eig_order))
    return out


def check_scales(k_scales):
    """include/pgeof.hpp:123-132."""
    prev = 1
    for s in k_scales:
        if s < prev:
            return False
        prev = s
    return True


def compute_features_multiscale(xyz, nn, nn_ptr, k_scales, eig_order="literal"):
    """include/pgeof.hpp:159-211 -> (N, S, 11)."""
    k_scales = [int(s) for s in k_scales]
    if not check_scales(k_scales):
        raise ValueError("k_scales should be > 1 and sorted in ascending order")
    n = len(nn_ptr) - 1
    out = np.zeros((n, len(k_scales), 11), dtype=np.float64)
    for i in range(n):
        a, b = int(nn_ptr[i]), int(nn_ptr[i + 1])
        for s, ks in enumerate(k_scales):
            if b - a < ks:                                # :193 early break
                break
            out[i, s] = features_from_pca(*pca_from_pointcloud(xyz[nn[a:a + ks]], eig_order))
    return out


def compute_features_optimal(xyz, nn, nn_ptr, k_min=1, k_step=1, k_min_search=1,
                             eig_order="literal", return_margin=False):
    """include/pgeof.hpp:243-310 -> (N, 12).

    ``return_margin`` additionally returns, per row, the gap between the best
    and the runner-up eigenentropy (inf when a single k was evaluated), which
    the parity tests use to set near-tie rows apart.
    """
    if k_min < 1 and k_min_search < 1:                    # :250 (sic, '&&')
        raise ValueError("k_min and k_min_search should be > 1")
    if k_step < 1:
        raise ValueError("k_step should be >= 1")         # reference: UB (mod 0, :283)
    n = len(nn_ptr) - 1
    out = np.zeros((n, 12), dtype=np.float64)
    margin = np.full(n, np.inf)
    for i in range(n):
        a, b = int(nn_ptr[i]), int(nn_ptr[i + 1])
        k_nn = b - a
        if k_nn >= k_min and k_nn >= k_min_search:        # :272
            k0 = min(max(k_min, k_min_search), k_nn)      # :274
            best, best_h, best_k, second = None, 1.0, k_nn, np.inf
            for k in range(k0, k_nn + 1):
                if k > k0 and k % k_step != 0 and k != k_nn:   # :283
                    continue
                pca = pca_from_pointcloud(xyz[nn[a:a + k]], eig_order)
                h = compute_eigentropy(pca[0])
                if k == k0 or h < best_h:                 # :289 strict '<'
                    if k != k0:
                        second = best_h
                    best, best_h, best_k = pca, h, k
                else:
                    second = min(second, h)
            out[i, :11] = features_from_pca(*best)
            out[i, 11] = float(best_k)
            margin[i] = second - best_h
    return (out, margin) if return_margin else out


# --------------------------------------------------------------------------
# Neighbour search (include/nn_search.hpp) -- defined float32 metric
# --------------------------------------------------------------------------
def sqdist_f32(q, pts):
    """fl(fl(fl(dx*dx)+fl(dy*dy))+fl(dz*dz)) for one query against (m,3) pts."""
    q = np.asarray(q, dtype=np.float32)
    pts = np.asarray(pts, dtype=np.float32)
    dx = q[0] - pts[:, 0]
    dy = q[1] - pts[:, 1]
    dz = q[2] - pts[:, 2]
    return (dx * dx + dy * dy) + dz * dz                  # float32 elementwise, no FMA


def _order(d2):
    """Indices sorted by (d2, index): stable argsort on d2."""
    return np.argsort(d2, kind="stable")


def knn_search(data, query, knn):
    """include/nn_search.hpp:31-67, brute force.  -> (uint32 (n,k), float32 (n,k))."""
    data = np.asarray(data, dtype=np.float32)
    query = np.asarray(query, dtype=np.float32)
    if knn > data.shape[0]:                               # :37
        raise ValueError("knn size is greater than the data point cloud size")
    nq = query.shape[0]
    idx = np.empty((nq, knn), dtype=np.uint32)
    d2o = np.empty((nq, knn), dtype=np.float32)
    for i in range(nq):
        d2 = sqdist_f32(query[i], data)
        o = _order(d2)[:knn]
        idx[i] = o
        d2o[i] = d2[o]
    return idx, d2o


def radius_search(data, query, search_radius, max_knn):
    """include/nn_search.hpp:85-132: strict d2 < fl(r*r), max_knn nearest, pad -1 / 0."""
    data = np.asarray(data, dtype=np.float32)
    query = np.asarray(query, dtype=np.float32)
    if max_knn > data.shape[0]:                           # :92-95
        raise ValueError("max knn size is greater than the data point cloud size")
    r = np.float32(search_radius)
    r2 = np.float32(r * r)                                # :98 (float)
    nq = query.shape[0]
    idx = np.full((nq, max_knn), -1, dtype=np.int32)      # :104
    d2o = np.zeros((nq, max_knn), dtype=np.float32)       # :108
    for i in range(nq):
        d2 = sqdist_f32(query[i], data)
        o = _order(d2)
        o = o[d2[o] < r2][:max_knn]                       # :117-122
        idx[i, :len(o)] = o
        d2o[i, :len(o)] = d2[o]
    return idx, d2o


def compute_features_selected(xyz, search_radius, max_knn, selected_features,
                              eig_order="literal"):
    """include/pgeof.hpp:325-375.  Metric in the dtype of ``xyz`` (f32 or f64)."""
    xyz = np.asarray(xyz)
    ids = [int(f) for f in selected_features]
    n = xyz.shape[0]
    out = np.zeros((n, len(ids)), dtype=np.float64)
    if xyz.dtype == np.float32:
        r = np.float32(search_radius)
        r2 = np.float32(r * r)
        dist = lambda i: sqdist_f32(xyz[i], xyz)          # noqa: E731
    else:
        r2 = float(search_radius) * float(search_radius)

        def dist(i):
            d = xyz[i] - xyz
            return (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    for i in range(n):
        d2 = dist(i)
        o = _order(d2)
        o = o[d2[o] < r2]                                 # :348-352 strict
        if len(o) < 2:                                    # :355
            continue
        o = o[:max_knn]                                   # :358-364
        pca = pca_from_pointcloud(xyz[o], eig_order)
        out[i] = selected_from_pca(*pca, ids)
    return out


# --------------------------------------------------------------------------
# CSR glue exactly as the README shows it (README.md:135-163)
# --------------------------------------------------------------------------
def knn_to_csr(knn_idx):
    n, k = knn_idx.shape
    nn_ptr = (np.arange(n + 1) * k).astype(np.uint32)
    nn = np.ascontiguousarray(knn_idx.reshape(-1)).astype(np.uint32)
    return nn, nn_ptr


def radius_to_csr(rad_idx):
    nn_ptr = np.r_[0, (rad_idx >= 0).sum(axis=1).cumsum()].astype(np.uint32)
    nn = rad_idx[rad_idx >= 0].astype(np.uint32)
    return nn, nn_ptr
