"""Generates the committed golden fixtures from the numpy oracle + scipy.

    PYTHONPATH=. python tests/golden/make_golden.py

The reference itself cannot be imported here (empty Eigen / nanoflann / Taskflow
submodules, no nanobind), so the vectors come from ``oracle/ref_numpy.py`` -- the float64
restatement that follows the reference line by line -- and, for neighbour indices, from
``scipy.spatial.KDTree`` used exactly as the reference's tests use it (tests/test_pgeof.py:8-27).
They freeze today's oracle so that later edits of the oracle, the C++ port or the kernels are
all checked against the same numbers.
"""
import os
import sys

import numpy as np
from scipy.spatial import KDTree

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ref_numpy as rn  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(20261017)
    # README example shape (README.md:129-143), reduced: uniform [0,1)^3, k = 20
    xyz = rng.random((600, 3), dtype=np.float32)
    knn_scipy = KDTree(xyz).query(xyz, k=20)[1].astype(np.uint32)
    idx, d2 = rn.knn_search(xyz, xyz, 20)
    assert (idx == knn_scipy).all()
    nn, nn_ptr = rn.knn_to_csr(idx)
    out = {"xyz": xyz, "knn_idx": idx, "knn_d2": d2}
    for order in ("literal", "documented"):
        out["features_" + order] = rn.compute_features(xyz, nn, nn_ptr, 1, order)
        out["multiscale_" + order] = rn.compute_features_multiscale(xyz, nn, nn_ptr, [5, 10, 20], order)
        opt, margin = rn.compute_features_optimal(xyz, nn, nn_ptr, 1, 1, 5, order, return_margin=True)
        out["optimal_" + order] = opt
        out["optimal_margin_" + order] = margin
    # radius search as in tests/test_pgeof.py:18-27
    r, max_knn = 0.2, 10
    _, rs = KDTree(xyz).query(xyz, k=max_knn, distance_upper_bound=r)
    rs[rs == xyz.shape[0]] = -1
    ridx, rd2 = rn.radius_search(xyz, xyz, r, max_knn)
    assert (ridx == rs).all()
    out["radius_idx"], out["radius_d2"] = ridx, rd2
    ids = [12, 10, 13, 0, 1, 2, 4, 5, 6, 7, 8, 9, 11]
    out["selected_ids"] = np.asarray(ids, np.int32)
    for order in ("literal", "documented"):
        out["selected_f32_" + order] = rn.compute_features_selected(xyz, 0.15, 12, ids, order)
        out["selected_f64_" + order] = rn.compute_features_selected(xyz.astype(np.float64) * 1.000000123, 0.15, 12, ids, order)
    np.savez_compressed(os.path.join(HERE, "readme_600.npz"), **out)
    print("wrote readme_600.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
