"""Small invocation of every entry point, meant to run under compute-sanitizer on the GPU box:
    compute-sanitizer --tool memcheck python tests/sanitize_small.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pgeof  # noqa: E402
import point_geometric_features_b200 as b200  # noqa: E402
from point_geometric_features_b200 import synth  # noqa: E402

xyz = synth.lidar_like_cloud(6000, seed=0)
for k in (5, 40, 100, 300):
    idx, d2 = pgeof.knn_search(xyz, xyz, k)
q = np.concatenate([xyz[:500], np.random.default_rng(0).uniform(-50, 200, (200, 3)).astype(np.float32)])
pgeof.knn_search(xyz, q, 20)
ridx, _ = pgeof.radius_search(xyz, xyz, 0.5, 32)
nn, nn_ptr = synth.radius_csr(ridx)
pgeof.compute_features(xyz, nn, nn_ptr, 3)
pgeof.compute_features_multiscale(xyz, nn, nn_ptr, [2, 4, 8, 16, 32])
pgeof.compute_features_optimal(xyz, nn, nn_ptr, 1, 2, 4)
nn2, ptr2 = b200.radius_search_csr(xyz, xyz, 0.5, 32)
assert (nn2 == nn).all() and (ptr2 == nn_ptr).all()
pgeof.compute_features_selected(xyz, 0.4, 10, [pgeof.EFeatureID.Verticality, pgeof.EFeatureID.Eigentropy])
pgeof.compute_features_selected(xyz.astype(np.float64), 0.4, 10, [pgeof.EFeatureID.Verticality, pgeof.EFeatureID.Eigentropy])
nn3, ptr3 = synth.knn_csr(idx)
pgeof.compute_features(xyz, nn3, ptr3)
# paths added in round 1: larger uniform cloud (tile kernel fast path + repairs), local queries on a clipped grid with
# balls that outgrow it, padded radius search on the tile kernel, permuted feature rows, fused knn_features
import torch  # noqa: E402
big = synth.uniform_cloud(60000, seed=1)
pgeof.knn_search(big, big, 50)
pgeof.knn_search(big, big[(big[:, 2] > 80) & (big[:, 2] < 100)], 20)
pgeof.knn_search(big, (np.float32([100, 100, 100]) + np.random.default_rng(1).uniform(-1, 1, (64, 3))).astype(np.float32), 64)
pgeof.radius_search(big, big, 6.0, 40)
i50, _ = pgeof.knn_search(big, big, 23)
pgeof.compute_features(big, *synth.knn_csr(i50))
t = torch.from_numpy(big).cuda()
b200.knn_features(t, 50)
b200.knn_features(t, 9, 1, True)
torch.cuda.synchronize()
print("sanitize_small ok")
# paths added in round 2: any k (global-scratch kernel), CSR-native kNN, 64-bit offsets, the chunked host pipeline, the optimal-k
# scan at k = 100, and -- SAN_VARIANTS=1 -- the kernels kept behind switches (pair, rolled rebuild, 160-key tile, no lock step)
pgeof.knn_search(xyz, xyz[:200], 700)
pgeof.radius_search(xyz, xyz[:200], 5.0, 600)
b200.knn_search_csr(xyz, xyz, 12)
i100, _ = pgeof.knn_search(big, big, 100)
n100, p100 = synth.knn_csr(i100)
pgeof.compute_features_optimal(big, n100, p100, 1, 1, 10)
pgeof.compute_features(big, n100, p100.astype(np.uint64))
os.environ["PGEOF_HOST_CHUNK_MB"] = "1"
pgeof.compute_features(big, n100, p100)
pgeof.compute_features_multiscale(big, n100, p100, [10, 50, 100])
del os.environ["PGEOF_HOST_CHUNK_MB"]
if os.environ.get("SAN_VARIANTS"):
    for name in ("PGEOF_KNN_PAIR", "PGEOF_KNN_ROLLED", "PGEOF_KNN_TILE128", "PGEOF_KNN_LOCK", "PGEOF_OPTIMAL_SCAN"):
        os.environ[name] = "0" if name in ("PGEOF_KNN_LOCK", "PGEOF_OPTIMAL_SCAN") else "1"
        pgeof.knn_search(big, big, 50)
        pgeof.knn_search(big, big[:20000], 100)
        pgeof.compute_features_optimal(big, n100, p100, 1, 1, 10)
        del os.environ[name]
# last session of round 2: device flavour of the single-search radius CSR pair (rows write their count and hits only), two-pass
# multiscale, closed-form optimal filter, texture gathers, two-step grid scatter -- and their switched-off counterparts
tx = torch.from_numpy(xyz).cuda()
rn, rp = b200.radius_search_csr(tx, tx, 0.5, 32)
assert (rn.cpu().numpy() == nn).all() and (rp.cpu().numpy() == nn_ptr).all()
rn2, rp2 = b200.radius_search_csr(tx, tx, 0.8, 100)                       # max_knn > 64: the warp-per-query routine writes the counts
pgeof.compute_features_multiscale(tx, rn, rp, [2, 4, 8, 16, 32])
for name, value in (("PGEOF_OPTIMAL_SCAN", "1"), ("PGEOF_MULTISCALE_SPLIT", "0"), ("PGEOF_FEATURES_TEX", "0"), ("PGEOF_GRID_SCATTER", "1"), ("PGEOF_RADIUS_TILE", "0")):
    os.environ[name] = value
    pgeof.knn_search(big, big[:20000], 30)
    b200.radius_search_csr(tx, tx, 0.5, 32)
    pgeof.compute_features_multiscale(big, n100, p100, [10, 50, 100])
    pgeof.compute_features_optimal(big, n100, p100, 1, 1, 10)
    del os.environ[name]
torch.cuda.synchronize()
print("sanitize_small round-2 paths ok")
