#!/usr/bin/env python
"""kNN on the LiDAR-like scene (non-uniform data): time split between the tile kernel and the warp-per-query routine."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pgeof
import point_geometric_features_b200 as b200
from point_geometric_features_b200 import synth
n = int(os.environ.get("N", 10_000_000)); k = int(os.environ.get("K", 50))
t = torch.from_numpy(synth.lidar_like_cloud(n, seed=0)).cuda()
for _ in range(2):
    pgeof.knn_search(t, t, k)
torch.cuda.synchronize()
ts = []
for _ in range(3):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); pgeof.knn_search(t, t, k); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
print("lidar kNN k=%d n=%d: %.2f ms (min of 3)" % (k, n, min(ts)))
