#!/usr/bin/env python
"""Fused knn_features against knn_search + compute_features: bit-identical rows and device time (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pgeof
import point_geometric_features_b200 as b200
from point_geometric_features_b200 import synth
n = int(os.environ.get("N", 10_000_000))
t = torch.from_numpy(synth.uniform_cloud(n, seed=0)).cuda()
b200.set_eig_order("literal")
for k in [int(x) for x in os.environ.get("KS", "20,50").split(",")]:
    idx, d2 = pgeof.knn_search(t, t, k)
    ptr = (torch.arange(n + 1, device="cuda", dtype=torch.int64) * k).to(torch.uint32)
    ref = pgeof.compute_features(t, idx.view(-1), ptr)
    del idx, d2
    ts = []
    for _ in range(4):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); f = b200.knn_features(t, k); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    same = bool((f.view(torch.int32) == ref.view(torch.int32)).all())
    nbad = int((f.view(torch.int32) != ref.view(torch.int32)).any(1).sum())
    print("k=%d fused knn_features %.2f ms (min of 4)  bit-identical to the two calls: %s (%d rows differ)" % (k, min(ts[1:]), same, nbad), flush=True)
