#!/usr/bin/env python
"""Device-resident timing of BASELINE.json configs 2-5 on one GPU (diagnostic; the bench line is config M)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pgeof
import point_geometric_features_b200 as b200
from point_geometric_features_b200 import synth

dev = torch.device("cuda", 0)
b200.set_eig_order("literal")


def timed(name, fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); out = fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    print("%-58s %8.2f ms (min of %d)" % (name, min(ts), reps), flush=True)
    return out


which = os.environ.get("CONFIGS", "2345")
if "2" in which:
    t = torch.from_numpy(synth.uniform_cloud(1_000_000, seed=0)).to(dev)
    idx, _ = timed("C2 1M uniform knn_search k=50", lambda: pgeof.knn_search(t, t, 50))
    ptr = (torch.arange(t.shape[0] + 1, device=dev, dtype=torch.int64) * 50).to(torch.uint32)
    timed("C2 1M compute_features", lambda: pgeof.compute_features(t, idx.view(-1), ptr))
if "3" in which:
    t = torch.from_numpy(synth.lidar_like_cloud(10_000_000, seed=0)).to(dev)
    ridx, _ = timed("C3 10M lidar radius_search r=0.2 max_k=64 (padded)", lambda: pgeof.radius_search(t, t, 0.2, 64))
    print("   mean neighbours/row %.1f" % float((ridx >= 0).sum(1).float().mean()))
    del ridx
    nn, ptr = timed("C3 10M lidar radius_search_csr", lambda: b200.radius_search_csr(t, t, 0.2, 64))
    timed("C3 10M compute_features on the radius CSR (nnz=%d)" % nn.shape[0], lambda: pgeof.compute_features(t, nn, ptr))
    idx, _ = timed("C3' 10M lidar knn_search k=50 (non-uniform data)", lambda: pgeof.knn_search(t, t, 50))
    del nn, ptr, idx
if "4" in which:
    t = torch.from_numpy(synth.uniform_cloud(10_000_000, seed=0)).to(dev)
    idx, _ = timed("C4 10M uniform knn_search k=100", lambda: pgeof.knn_search(t, t, 100))
    ptr = (torch.arange(t.shape[0] + 1, device=dev, dtype=torch.int64) * 100).to(torch.uint32)
    timed("C4 10M compute_features_multiscale [10,20,50,100]", lambda: pgeof.compute_features_multiscale(t, idx.view(-1), ptr, [10, 20, 50, 100]))
    if "5" in which:
        timed("C5 10M-row shard compute_features_optimal 10..100 step 1", lambda: pgeof.compute_features_optimal(t, idx.view(-1), ptr, 10, 1, 10))
