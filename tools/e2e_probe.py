#!/usr/bin/env python
"""Host-buffer (numpy in / numpy out) timing of the bench pipeline, call by call (diagnostic)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pgeof
import point_geometric_features_b200 as b200
from point_geometric_features_b200 import synth

n, k = int(os.environ.get("N", 10_000_000)), int(os.environ.get("K", 50))
xyz = synth.uniform_cloud(n, seed=0)
hx = torch.from_numpy(xyz).pin_memory().numpy() if not os.environ.get("PAGEABLE") else xyz
b200.set_eig_order("literal")
for it in range(int(os.environ.get("ITS", 5))):
    t0 = time.perf_counter()
    knn, d2 = pgeof.knn_search(hx, hx, k)
    t1 = time.perf_counter()
    nn_ptr = (np.arange(n + 1, dtype=np.uint64) * k).astype(np.uint32)
    nn = knn.reshape(-1)
    t2 = time.perf_counter()
    f = pgeof.compute_features(hx, nn, nn_ptr)
    t3 = time.perf_counter()
    print("it %d knn_search %.1f ms  glue %.1f ms  compute_features %.1f ms  total %.1f ms" % (it, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t3 - t0)), flush=True)
    del knn, d2, f, nn
