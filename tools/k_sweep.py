#!/usr/bin/env python
"""knn_search + compute_features device time for several k on the 10 M uniform cloud (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pgeof
import point_geometric_features_b200 as b200
from point_geometric_features_b200 import synth
n = int(os.environ.get("N", 10_000_000))
t = torch.from_numpy(synth.uniform_cloud(n, seed=0)).cuda()
b200.set_eig_order("literal")
for k in [int(x) for x in os.environ.get("KS", "8,16,20,32,33,50,52,53,64,65,100").split(",")]:
    res = []
    for it in range(3):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(); idx, d2 = pgeof.knn_search(t, t, k); e[1].record()
        ptr = (torch.arange(n + 1, device="cuda", dtype=torch.int64) * k).to(torch.uint32)
        f = pgeof.compute_features(t, idx.view(-1), ptr); e[2].record(); torch.cuda.synchronize()
        res.append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
        del idx, d2, f
    print("k=%3d knn_search %7.2f ms  glue+compute_features %6.2f ms" % (k, min(r[0] for r in res), min(r[1] for r in res)), flush=True)
