#!/usr/bin/env python
"""Per-call device time of one rank's step of an N-way slab-sharded run, emulated on one GPU (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pgeof
from point_geometric_features_b200 import synth, shard
world, rank = int(os.environ.get("WORLD", 8)), int(os.environ.get("RANK_", 3))
t = torch.from_numpy(synth.uniform_cloud(10_000_000, seed=0)).cuda()


def timeit(name, fn, reps=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); out = fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    print("%-44s %.3f ms" % (name, min(ts)), flush=True)
    return out


rows = timeit("spatial_shard(rank %d of %d)" % (rank, world), lambda: shard.spatial_shard(t, rank, world))
q = timeit("t[rows]", lambda: t[rows])
idx, d2 = timeit("knn_search(t, q, 50)", lambda: pgeof.knn_search(t, q, 50))
ptr = timeit("csr glue", lambda: (torch.arange(q.shape[0] + 1, device="cuda", dtype=torch.int64) * 50).to(torch.uint32))
f = timeit("compute_features", lambda: pgeof.compute_features(t, idx.view(-1), ptr))
