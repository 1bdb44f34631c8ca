#!/usr/bin/env python
"""Per-call device time of one rank's step of an N-way slab-sharded run, emulated on one GPU (diagnostic).
PROBE_NCU=1 runs the step a few times only (for `ncu --metrics gpu__time_duration.sum`)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pgeof
import point_geometric_features_b200 as b200
from point_geometric_features_b200 import synth, shard
world, rank = int(os.environ.get("WORLD", 8)), int(os.environ.get("RANK_", 3))
t = torch.from_numpy(synth.uniform_cloud(10_000_000, seed=0)).cuda()


def step():
    q = shard.slab_queries(t, rank, world)[1]
    idx, d2 = pgeof.knn_search(t, q, 50)
    ptr = (torch.arange(q.shape[0] + 1, device="cuda", dtype=torch.int64) * 50).to(torch.uint32)
    return pgeof.compute_features(t, idx.view(-1), ptr)


def timeit(name, fn, reps=8):
    fn(); torch.cuda.synchronize(); ts, hs = [], []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0 = time.perf_counter(); s.record(); out = fn(); e.record(); torch.cuda.synchronize(); hs.append(1e3 * (time.perf_counter() - h0)); ts.append(s.elapsed_time(e))
    print("%-44s dev %.3f ms  host %.3f ms" % (name, min(ts), min(hs)), flush=True)
    return out


if os.environ.get("PROBE_NCU"):
    for _ in range(4):
        step()
    torch.cuda.synchronize()
    sys.exit(0)
rows, q = timeit("slab_queries(rank %d of %d)" % (rank, world), lambda: shard.slab_queries(t, rank, world))
idx, d2 = timeit("knn_search(t, q, 50)", lambda: pgeof.knn_search(t, q, 50))
ptr = timeit("csr glue", lambda: (torch.arange(q.shape[0] + 1, device="cuda", dtype=torch.int64) * 50).to(torch.uint32))
f = timeit("compute_features", lambda: pgeof.compute_features(t, idx.view(-1), ptr))
timeit("whole step", step)
b200.profile_reset(); b200.profile_enable(True)
for _ in range(8):
    step()
torch.cuda.synchronize()
for name in ("knn_search", "features", "grid_build", "row_order"):
    ms, cnt = b200.profile_read(name)
    print("   timer %-12s %.3f ms/step" % (name, ms / 8))
