#!/usr/bin/env python
"""The reference's own benchmark workloads (tests/bench_knn.py, tests/bench_jakteristics.py) through the numpy API."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pgeof
from pgeof import EFeatureID


def timed(name, fn, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); fn(); ts.append(1e3 * (time.perf_counter() - t))
    print("%-70s %8.2f ms (min of %d, host numpy in/out)" % (name, min(ts), reps), flush=True)


rng = np.random.default_rng(0)
x1m = rng.uniform(0.0, 200.0, size=(1000000, 3)).astype(np.float32)
timed("bench_knn.py: knn_search 1M uniform k=50", lambda: pgeof.knn_search(x1m, x1m, 50))
timed("bench_knn.py: radius_search 1M uniform r=0.2 max_knn=30", lambda: pgeof.radius_search(x1m, x1m, 0.2, 30))
x10k = rng.uniform(0.0, 200.0, size=(10000, 3))
timed("bench_jakteristics.py: compute_features_selected 10k f64 r=5 knn=50 [Verticality]",
      lambda: pgeof.compute_features_selected(x10k, 5.0, 50, [EFeatureID.Verticality]))
x1md = x1m.astype(np.float64)
timed("compute_features_selected 1M f64 r=5 knn=50 [Verticality]", lambda: pgeof.compute_features_selected(x1md, 5.0, 50, [EFeatureID.Verticality]), reps=3)
timed("compute_features_selected 1M f32 r=5 knn=50 [Verticality]", lambda: pgeof.compute_features_selected(x1m, 5.0, 50, [EFeatureID.Verticality]), reps=3)
