#!/usr/bin/env python
"""Per-call device / host timing of the bench step (diagnostic, not a bench)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pgeof
import point_geometric_features_b200 as b200
from point_geometric_features_b200 import synth

n, k = int(os.environ.get("N", 10_000_000)), int(os.environ.get("K", 50))
dev = torch.device("cuda", 0)
t = torch.from_numpy(synth.uniform_cloud(n, seed=0)).to(dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
b200.set_eig_order("literal")
E = lambda: torch.cuda.Event(enable_timing=True)
if os.environ.get('PROFILE'): b200.profile_enable(True)
if os.environ.get('SAMPLER'):
    import threading, pynvml
    pynvml.nvmlInit(); hnd = pynvml.nvmlDeviceGetHandleByIndex(0)
    def poll():
        while True:
            pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM); pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(hnd); time.sleep(0.025)
    threading.Thread(target=poll, daemon=True).start()
if os.environ.get('HOLD'): held = None
import gc
if os.environ.get('NOGC'): gc.disable()
if os.environ.get('NICE'): os.nice(-15)
for it in range(int(os.environ.get('ITS', 10))):
    flush.zero_()
    e = [E() for _ in range(4)]
    h0 = time.perf_counter(); e[0].record()
    idx, d2 = pgeof.knn_search(t, t, k)
    h1 = time.perf_counter(); e[1].record()
    nn_ptr = (torch.arange(n + 1, device=dev, dtype=torch.int64) * k).to(torch.uint32)
    h2 = time.perf_counter(); e[2].record()
    f = pgeof.compute_features(t, idx.view(-1), nn_ptr)
    h3 = time.perf_counter(); e[3].record()
    if not os.environ.get('NOSYNC'): torch.cuda.synchronize()
    h4 = time.perf_counter()
    print("it %d dev: knn %.2f glue %.2f feat %.2f total %.2f | host: knn %.2f glue %.2f feat %.2f sync %.2f" % (
        it, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3]), e[0].elapsed_time(e[3]),
        1e3 * (h1 - h0), 1e3 * (h2 - h1), 1e3 * (h3 - h2), 1e3 * (h4 - h3)), flush=True)
