#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): points/s of knn_search(k=50) + compute_features on 10 M
synthetic points, 1/2/4/8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One *step* = one pass of the hot path over the whole cloud:
    idx, d2 = knn_search(xyz, xyz[shard], 50); nn = idx.view(-1); nn_ptr = arange * 50;
    feats   = compute_features(xyz, nn, nn_ptr)
`value`  : device-resident throughput (inputs already in HBM, CUDA events, max over ranks).
`e2e`    : the same pipeline through the drop-in module with HOST (numpy, pinned) buffers --
           H2D / D2H copies inside the timed region.
`roofline`: the dominant kernel's algorithmic bytes / its measured duration against the measured
           HBM copy bandwidth (MEASURED_PEAKS.json).
`cpu_baseline` / `--impl reference`: the C++ restatement of the reference's CPU path (oracle/, kind
           "port": the reference itself is unbuildable here) on all host threads, bounded sample.
Multi-GPU: the cloud is replicated, queries are sharded contiguously across ranks (strong scaling of
the one 10 M cloud; no data-path collective), rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "points/s: knn_search k=50 + compute_features, 10M pts"
UNIT = "points/s"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--points", type=int, default=10_000_000)
    p.add_argument("--knn", type=int, default=50)
    p.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    p.add_argument("--cpu-sample", type=int, default=500_000, help="points of the bounded CPU sample")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--e2e-steps", type=int, default=0, help="0 = same as --steps")
    return p.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  In-process NVML
    (two cheap queries every 25 ms from a thread); falls back to an `nvidia-smi -lms` child process."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.rows, self.proc, self.nvml, self.stop_flag = [], None, None, False
        if os.environ.get('PGEOF_BENCH_NO_CLOCKS'):
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "250"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if index < len(ids) and ids[index].isdigit():
                return int(ids[index])
        return index

    def _poll(self):
        nv = self.nvml
        bits = [nv.nvmlClocksThrottleReasonHwSlowdown, nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                nv.nvmlClocksThrottleReasonSwThermalSlowdown, nv.nvmlClocksThrottleReasonSwPowerCap]
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.rows.append((time.perf_counter(), sm, [bool(r & b) for b in bits]))
            except Exception:
                pass
            time.sleep(float(os.environ.get('PGEOF_BENCH_CLOCK_PERIOD', '0.025')))

    def _read(self):
        for line in self.proc.stdout:
            c = [x.strip() for x in line.split(",")]
            if len(c) >= 7 and c[0].replace(".", "").isdigit():
                self.sm_max = float(c[1]) if c[1].replace(".", "").isdigit() else None
                self.rows.append((time.perf_counter(), float(c[0]), [x.lower().startswith("active") for x in c[3:7]]))

    def stop(self, t0, t1):
        if self.nvml is None and not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock source"]}
        time.sleep(0.06)
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        rows = [r for r in self.rows if t0 <= r[0] <= t1] or list(self.rows)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        reasons = [n for i, n in enumerate(self.NAMES) if any(r[2][i] for r in rows)]
        return {"sm_mhz": statistics.median(r[1] for r in rows), "sm_max_mhz": getattr(self, "sm_max", None), "reasons": reasons,
                "samples": len(rows), "source": "nvml" if self.nvml else "nvidia-smi"}


def sample_cloud(n_total, n_sample, seed=0):
    """Bounded CPU sample of the same workload: a sub-cube of the uniform cloud at EQUAL density."""
    from point_geometric_features_b200 import synth
    extent = 200.0 * (n_sample / float(n_total)) ** (1.0 / 3.0)
    return synth.uniform_cloud(n_sample, seed=seed, extent=extent)


def cpu_step(cpu, xyz, k):
    import numpy as np
    idx, _ = cpu.knn_search(xyz, xyz, k)                       # KD-tree build + query, all host threads
    nn_ptr = (np.arange(xyz.shape[0] + 1) * k).astype(np.uint32)   # README glue (README.md:135-141)
    nn = idx.reshape(-1)
    return cpu.compute_features(xyz, nn, nn_ptr, 1, "literal", f64=False)


def run_reference(args, rank, real_stdout):
    """--impl reference: the reference's CPU path.  The reference binary cannot be built here (empty
    third_party submodules), so this times oracle/cpu_ref.cpp, the C++ restatement (kind "port")."""
    if rank != 0:
        return
    from oracle import cpu
    cpu.build()
    n_s = min(args.cpu_sample, args.points)
    xyz = sample_cloud(args.points, n_s)
    for _ in range(args.warmup):
        cpu_step(cpu, xyz, args.knn)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(cpu, xyz, args.knn)
    dt = time.perf_counter() - t0
    value = n_s * args.steps / dt
    sample = "%d-point sub-cube of the %d-point uniform cloud at equal density (same k=%d), KD-tree build included" % (n_s, args.points, args.knn)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": "uniform [0,200)^3 float32 cloud, knn_search(k=%d) -> CSR -> compute_features (11 features)" % args.knn,
                                            "points": args.points, "knn": args.knn, "sample_points": n_s},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu.hardware_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(real_stdout, line)


def main():
    # stdout carries exactly ONE JSON line: NCCL / torchrun banners written to fd 1 by native code go to stderr instead
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        _main(real_stdout)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)


def emit(real_stdout, line):
    os.write(real_stdout, (json.dumps(line) + "\n").encode())


def _main(real_stdout):
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, real_stdout)
        return

    # The drop-in API returns freshly allocated outputs (2 x 2 GB + 0.44 GB per step here), served by torch's caching
    # allocator.  Left to split its 2 GB blocks for the 40-80 MB CSR glue tensors, it periodically has to cudaMalloc a
    # new 2 GB block (80-500 ms, inside knn_search's host time); blocks above 256 MB are therefore never split.
    os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "max_split_size_mb:256")
    import numpy as np
    import torch
    import torch.distributed as dist

    import pgeof
    import point_geometric_features_b200 as b200
    from point_geometric_features_b200 import shard, synth

    if b200.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    b200.set_eig_order("literal")

    n, k = args.points, args.knn
    weak = args.scaling == "weak"
    xyz = synth.uniform_cloud(n, seed=rank if weak else 0)         # strong: the same cloud on every rank (replicated)
    lo, hi = (0, n) if weak else shard.shard_range(n, rank, world)
    n_total = n * world if weak else n
    t = torch.from_numpy(xyz).to(dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    debug = bool(os.environ.get("PGEOF_BENCH_DEBUG"))
    host_log = []

    def step():
        h0 = time.perf_counter()
        if world == 1 or weak:
            q = t
        else:                                                      # this rank's slab of the replicated cloud (no collective)
            q = t[shard.spatial_shard(t, rank, world)]
        idx, d2 = pgeof.knn_search(t, q, k)
        h1 = time.perf_counter()
        nn_ptr = (torch.arange(q.shape[0] + 1, device=dev, dtype=torch.int64) * k).to(torch.uint32)
        h2 = time.perf_counter()
        feats = pgeof.compute_features(t, idx.view(-1), nn_ptr)
        if debug:
            host_log.append((1e3 * (h1 - h0), 1e3 * (h2 - h1), 1e3 * (time.perf_counter() - h2)))
        return idx, d2, feats

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    out = None
    for _ in range(max(args.warmup, 3)):
        out = None                                                 # a caller's loop overwrites its previous results
        out = step()
    barrier()
    b200.profile_reset()
    b200.profile_enable(not os.environ.get('PGEOF_BENCH_NO_PROFILE'))
    b200.reset_launch_count()
    sampler = ClockSampler(local_rank)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall0 = time.perf_counter()
    for s, e in ev:
        out = None
        flush.zero_()                                              # L2 flush between timed iterations (not timed)
        s.record()
        out = step()
        e.record()
        torch.cuda.synchronize()                                   # outside the step's event pair, like the flush: keeps the host from
                                                                   # racing ahead into the next step's allocations (measured: removes 2-500 ms host stalls)
    barrier()
    t_wall1 = time.perf_counter()
    clocks = sampler.stop(t_wall0, t_wall1)
    launches = b200.launch_count() + 2 * args.steps               # + torch arange / cast per step (not ours, listed for honesty)
    own_launches = b200.launch_count()
    b200.profile_enable(False)
    step_ms = [s.elapsed_time(e) for s, e in ev]
    if debug and rank == 0:
        for i, ms in enumerate(step_ms):
            h = host_log[len(host_log) - len(step_ms) + i]
            print("step %d dev %.2f ms | host knn %.2f glue %.2f feat %.2f" % (i, ms, h[0], h[1], h[2]), file=sys.stderr)
    dev_ms = sum(step_ms)
    tm = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    dev_ms_max = float(tm.item())
    value = n_total * args.steps / (dev_ms_max * 1e-3)
    kernels = {}
    for name in ("knn_search", "features", "grid_build"):
        ms, cnt = b200.profile_read(name)
        kernels[name] = {"ms_per_launch": ms / max(cnt, 1), "launches": int(cnt)}
    # ---- roofline of the dominant kernel (algorithmic bytes of SURVEY.md 8d) --------------------
    peak, peak_src = peaks()
    rows = int(out[0].shape[0]) if out is not None else hi - lo
    alg = {"knn_search": (24 + 8 * k) * rows, "features": (48 + 16 * k) * rows}
    dom = max(("knn_search", "features"), key=lambda nme: kernels[nme]["ms_per_launch"])
    ach = alg[dom] / (kernels[dom]["ms_per_launch"] * 1e-3) / 1e9
    # dram__bytes_read + dram__bytes_write per row from the latest `ncu --set full` capture (profiles/r1c_summary.md)
    traffic_per_row = {"knn_search": 571.5, "features": 680.4}
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic_per_row[dom] * rows, "traffic_source": "profiles/r1c_summary.md (ncu --set full, 10 M rows, scaled by rows)",
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg[dom],
                "note": "knn tile kernel is FP32/ALU-issue bound (622 candidate evaluations + a 985-comparator sorting network per query), not HBM bound; the gather-bound feature kernel is the one to read against the HBM roofline (all_kernels.features); see DESIGN.md",
                "all_kernels": {nme: {"ms": kernels[nme]["ms_per_launch"],
                                      "achieved_gbs": (alg[nme] / (kernels[nme]["ms_per_launch"] * 1e-3) / 1e9) if nme in alg and kernels[nme]["ms_per_launch"] > 0 else None}
                                for nme in kernels}}

    rows_dev = rows
    del out
    rows = hi - lo

    # ---- end to end through the drop-in module with host buffers --------------------------------
    e2e = None
    if not args.no_e2e:
        e2e_steps = args.e2e_steps or args.steps
        hx = torch.from_numpy(xyz).pin_memory().numpy()             # pinned host input
        hq = hx if (lo, hi) == (0, n) else hx[lo:hi]

        def host_step():
            knn, _d2 = pgeof.knn_search(hx, hq, k)                  # numpy in -> numpy out (H2D + D2H inside)
            nn_ptr = np.arange(0, (hi - lo + 1) * k, k, dtype=np.uint32)   # README glue (README.md:135-141) in one numpy pass
            nn = knn.reshape(-1)                                    # zero-copy: knn is already uint32
            f = pgeof.compute_features(hx, nn, nn_ptr)
            return float(f[0, 0])                                   # read the result on the host

        for _ in range(2):
            host_step()
        barrier()
        t0 = time.perf_counter()
        e2e_ms = []
        for _ in range(e2e_steps):
            ts = time.perf_counter()
            host_step()
            e2e_ms.append(1e3 * (time.perf_counter() - ts))
        barrier()
        dt = time.perf_counter() - t0
        tm = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dt = float(tm.item())
        h2d = n * 12 + (0 if (lo, hi) == (0, n) else rows * 12) + n * 12 + rows * k * 4 + (rows + 1) * 4
        d2h = rows * k * 8 + rows * 44
        e2e = {"value": n_total * e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": 1e3 * dt / e2e_steps, "steps": e2e_steps,
               "step_ms_min_median_max": [min(e2e_ms), statistics.median(e2e_ms), max(e2e_ms)],
               "path": "pgeof.knn_search(numpy) -> reshape/arange glue -> pgeof.compute_features(numpy); pinned host input, pinned pooled outputs"}

    # ---- CPU baseline beside it (rank 0, N = 1 only) ---------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import cpu
        cpu.build()
        n_s = min(args.cpu_sample, n)
        cx = sample_cloud(n, n_s)
        cpu_step(cpu, cx, k)
        reps, t0 = 0, time.perf_counter()
        while reps < 3 or time.perf_counter() - t0 < 10.0:
            cpu_step(cpu, cx, k)
            reps += 1
            if time.perf_counter() - t0 > 30.0:
                break
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": n_s * reps / dt, "unit": UNIT, "cores": cpu.hardware_threads(), "kind": "port",
                        "sample": "%d-point sub-cube at equal density, %d repetitions, KD-tree build + kNN(k=%d) + CSR glue + compute_features" % (n_s, reps, k)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": "uniform [0,200)^3 float32 cloud (seed 0), knn_search(k=%d) -> CSR view -> compute_features (11 features); grid build included in every step" % k,
                           "points": n_total, "knn": k, "rows_per_rank": rows_dev,
                           "parallelism": "query-sharded x%d (z slabs of ~n/N points from sample quantiles, computed inside the step), cloud and grid replicated, no data-path collective; e2e shards by contiguous row range" % world,
                           "l2": "512 MiB buffer zeroed between timed iterations; per-step working set %.1f GB >> 126 MB L2" % ((rows * (k * 12 + 44) + n * 28) / 1e9),
                           "eig_order": "literal"},
                "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "clocks": clocks,
                "gpu_launches": int(own_launches), "gpu_launches_per_step": own_launches / args.steps,
                "wall_s_timed_region": t_wall1 - t_wall0,
                "step_ms_min_median_max": [min(step_ms), statistics.median(step_ms), max(step_ms)]}
        emit(real_stdout, line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
