#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): points/s of knn_search(k=50) + compute_features on 10 M
synthetic points, 1/2/4/8 B200 -- plus the other BASELINE.json configurations behind --config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config M|C2|C3|C4|C5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One *step* = one pass of the hot path over the whole cloud (config M, the metric):
    idx, d2 = knn_search(xyz, xyz[shard], 50); nn = idx.view(-1); nn_ptr = arange * 50;
    feats   = compute_features(xyz, nn, nn_ptr)
`value`  : device-resident throughput (inputs already in HBM, CUDA events, max over ranks).
`e2e`    : the same pipeline through the drop-in module with HOST (numpy, pinned) buffers --
           H2D / D2H copies inside the timed region.
`roofline`: the dominant kernel's algorithmic bytes / its measured duration against the
           measured HBM copy bandwidth (MEASURED_PEAKS.json); `traffic` from the tracked ncu summary
           profiles/traffic.json (written by tools/ncu_traffic.py from an `ncu --set full` capture).
`cpu_baseline` / `--impl reference`: the C++ restatement of the reference's CPU path (oracle/, kind
           "port": the reference itself is unbuildable here) on all host threads, bounded sample.
Multi-GPU: the cloud is replicated, every rank owns a spatial slab of the queries (strong scaling of the
one cloud; no data-path collective) in the device leg AND in the e2e leg, rank 0 prints ONE JSON line.
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

UNIT = "points/s"

# BASELINE.json configs (C1 is the README example = a parity test, not a bench line)
CONFIGS = {
    "M": dict(points=10_000_000, cloud="uniform", search=("knn", 50), feat=("features",), cpu_sample=500_000,
              metric="points/s: knn_search k=50 + compute_features, 10M pts",
              workload="uniform [0,200)^3 float32 cloud (seed 0), knn_search(k=50) -> CSR view -> compute_features (11 features); grid build included in every step"),
    "C2": dict(points=1_000_000, cloud="uniform", search=("knn", 50), feat=("features",), cpu_sample=500_000,
               metric="points/s: knn_search k=50 + compute_features, 1M pts",
               workload="BASELINE configs[1]: 1 M uniform points, knn_search(k=50) -> CSR view -> compute_features (11 features)"),
    "C3": dict(points=10_000_000, cloud="lidar", search=("radius", 0.2, 64), feat=("features",), cpu_sample=300_000,
               metric="points/s: radius_search r=0.2 max_k=64 -> CSR -> compute_features, 10M LiDAR-like pts",
               workload="BASELINE configs[2]: 10 M LiDAR-like scene (ground / walls / poles / scatter, seed 0), radius_search r=0.2 max_k=64 emitted as CSR -> compute_features"),
    "C4": dict(points=10_000_000, cloud="uniform", search=("knn", 100), feat=("multiscale", [10, 20, 50, 100]), cpu_sample=200_000,
               metric="points/s: knn_search k=100 + compute_features_multiscale [10,20,50,100], 10M pts",
               workload="BASELINE configs[3]: 10 M uniform points, one k=100 kNN pass -> compute_features_multiscale k_scales=[10,20,50,100]"),
    "C5": dict(points=50_000_000, cloud="uniform", search=("knn", 100), feat=("optimal", 1, 1, 10), cpu_sample=100_000,
               metric="points/s: knn_search k=100 + compute_features_optimal k_min_search=10..100 k_step=1, 50M pts",
               workload="BASELINE configs[4]: 50 M uniform points, k=100 kNN -> compute_features_optimal(k_min_search=10, k_step=1), rows in shard-local CSR blocks of <= 2^32-1 neighbours"),
}


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--config", default="M", choices=sorted(CONFIGS))
    p.add_argument("--points", type=int, default=0, help="0 = the size the config names")
    p.add_argument("--knn", type=int, default=0, help="0 = the k the config names")
    p.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    p.add_argument("--cpu-sample", type=int, default=0, help="points of the bounded CPU sample (0 = per config)")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--e2e-steps", type=int, default=0, help="0 = same as --steps")
    a = p.parse_args()
    cfg = dict(CONFIGS[a.config])
    if a.points:
        cfg["points"] = a.points
    if a.knn and cfg["search"][0] == "knn":
        cfg["search"] = ("knn", a.knn)
    if a.cpu_sample:
        cfg["cpu_sample"] = a.cpu_sample
    cfg["cpu_sample"] = min(cfg["cpu_sample"], cfg["points"])
    a.cfg = cfg
    return a


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def tracked_traffic():
    """dram read + write bytes per row of the hot kernels, from the tracked summary of the latest `ncu --set full` capture."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(path)), "profiles/traffic.json"
    except Exception:
        return {}, None


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


# ------------------------------------------------------------------------------------------------
# synthetic clouds: pure numpy here, so that the CPU arms never map the product's shared libraries
# ------------------------------------------------------------------------------------------------
def load_synth():
    """point_geometric_features_b200/synth.py loaded as a stand-alone module (numpy only): importing the package would
    import the CUDA extension, which the reference arm must not touch."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_pgeof_synth", os.path.join(ROOT, "point_geometric_features_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_cloud(synth, cfg, n, seed=0, n_total=None):
    """The config's cloud at `n` points; a bounded sample (n < n_total) keeps the DENSITY of the full-size cloud."""
    n_total = n_total or n
    frac = n / float(n_total)
    if cfg["cloud"] == "lidar":
        return synth.lidar_like_cloud(n, seed=seed, extent=140.0 * frac ** 0.5)
    return synth.uniform_cloud(n, seed=seed, extent=200.0 * frac ** (1.0 / 3.0))


def cpu_step(cpu, cfg, xyz):
    """The config's pipeline on the CPU port, float32 arithmetic as the reference, README glue included."""
    import numpy as np
    s = cfg["search"]
    if s[0] == "knn":
        idx, _ = cpu.knn_search(xyz, xyz, s[1])                      # KD-tree build + query, all host threads
        nn_ptr = (np.arange(xyz.shape[0] + 1) * s[1]).astype(np.uint32)   # README glue (README.md:135-141)
        nn = idx.reshape(-1)
    else:
        idx, _ = cpu.radius_search(xyz, xyz, s[1], s[2])
        nn_ptr = np.r_[0, (idx >= 0).sum(axis=1).cumsum()].astype(np.uint32)   # README.md:157-163
        nn = idx[idx >= 0].astype(np.uint32)
    f = cfg["feat"]
    if f[0] == "features":
        return cpu.compute_features(xyz, nn, nn_ptr, 1, "literal", f64=False)
    if f[0] == "multiscale":
        return cpu.compute_features_multiscale(xyz, nn, nn_ptr, f[1], "literal", f64=False)
    return cpu.compute_features_optimal(xyz, nn, nn_ptr, f[1], f[2], f[3], "literal", f64=False)


def sample_text(cfg, n_s):
    return "%d-point sample of the %d-point %s cloud at equal density, same search and feature parameters, spatial index build included" % (
        n_s, cfg["points"], cfg["cloud"])


def run_reference(args, rank, real_stdout):
    """--impl reference: the reference's CPU path.  The reference binary cannot be built here (empty
    third_party submodules), so this times oracle/cpu_ref.cpp, the C++ restatement (kind "port")."""
    if rank != 0:
        return
    from oracle import cpu
    cpu.build()
    cfg = args.cfg
    n_s = cfg["cpu_sample"]
    xyz = make_cloud(load_synth(), cfg, n_s, 0, cfg["points"])
    for _ in range(args.warmup):
        cpu_step(cpu, cfg, xyz)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(cpu, cfg, xyz)
    dt = time.perf_counter() - t0
    value = n_s * args.steps / dt
    line = {"impl": "reference", "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": cfg["workload"], "points": cfg["points"], "sample_points": n_s},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu.hardware_threads(), "kind": "port", "sample": sample_text(cfg, n_s)},
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


def alg_bytes(cfg, rows, nnz):
    """Algorithmic bytes per launch: search 24 + 8k per row (CSR radius: 24 + 4 + 4 per neighbour kept);
    features 48 + 16 k; multiscale 4k + 4 + 12k + 44 S; optimal 4k + 4 + 12k + 48."""
    s, f = cfg["search"], cfg["feat"]
    out = {}
    if s[0] == "knn":
        out["knn_search"] = (24 + 8 * s[1]) * rows
    else:
        out["radius_search"] = 28 * rows + 4 * nnz
    if f[0] == "features":
        out["features"] = 48 * rows + 16 * nnz
    elif f[0] == "multiscale":
        out["multiscale"] = (4 + 44 * len(f[1])) * rows + 16 * nnz
    else:
        out["optimal"] = 52 * rows + 16 * nnz
    return out


def _main(real_stdout):
    args = parse()
    cfg = args.cfg
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

    n = cfg["points"]
    search, feat = cfg["search"], cfg["feat"]
    kk = search[1] if search[0] == "knn" else search[2]           # row capacity
    weak = args.scaling == "weak"
    xyz = make_cloud(synth, cfg, n, seed=rank if weak else 0)       # strong: the same cloud on every rank (replicated)
    n_total = n * world if weak else n
    t = torch.from_numpy(xyz).to(dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    # rows per CSR block: uint32 offsets hold 2^32-1 neighbours -> shard-local CSR blocks beyond that
    block_rows = max(1, min(n, (0xFFFFFFFF // kk) // 1024 * 1024))

    debug = bool(os.environ.get("PGEOF_BENCH_DEBUG"))
    host_log = []

    def features_of(cloud, nn, nn_ptr):
        if feat[0] == "features":
            return pgeof.compute_features(cloud, nn, nn_ptr)
        if feat[0] == "multiscale":
            return pgeof.compute_features_multiscale(cloud, nn, nn_ptr, feat[1])
        return pgeof.compute_features_optimal(cloud, nn, nn_ptr, feat[1], feat[2], feat[3])

    def step():
        h0 = time.perf_counter()
        if world == 1 or weak:
            q = t
        else:                                                      # this rank's slab of the replicated cloud (no collective)
            q = shard.slab_queries(t, rank, world)[1]
        outs, nnz = [], 0
        for lo in range(0, q.shape[0], block_rows):                 # one block unless the CSR would overflow uint32 offsets
            qb = q if block_rows >= q.shape[0] else q[lo:lo + block_rows]
            if search[0] == "knn":
                idx, d2 = pgeof.knn_search(t, qb, search[1])
                nn = idx.view(-1)
                nn_ptr = (torch.arange(qb.shape[0] + 1, device=dev, dtype=torch.int64) * search[1]).to(torch.uint32)
                keep = (idx, d2)
            else:
                nn, nn_ptr = b200.radius_search_csr(t, qb, search[1], search[2])
                keep = (nn, nn_ptr)
            nnz += int(nn.shape[0])
            f = features_of(t, nn, nn_ptr)
            outs.append(keep + (f,))
        if debug:
            host_log.append(1e3 * (time.perf_counter() - h0))
        return outs, nnz, q.shape[0]

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
    own_launches = b200.launch_count()
    b200.profile_enable(False)
    step_ms = [s.elapsed_time(e) for s, e in ev]
    if debug and rank == 0:
        for i, ms in enumerate(step_ms):
            print("step %d dev %.2f ms | host %.2f ms" % (i, ms, host_log[len(host_log) - len(step_ms) + i]), file=sys.stderr)
    dev_ms = sum(step_ms)
    tm = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    dev_ms_max = float(tm.item())
    value = n_total * args.steps / (dev_ms_max * 1e-3)
    kernels = {}
    for name in ("knn_search", "radius_search", "features", "multiscale", "optimal", "grid_build", "row_order"):
        ms, cnt = b200.profile_read(name)
        if cnt:
            kernels[name] = {"ms_per_step": ms / args.steps, "launches_per_step": cnt / args.steps}
    # ---- roofline of the dominant kernel (algorithmic bytes of SURVEY.md 8d) --------------------
    peak, peak_src = peaks()
    _, nnz, rows_dev = out
    alg = alg_bytes(cfg, rows_dev, nnz)
    dom = max(alg, key=lambda nme: kernels.get(nme, {"ms_per_step": 0.0})["ms_per_step"])
    dom_ms = kernels.get(dom, {"ms_per_step": float("nan")})["ms_per_step"]
    ach = alg[dom] / (dom_ms * 1e-3) / 1e9
    traffic, traffic_src = tracked_traffic()
    tcfg = traffic.get(cfg.get("traffic_key", args.config), {})
    per_row = tcfg.get(dom)
    if per_row and tcfg.get("_source"):
        traffic_src = "%s <- %s" % (traffic_src, tcfg["_source"])
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": per_row * rows_dev if per_row else None, "traffic_source": traffic_src if per_row else None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg[dom], "kernel_ms_per_step": dom_ms,
                "note": "search kernels are FP32/ALU-issue bound (distance evaluation + selection), not HBM bound; the gather-bound feature kernels are the ones to read against the HBM roofline (all_kernels); see DESIGN.md",
                "all_kernels": {nme: {"ms": kernels[nme]["ms_per_step"],
                                      "achieved_gbs": (alg[nme] / (kernels[nme]["ms_per_step"] * 1e-3) / 1e9) if nme in alg and kernels[nme]["ms_per_step"] > 0 else None}
                                for nme in kernels}}
    del out

    # ---- end to end through the drop-in module with host buffers --------------------------------
    e2e = None
    if not args.no_e2e:
        e2e_steps = args.e2e_steps or args.steps
        hx = torch.from_numpy(xyz).pin_memory().numpy()             # pinned host input
        if world == 1 or weak:
            hq = hx
        else:                                                       # the same spatial slab as the device leg
            rows_t = shard.slab_queries(t, rank, world)[0]
            hq = torch.from_numpy(xyz[rows_t.cpu().numpy()]).pin_memory().numpy()
        rows = hq.shape[0]
        e2e_block = max(1, min(rows, block_rows, int(os.environ.get("PGEOF_BENCH_E2E_BLOCK", "16000000"))))

        def host_step():
            last, nnz_h = 0.0, 0
            for lo in range(0, rows, e2e_block):
                qb = hq if e2e_block >= rows else hq[lo:lo + e2e_block]
                if search[0] == "knn":
                    knn, _d2 = pgeof.knn_search(hx, qb, search[1])          # numpy in -> numpy out (H2D + D2H inside)
                    nn_ptr = np.arange(0, (qb.shape[0] + 1) * search[1], search[1], dtype=np.uint32)   # README glue (README.md:135-141) in one numpy pass
                    nn = knn.reshape(-1)                                    # zero-copy: knn is already uint32
                else:
                    ridx, _d2 = pgeof.radius_search(hx, qb, search[1], search[2])
                    nn_ptr = np.r_[0, (ridx >= 0).sum(axis=1).cumsum()].astype(np.uint32)              # README.md:157-163
                    nn = ridx[ridx >= 0].astype(np.uint32)
                nnz_h += nn.shape[0]
                f = features_of(hx, nn, nn_ptr)
                last = float(f.reshape(-1)[0])                              # read the result on the host
            return last, nnz_h

        nnz_h = 0
        for _ in range(2):
            _, nnz_h = host_step()
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
        n_blocks = (rows + e2e_block - 1) // e2e_block
        out_floats = {"features": 11, "multiscale": 11 * (len(feat[1]) if feat[0] == "multiscale" else 1), "optimal": 12}[feat[0]]
        h2d = n_blocks * (n * 12 + n * 12) + (0 if hq is hx else rows * 12) + nnz_h * 4 + (rows + n_blocks) * 4
        d2h = rows * kk * 8 + rows * out_floats * 4
        e2e = {"value": n_total * e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": 1e3 * dt / e2e_steps, "steps": e2e_steps,
               "step_ms_min_median_max": [min(e2e_ms), statistics.median(e2e_ms), max(e2e_ms)],
               "path": "pgeof search (numpy) -> README CSR glue in numpy -> pgeof features (numpy); pinned host input, pinned pooled outputs; %d query block(s) per step" % n_blocks}

    # ---- CPU baseline beside it (rank 0, N = 1 only) ---------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import cpu
        cpu.build()
        n_s = cfg["cpu_sample"]
        cx = make_cloud(synth, cfg, n_s, 0, n)
        cpu_step(cpu, cfg, cx)
        reps, t0 = 0, time.perf_counter()
        while reps < 3 or time.perf_counter() - t0 < 10.0:
            cpu_step(cpu, cfg, cx)
            reps += 1
            if time.perf_counter() - t0 > 30.0:
                break
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": n_s * reps / dt, "unit": UNIT, "cores": cpu.hardware_threads(), "kind": "port",
                        "sample": sample_text(cfg, n_s) + "; %d repetitions" % reps}

    if rank == 0:
        line = {"metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": cfg["workload"], "name": args.config, "points": n_total, "rows_per_rank": rows_dev, "neighbours_per_rank": nnz,
                           "parallelism": "query-sharded x%d (z slabs of ~n/N points from a histogram of the replicated cloud, computed inside the step by the library), cloud replicated, grid clipped to the slab, no data-path collective; the e2e leg uses the same slabs" % world,
                           "l2": "512 MiB buffer zeroed between timed iterations; per-step working set %.1f GB >> 126 MB L2" % ((nnz * 12 + rows_dev * 44 + n * 28) / 1e9),
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
