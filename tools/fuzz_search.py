#!/usr/bin/env python
"""Randomised exactness check of knn_search / radius_search against a brute-force torch evaluation of the defined
float32 metric fl(fl(dx^2 + dy^2) + dz^2) with (d2, index) order.  Run on a GPU box:  python tools/fuzz_search.py [seconds]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pgeof

dev = torch.device("cuda", 0)


def cloud(rng, n, kind):
    if kind == "uniform":
        return rng.uniform(0, rng.choice([1.0, 50.0, 200.0]), (n, 3))
    if kind == "clusters":
        c = rng.uniform(0, 30, (max(1, n // 500), 3))
        return c[rng.integers(0, len(c), n)] + rng.normal(0, rng.choice([0.01, 0.2]), (n, 3))
    if kind == "sheet":
        return np.c_[rng.uniform(0, 20, (n, 2)), rng.normal(0, 0.005, n)]
    if kind == "lattice":
        m = int(round(n ** (1 / 3))) + 1
        g = np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3)[:n].astype(float)
        return g[rng.permutation(len(g))]
    if kind == "dups":
        base = rng.uniform(0, 5, (max(1, n // 20), 3))
        return base[rng.integers(0, len(base), n)]
    if kind == "mixed":
        a = cloud(rng, n // 2, "sheet"); b = cloud(rng, n - n // 2, "uniform") * 0.1
        return np.concatenate([a, b])[rng.permutation(n)]
    raise ValueError(kind)


def brute(data, query, chunk=1024):
    """yields (rows, d2 sorted, idx sorted) with (d2, index) order; all float32 ops separate (no FMA)"""
    n = data.shape[0]
    for s in range(0, query.shape[0], chunk):
        q = query[s:s + chunk]
        dx = q[:, None, 0] - data[None, :, 0]; dy = q[:, None, 1] - data[None, :, 1]; dz = q[:, None, 2] - data[None, :, 2]
        d2 = (dx * dx + dy * dy) + dz * dz
        srt, idx = torch.sort(d2, dim=1, stable=True)            # stable: ties keep ascending index
        yield s, srt, idx


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    t0, it, bad = time.time(), 0, 0
    sizes = {}
    while time.time() - t0 < budget:
        rng = np.random.default_rng(1000 + it)
        kind = rng.choice(["uniform", "clusters", "sheet", "lattice", "dups", "mixed"])
        n = int(rng.choice([int(x) for x in os.environ.get("FUZZ_N", "1,7,100,3000,40000,150000").split(",")]))
        xyz = cloud(rng, n, kind).astype(np.float32)
        n = len(xyz)
        data = torch.from_numpy(xyz).to(dev)
        u = rng.random()
        if u < 0.35:
            query = data
        elif u < 0.6:
            m = int(rng.integers(1, min(n, 5000) + 1))
            qn = xyz[rng.integers(0, n, m)] + rng.normal(0, 0.05, (m, 3)).astype(np.float32) * (rng.random() < 0.7)
            query = torch.from_numpy(qn.astype(np.float32)).to(dev)
        else:
            # spatially local queries (a slab / a blob of the cloud, sometimes pushed partly outside it): the library then
            # indexes only the neighbourhood of the query box and must hand balls that outgrow it to the full grid
            ax = int(rng.integers(0, 3))
            c = xyz[:, ax]
            lo, hi = np.quantile(c, sorted(rng.uniform(0, 1, 2)))
            sel = np.nonzero((c >= lo) & (c <= hi))[0]
            if rng.random() < 0.5 and len(sel):
                ctr = xyz[rng.choice(sel)]
                d = np.linalg.norm(xyz - ctr, axis=1)
                sel = np.argsort(d)[: max(1, int(rng.integers(1, max(2, n // 10))))]
            if len(sel) == 0:
                sel = np.arange(min(n, 10))
            sel = sel[: n // 2] if n >= 4 else sel
            qn = xyz[sel].copy()
            if rng.random() < 0.3:
                qn += rng.normal(0, 1, 3).astype(np.float32) * float(np.ptp(xyz, axis=0).max()) * float(rng.choice([0.01, 0.3, 2.0]))
            query = torch.from_numpy(np.ascontiguousarray(qn, dtype=np.float32)).to(dev)
        rows = torch.from_numpy(rng.choice(query.shape[0], min(query.shape[0], 1500), replace=False)).to(dev)
        k = int(min(n, rng.choice([1, 2, 5, 16, 20, 31, 32, 33, 50, 52, 53, 60, 64, 65, 100])))
        idx, d2 = pgeof.knn_search(data, query, k)
        r_ok = True
        radius = float(rng.choice([0.0, 0.01, 0.05, 0.2, 1.0, 3.0]))
        max_k = int(min(n, rng.choice([1, 4, 16, 32, 33, 64, 100])))
        ridx, rd2 = pgeof.radius_search(data, query, radius, max_k)
        r2 = np.float32(radius) * np.float32(radius)
        for s, srt, sidx in brute(data, query[rows]):
            e = min(s + 1024, rows.shape[0])
            rr = rows[s:e]
            ok = bool((idx.view(torch.int32)[rr] == sidx[:, :k].to(torch.int32)).all()) and bool((d2[rr] == srt[:, :k]).all())
            cnt = (srt < float(r2)).sum(1).clamp(max=max_k)
            col = torch.arange(max_k, device=dev)[None, :]
            hit = col < cnt[:, None]
            want_i = torch.where(hit, sidx[:, :max_k].to(torch.int32), torch.full_like(sidx[:, :max_k], -1, dtype=torch.int32))
            want_d = torch.where(hit, srt[:, :max_k], torch.zeros_like(srt[:, :max_k]))
            r_ok = bool((ridx[rr] == want_i).all()) and bool((rd2[rr] == want_d).all())
            if not (ok and r_ok):
                bad += 1
                print("MISMATCH it=%d kind=%s n=%d k=%d knn_ok=%s radius=%g max_k=%d radius_ok=%s" % (it, kind, n, k, ok, radius, max_k, r_ok), flush=True)
                break
        it += 1
        sizes[n] = sizes.get(n, 0) + 1
    print("fuzz: %d cases, %d mismatches, %.0f s; cases per cloud size: %s" % (it, bad, time.time() - t0, dict(sorted(sizes.items()))))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
