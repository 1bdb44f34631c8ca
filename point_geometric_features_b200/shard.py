"""Query sharding across the GPUs of one box.

Every query row depends only on the read-only cloud (``nn_search.hpp:51-58``,
``pgeof.hpp:95-108``), so the path shards with NO data-path collective: the cloud is
replicated, rank r owns the contiguous query range ``shard_range(n, r, world)``, builds its
own grid and writes its own rows.  A CSR over a shard uses LOCAL offsets, which the reference
explicitly allows (``pgeof.hpp:83``: "number of points is not determined by xyz") and which the
uint32 ``nn_ptr`` requires beyond 2^32-1 neighbours.

``gather_rows`` (optional, off the timed path) reassembles per-rank row blocks with
``torch.distributed.all_gather`` -- NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations


def shard_range(n, rank, world):
    """Contiguous, balanced split of ``n`` rows: the first ``n % world`` ranks get one extra row."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n, world):
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def local_knn_csr(n_local, k, torch, device):
    """(nn_ptr) of a dense (n_local, k) kNN block with shard-local offsets (zero-copy nn = idx.view(-1))."""
    if n_local * k > 0xFFFFFFFF:
        raise ValueError("shard holds more than 2^32-1 neighbours; use more shards")
    return (torch.arange(n_local + 1, device=device, dtype=torch.int64) * k).to(torch.uint32)


def knn_features_shard(xyz, k, rank, world, k_min=1):
    """knn_search(xyz, xyz[lo:hi], k) -> local CSR -> compute_features for this rank's rows.

    ``xyz`` is the full cloud as a CUDA tensor on this rank's device.  Returns
    ``(lo, hi, indices, sqr_dist, features)`` with the outputs kept sharded.
    """
    import torch

    from . import compute_features, knn_search

    lo, hi = shard_range(xyz.shape[0], rank, world)
    query = xyz if world == 1 else xyz[lo:hi]
    idx, d2 = knn_search(xyz, query, k)
    nn_ptr = local_knn_csr(hi - lo, k, torch, xyz.device)
    feats = compute_features(xyz, idx.view(-1), nn_ptr, k_min)
    return lo, hi, idx, d2, feats


def gather_rows(local, n_total, dist=None):
    """All-gather row blocks of unequal length (``shard_range`` order) into one (n_total, ...) tensor."""
    import torch

    if dist is None:
        import torch.distributed as dist
    world = dist.get_world_size()
    sizes = shard_sizes(n_total, world)
    pad = max(sizes)
    buf = local.new_zeros((pad,) + tuple(local.shape[1:]))
    buf[: local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], 0)
