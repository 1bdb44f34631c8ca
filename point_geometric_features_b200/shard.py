"""Query sharding across the GPUs of one box.

Every query row depends only on the read-only cloud (``nn_search.hpp:51-58``,
``pgeof.hpp:95-108``), so the path shards with NO data-path collective: the cloud is
replicated, rank r owns the contiguous query range ``shard_range(n, r, world)``, builds its
own grid and writes its own rows.  A CSR over a shard uses LOCAL offsets, which the reference
explicitly allows (``pgeof.hpp:83``: "number of points is not determined by xyz") and which the
uint32 ``nn_ptr`` requires beyond 2^32-1 neighbours.

Which rows a rank owns is a free choice.  ``shard_range`` (contiguous row ranges) is zero-copy but, on a
cloud in random order, leaves every rank with queries spread thinly over the WHOLE volume: the kNN tile
kernel shares one candidate region between 32 neighbouring queries, so its cost per query grows as the
queries thin out (measured: 5 M of 10 M random rows cost 7.1 ms against 7.7 ms for all 10 M).
``slab_queries`` / ``spatial_shard`` therefore give rank r a SLAB along one axis holding ~n/world points (edges from
a 4096-bin histogram of the replicated cloud, computed identically on every rank, no collective), which keeps the
query density of a single-GPU run; ``gather_rows_indexed`` puts such row blocks back in input order.

``gather_rows`` / ``gather_rows_indexed`` (optional, off the timed path) reassemble per-rank row blocks
with ``torch.distributed.all_gather`` -- NCCL on GPUs, gloo in the CPU tests.
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


SLAB_BINS = 4096


def slab_queries(xyz, rank, world, axis=2):
    """``(rows, query)`` of rank ``rank``'s slab along ``axis``: row ids (ascending int64) and their coordinates.

    The slabs partition the rows.  Their edges come from a 4096-bin histogram of the replicated cloud (slab r = the bins
    whose exclusive prefix count lies in [r n / world, (r + 1) n / world)), so every slab holds n/world points up to the
    mass of one bin and every rank derives the same edges with no exchange.  CUDA tensors go through the library
    (``pgeof_slab_plan_dev`` / ``pgeof_slab_fill_dev``: three streaming passes and one host synchronisation); host tensors
    use the torch restatement below, bin for bin the same arithmetic (it is what the gloo tests exercise)."""
    import torch

    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    n = xyz.shape[0]
    if xyz.is_cuda:
        from . import slab_select
        return slab_select(xyz, rank, world, axis)
    if n == 0:
        return torch.arange(0), xyz
    c = xyz[:, axis].to(torch.float32)
    lo, hi = c.min(), c.max()
    ext = hi - lo
    scale = torch.tensor(float(SLAB_BINS), dtype=torch.float32) / ext if float(ext) > 0 else torch.tensor(0.0)
    bins = ((c - lo) * scale).floor().clamp(0, SLAB_BINS - 1).to(torch.int64)
    cum = torch.cat([torch.zeros(1, dtype=torch.int64), torch.bincount(bins, minlength=SLAB_BINS).cumsum(0)])
    target = torch.tensor([(rank * n) // world, ((rank + 1) * n) // world])
    b_lo, b_hi = (int(v) for v in torch.searchsorted(cum, target, right=False))     # smallest b with cum[b] >= target
    rows = torch.nonzero((bins >= b_lo) & (bins < b_hi)).squeeze(1)
    return rows, xyz[rows]


def spatial_shard(xyz, rank, world, axis=2):
    """Row ids (ascending, int64) of rank ``rank``'s slab along ``axis`` (see ``slab_queries``)."""
    return slab_queries(xyz, rank, world, axis)[0]


def knn_features_shard(xyz, k, rank, world, k_min=1, spatial=True):
    """knn_search(xyz, xyz[rows], k) -> local CSR -> compute_features for this rank's rows.

    ``xyz`` is the full cloud as a CUDA tensor on this rank's device.  Returns
    ``(rows, indices, sqr_dist, features)`` with the outputs kept sharded; ``rows`` holds the input row of
    every output row (a slab of ``spatial_shard``, or the contiguous ``shard_range`` with ``spatial=False``).
    """
    import torch

    from . import compute_features, knn_search

    if world == 1:
        rows, query = torch.arange(xyz.shape[0], device=xyz.device), xyz
    elif spatial:
        rows, query = slab_queries(xyz, rank, world)
    else:
        lo, hi = shard_range(xyz.shape[0], rank, world)
        rows, query = torch.arange(lo, hi, device=xyz.device), xyz[lo:hi]
    idx, d2 = knn_search(xyz, query, k)
    nn_ptr = local_knn_csr(query.shape[0], k, torch, xyz.device)
    feats = compute_features(xyz, idx.view(-1), nn_ptr, k_min)
    return rows, idx, d2, feats


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


def gather_rows_indexed(local, rows, n_total, dist=None):
    """All-gather row blocks whose input rows are listed in ``rows`` (``spatial_shard``) into one
    (n_total, ...) tensor in input order."""
    import torch

    if dist is None:
        import torch.distributed as dist
    world = dist.get_world_size()
    cnt = torch.tensor([local.shape[0]], device=local.device, dtype=torch.int64)
    cnts = [torch.empty_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    sizes = [int(c.item()) for c in cnts]
    pad = max(sizes + [1])
    buf = local.new_zeros((pad,) + tuple(local.shape[1:]))
    buf[: local.shape[0]] = local
    ids = rows.new_zeros((pad,))
    ids[: rows.shape[0]] = rows
    parts = [torch.empty_like(buf) for _ in range(world)]
    id_parts = [torch.empty_like(ids) for _ in range(world)]
    dist.all_gather(parts, buf)
    dist.all_gather(id_parts, ids)
    out = local.new_zeros((n_total,) + tuple(local.shape[1:]))
    for p, i, s in zip(parts, id_parts, sizes):
        out[i[:s]] = p[:s]
    return out
