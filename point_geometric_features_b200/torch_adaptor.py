"""torch adaptor: FRNN-shaped fixed-radius search and SuperPoint-Transformer style kNN helpers on top of the
B200 search kernels.

The reference names this use itself: ``radius_search`` "could be a fallback replacement for FRNN into
SuperPointTransformer code base" (``include/nn_search.hpp:72``; README credits).  That code base calls

    distances, neighbors, _, _ = frnn.frnn_grid_points(xyz_query[None], xyz_search[None], K=k, r=r_max)

on CUDA tensors and handles batches either through FRNN's padded ``(N, P, 3)`` + ``lengths`` layout or through a
sorted ``batch`` vector.  Both conventions are offered here with the same argument names and return shapes; every
search is one ``radius_search`` / ``knn_search`` call of this package per cloud (device tensors stay on the device,
results are exact and ordered by (squared distance, index)).  torch is plumbing only: slicing, padding, offsets.
"""
from __future__ import annotations


def _search_one(search, query, K, r):
    """max K nearest points of ``search`` within ``r`` of every ``query`` row -> (idx int64 (Q, K) padded -1, d2 (Q, K) padded -1)."""
    import torch

    from . import radius_search

    Q = query.shape[0]
    idx = torch.full((Q, K), -1, dtype=torch.int64, device=query.device)
    d2 = torch.full((Q, K), -1.0, dtype=torch.float32, device=query.device)
    if Q == 0 or search.shape[0] == 0 or K == 0:
        return idx, d2
    kk = min(K, search.shape[0])                                  # max_knn > len(data) is an error in pgeof (nn_search.hpp:92-95)
    i, d = radius_search(search.contiguous(), query.contiguous(), float(r), int(kk))
    i = torch.as_tensor(i, device=query.device).to(torch.int64)
    d = torch.as_tensor(d, device=query.device)
    miss = i < 0
    idx[:, :kk] = i
    d2[:, :kk] = torch.where(miss, torch.full_like(d, -1.0), d)   # pgeof pads distances with 0, FRNN with -1
    return idx, d2


def frnn_grid_points(points1, points2, lengths1=None, lengths2=None, K=-1, r=-1.0, grid=None, return_nn=False,
                     return_sorted=True, radius_cell_ratio=2.0):
    """Drop-in for ``frnn.frnn_grid_points``: for every point of ``points1`` (N, P1, 3) the K nearest points of
    ``points2`` (N, P2, 3) within radius ``r`` (per cloud: scalar, or a tensor of N radii).

    Returns ``(dists, idxs, nn, grid)``: squared distances (N, P1, K) and indices into ``points2[n]`` (N, P1, K), both
    padded with -1 where a point has fewer than K neighbours within ``r`` or lies beyond ``lengths1``; ``nn`` (N, P1, K, 3)
    when ``return_nn``; ``grid`` is always None (the uniform grid is rebuilt per call: 0.7 ms for 10 M points).
    Results are always sorted by distance (``return_sorted`` is accepted for compatibility)."""
    import torch

    if points1.dim() != 3 or points2.dim() != 3 or points1.shape[2] != 3 or points2.shape[2] != 3:
        raise ValueError("points1 and points2 must have shape (N, P, 3)")
    if points1.shape[0] != points2.shape[0]:
        raise ValueError("points1 and points2 must hold the same number of clouds")
    if K < 0 or (not torch.is_tensor(r) and r < 0):
        raise ValueError("K and r must be given")
    N, P1 = points1.shape[0], points1.shape[1]
    p1, p2 = points1.to(torch.float32), points2.to(torch.float32)
    l1 = [P1] * N if lengths1 is None else [int(v) for v in lengths1.tolist()]
    l2 = [points2.shape[1]] * N if lengths2 is None else [int(v) for v in lengths2.tolist()]
    radii = [float(v) for v in r.reshape(-1).tolist()] if torch.is_tensor(r) else [float(r)] * N
    if len(radii) == 1:
        radii = radii * N
    dists = torch.full((N, P1, K), -1.0, dtype=torch.float32, device=points1.device)
    idxs = torch.full((N, P1, K), -1, dtype=torch.int64, device=points1.device)
    for n in range(N):
        i, d = _search_one(p2[n, :l2[n]], p1[n, :l1[n]], K, radii[n])
        idxs[n, :l1[n]] = i
        dists[n, :l1[n]] = d
    nn = None
    if return_nn:
        gather = idxs.clamp(min=0)
        nn = torch.gather(p2[:, None].expand(-1, P1, -1, -1), 2, gather[..., None].expand(-1, -1, -1, 3))
        nn = torch.where((idxs < 0)[..., None], torch.zeros_like(nn), nn)
    return dists, idxs, nn, None


def _segments(batch, n):
    """[lo, hi) of every batch item of a SORTED batch vector (None: one item)."""
    import torch

    if batch is None:
        return [(0, n)]
    if batch.numel() != n:
        raise ValueError("batch must hold one entry per point")
    if n and bool((batch[1:] < batch[:-1]).any()):
        raise ValueError("batch must be sorted")
    counts = torch.bincount(batch.to(torch.int64)).tolist() if n else []
    out, lo = [], 0
    for c in counts:
        out.append((lo, lo + c))
        lo += c
    return out


def knn_2(x_search, x_query, k, r_max=1.0, batch_search=None, batch_query=None):
    """SuperPoint-Transformer ``knn_2``: k nearest ``x_search`` points within ``r_max`` of every ``x_query`` point, clouds
    separated by the sorted ``batch_*`` vectors.  Returns ``(neighbors (Q, k) int64, distances (Q, k) float32)``: indices
    into ``x_search`` (global, -1 where missing), EUCLIDEAN distances (-1 where missing)."""
    import torch

    if (batch_search is None) != (batch_query is None):
        raise ValueError("give both batch vectors or neither")
    seg_s, seg_q = _segments(batch_search, x_search.shape[0]), _segments(batch_query, x_query.shape[0])
    if len(seg_s) < len(seg_q):
        seg_s = seg_s + [(x_search.shape[0], x_search.shape[0])] * (len(seg_q) - len(seg_s))
    Q = x_query.shape[0]
    neighbors = torch.full((Q, k), -1, dtype=torch.int64, device=x_query.device)
    distances = torch.full((Q, k), -1.0, dtype=torch.float32, device=x_query.device)
    xs, xq = x_search.to(torch.float32), x_query.to(torch.float32)
    for (s0, s1), (q0, q1) in zip(seg_s, seg_q):
        i, d2 = _search_one(xs[s0:s1], xq[q0:q1], k, r_max)
        miss = i < 0
        neighbors[q0:q1] = torch.where(miss, i, i + s0)
        distances[q0:q1] = torch.where(miss, d2, d2.clamp(min=0).sqrt())
    return neighbors, distances


def knn_1(xyz, k, r_max=1.0, batch=None, self_is_neighbor=False):
    """SuperPoint-Transformer ``knn_1``: k nearest neighbours of every point within its own cloud (``batch``: sorted cloud
    ids) and within ``r_max``.  With ``self_is_neighbor=False`` the point itself is searched for and dropped, as SPT does."""
    kk = k if self_is_neighbor else k + 1
    neighbors, distances = knn_2(xyz, xyz, kk, r_max, batch, batch)
    if self_is_neighbor:
        return neighbors, distances
    # the point itself comes first (distance 0, lowest index among exact duplicates): drop column 0
    return neighbors[:, 1:].contiguous(), distances[:, 1:].contiguous()
