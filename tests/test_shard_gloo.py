"""world_size-2 gloo test of the multi-GPU host logic: contiguous query shards,
shard-local CSR offsets, optional all-gather of row blocks.  The per-shard compute is done by the
CPU oracle here (no GPU in this container); what is under test is the partition + gather plumbing
of point_geometric_features_b200/shard.py, which the N-GPU bench uses unchanged over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest

from point_geometric_features_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 10, 1000003):
        for world in (1, 2, 3, 8):
            edges = [shard.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1 and sizes == shard.shard_sizes(n, world)
    with pytest.raises(ValueError):
        shard.shard_range(10, 2, 2)


def _worker(rank, world, port, n, k, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import cpu
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    xyz = np.random.default_rng(0).uniform(0, 50, (n, 3)).astype(np.float32)      # replicated cloud
    lo, hi = shard.shard_range(n, rank, world)
    idx, _ = cpu.knn_search(xyz, xyz[lo:hi], k)                                    # this rank's rows only
    nn_ptr = shard.local_knn_csr(hi - lo, k, torch, "cpu").numpy()                 # shard-local offsets
    feats = cpu.compute_features(xyz, idx.reshape(-1), nn_ptr, 1, f64=False)
    full = shard.gather_rows(torch.from_numpy(feats), n, dist)
    full_idx = shard.gather_rows(torch.from_numpy(idx.astype(np.int64)), n, dist)
    # spatial shards (z slabs): same pipeline on the rows spatial_shard hands this rank
    rows = shard.spatial_shard(torch.from_numpy(xyz), rank, world)
    sidx, _ = cpu.knn_search(xyz, xyz[rows.numpy()], k)
    sptr = shard.local_knn_csr(rows.shape[0], k, torch, "cpu").numpy()
    sfeats = cpu.compute_features(xyz, sidx.reshape(-1), sptr, 1, f64=False)
    sfull = shard.gather_rows_indexed(torch.from_numpy(sfeats), rows, n, dist)
    if rank == 0:
        np.save(os.path.join(out_dir, "feats.npy"), full.numpy())
        np.save(os.path.join(out_dir, "idx.npy"), full_idx.numpy())
        np.save(os.path.join(out_dir, "feats_spatial.npy"), sfull.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_reassemble_to_single_process_result(tmp_path):
    import torch.multiprocessing as mp
    from oracle import cpu
    n, k, world = 1001, 12, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, n, k, str(tmp_path)), nprocs=world, join=True)
    xyz = np.random.default_rng(0).uniform(0, 50, (n, 3)).astype(np.float32)
    idx, _ = cpu.knn_search(xyz, xyz, k)
    nn_ptr = (np.arange(n + 1) * k).astype(np.uint32)
    ref = cpu.compute_features(xyz, idx.reshape(-1), nn_ptr, 1, f64=False)
    np.testing.assert_array_equal(np.load(tmp_path / "idx.npy"), idx.astype(np.int64))
    np.testing.assert_array_equal(np.load(tmp_path / "feats.npy"), ref)
    np.testing.assert_array_equal(np.load(tmp_path / "feats_spatial.npy"), ref)


def test_spatial_shards_partition_the_rows():
    import torch
    x = torch.from_numpy(np.random.default_rng(1).normal(0, 30, (20011, 3)).astype(np.float32))
    for world in (1, 2, 3, 8):
        sels = [shard.spatial_shard(x, r, world) for r in range(world)]
        assert torch.equal(torch.cat(sels).sort().values, torch.arange(x.shape[0]))
        sizes = [int(s.shape[0]) for s in sels]
        assert max(sizes) - min(sizes) <= 3 * 0.4 * x.shape[0] * 8 / 4096 + 1       # balanced up to the mass of a few of the 4096 bins (normal data: 0.4 n 8 sigma / 4096 at the mode)
        for r in range(world - 1):                                              # slabs are ordered along z
            assert x[sels[r], 2].max() <= x[sels[r + 1], 2].min()
