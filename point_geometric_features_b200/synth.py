"""Seeded synthetic point clouds for the BASELINE.json configurations.

The reference's own generators are unseeded (``tests/helpers.py:7``, ``tests/bench_knn.py:8-11``);
these reproduce their distributions with a fixed seed so that parity runs are repeatable.
"""
from __future__ import annotations

import numpy as np


def uniform_cloud(n, seed=0, extent=200.0):
    """``rng.uniform(0, extent, (n, 3)).astype(float32)`` -- tests/bench_knn.py:11, tests/helpers.py:8."""
    rng = np.random.default_rng(seed)
    return rng.uniform(0.0, extent, size=(n, 3)).astype(np.float32)


def lidar_like_cloud(n, seed=0, dup_fraction=1e-3, extent=140.0):
    """LiDAR-like scene of SURVEY.md 8d: 140 m x 140 m x 30 m, float32.

    55 % undulating ground with 2 cm noise, 20 % vertical walls (1 cm thick), 5 % poles
    (r = 0.1 m cylinders, 6 m tall), 20 % Gaussian scatter blobs, plus ``dup_fraction`` exact
    duplicates so that d2 = 0 ties are exercised.

    ``extent`` shrinks the footprint (positions, wall lengths, object counts) while keeping thicknesses,
    noise and heights: ``extent = 140 * sqrt(n / 1e7)`` gives a small cloud the point density -- hence
    the neighbourhood sizes at r = 0.2 -- of the 10 M-point configuration (BASELINE.json configs[2]).
    """
    rng = np.random.default_rng(seed)
    s = float(extent) / 140.0
    n_dup = int(n * dup_fraction)
    m = n - n_dup
    n_ground = int(0.55 * m)
    n_wall = int(0.20 * m)
    n_pole = int(0.05 * m)
    n_scat = m - n_ground - n_wall - n_pole
    parts = []
    # ground
    gx = rng.uniform(0, 140 * s, n_ground)
    gy = rng.uniform(0, 140 * s, n_ground)
    gz = 0.5 * np.sin(gx / 15.0) + 0.3 * np.cos(gy / 11.0) + rng.normal(0, 0.02, n_ground)
    parts.append(np.stack([gx, gy, gz], 1))
    # walls
    n_walls = max(4, int(round(40 * s * s)))
    wid = rng.integers(0, n_walls, n_wall)
    wx0, wy0 = rng.uniform(10 * s, 130 * s, n_walls), rng.uniform(10 * s, 130 * s, n_walls)
    yaw = rng.uniform(0, np.pi, n_walls)
    wlen, wh = rng.uniform(10, 30, n_walls) * min(1.0, max(s, 0.3)), rng.uniform(3, 12, n_walls)
    t = rng.uniform(0, 1, n_wall) * wlen[wid]
    off = rng.normal(0, 0.01, n_wall)
    wx = wx0[wid] + t * np.cos(yaw[wid]) - off * np.sin(yaw[wid])
    wy = wy0[wid] + t * np.sin(yaw[wid]) + off * np.cos(yaw[wid])
    wz = rng.uniform(0, 1, n_wall) * wh[wid]
    parts.append(np.stack([wx, wy, wz], 1))
    # poles
    n_poles = max(10, int(round(400 * s * s)))
    pid = rng.integers(0, n_poles, n_pole)
    px0, py0 = rng.uniform(5 * s, 135 * s, n_poles), rng.uniform(5 * s, 135 * s, n_poles)
    ang = rng.uniform(0, 2 * np.pi, n_pole)
    parts.append(np.stack([px0[pid] + 0.1 * np.cos(ang), py0[pid] + 0.1 * np.sin(ang), rng.uniform(0, 6, n_pole)], 1))
    # scatter
    n_blobs = max(20, int(round(2000 * s * s)))
    bid = rng.integers(0, n_blobs, n_scat)
    bc = np.stack([rng.uniform(5 * s, 135 * s, n_blobs), rng.uniform(5 * s, 135 * s, n_blobs), rng.uniform(2, 10, n_blobs)], 1)
    bs = rng.uniform(0.5, 1.5, n_blobs)
    parts.append(bc[bid] + rng.normal(0, 1, (n_scat, 3)) * bs[bid, None])
    xyz = np.concatenate(parts, 0)
    np.clip(xyz[:, 0], 0, 140 * s, out=xyz[:, 0])
    np.clip(xyz[:, 1], 0, 140 * s, out=xyz[:, 1])
    np.clip(xyz[:, 2], -2, 28, out=xyz[:, 2])
    if n_dup:
        xyz = np.concatenate([xyz, xyz[rng.integers(0, m, n_dup)]], 0)
    xyz = xyz[rng.permutation(xyz.shape[0])]
    return np.ascontiguousarray(xyz, dtype=np.float32)


def knn_csr(knn_idx):
    """CSR view of a dense (n, k) kNN index array, exactly the README glue (README.md:135-141)."""
    n, k = knn_idx.shape
    nn_ptr = (np.arange(n + 1, dtype=np.uint64) * k).astype(np.uint32)
    return knn_idx.reshape(-1), nn_ptr


def radius_csr(rad_idx):
    """CSR from a -1 padded radius result, README.md:157-163."""
    nn_ptr = np.r_[0, (rad_idx >= 0).sum(axis=1).cumsum()].astype(np.uint32)
    nn = rad_idx[rad_idx >= 0].astype(np.uint32)
    return nn, nn_ptr
