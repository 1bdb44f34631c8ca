"""CPU tests that pin the oracle: against scipy exactly as the reference's own
tests do (tests/test_pgeof.py:8-46 of the reference), against LAPACK, against analytic known
answers and against the committed golden fixtures."""
import os

import numpy as np
import pytest
from scipy.spatial import KDTree

from oracle import cpu, ref_numpy as rn
from tests.helpers import compare_features, knn_csr, radius_csr, row_eigvals

GOLD = os.path.join(os.path.dirname(__file__), "golden", "readme_600.npz")


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_knn_oracles_match_scipy(seed):
    """reference tests/test_pgeof.py:8-15 (1000 pts uniform [0,200)^3, k=10), seeded."""
    xyz = np.random.default_rng(seed).uniform(0.0, 200.0, size=(1000, 3)).astype(np.float32)
    _, k_legacy = KDTree(xyz).query(xyz, k=10, workers=-1)
    for idx, d2 in (rn.knn_search(xyz, xyz, 10), cpu.knn_search(xyz, xyz, 10), cpu.knn_search(xyz, xyz, 10, brute=True)):
        np.testing.assert_equal(k_legacy, idx)
        assert (np.diff(d2, axis=1) >= 0).all()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_radius_oracles_match_scipy(seed):
    """reference tests/test_pgeof.py:18-27 (1000 pts [0,1)^3, r=0.2, max_knn=10), seeded."""
    xyz = np.random.default_rng(seed).random(size=(1000, 3), dtype=np.float32)
    _, k_legacy = KDTree(xyz).query(xyz, k=10, distance_upper_bound=0.2, workers=-1)
    k_legacy[k_legacy == xyz.shape[0]] = -1
    for idx, d2 in (rn.radius_search(xyz, xyz, 0.2, 10), cpu.radius_search(xyz, xyz, 0.2, 10)):
        np.testing.assert_equal(k_legacy, idx)
        assert (d2[idx < 0] == 0).all()


def test_kdtree_equals_bruteforce_with_ties():
    """Duplicates and lattice points: (d2, index) ties must resolve identically in every oracle."""
    rng = np.random.default_rng(3)
    g = np.stack(np.meshgrid(*[np.arange(8)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    xyz = np.concatenate([g, g[rng.integers(0, len(g), 200)], rng.uniform(0, 7, (300, 3)).astype(np.float32)])
    xyz = xyz[rng.permutation(len(xyz))]
    q = np.concatenate([xyz[:300], rng.uniform(-3, 10, (100, 3)).astype(np.float32)])
    for k in (1, 7, 27, 64):
        a = cpu.knn_search(xyz, q, k)
        b = cpu.knn_search(xyz, q, k, brute=True)
        c = rn.knn_search(xyz, q, k)
        np.testing.assert_equal(a[0], b[0]); np.testing.assert_equal(a[1], b[1])
        np.testing.assert_equal(a[0], c[0]); np.testing.assert_equal(a[1], c[1])
    for r, m in ((1.0, 5), (1.5, 40), (0.0, 3)):
        a = cpu.radius_search(xyz, q, r, m)
        c = rn.radius_search(xyz, q, r, m)
        np.testing.assert_equal(a[0], c[0]); np.testing.assert_equal(a[1], c[1])


def test_eig3_restatement_matches_lapack():
    rng = np.random.default_rng(4)
    for _ in range(200):
        p = rng.normal(size=(rng.integers(3, 30), 3)) * rng.uniform(1e-3, 1e3)
        c = np.cov(p.T, bias=True)
        w, v = cpu.eig3(c)
        wl = np.linalg.eigvalsh(c)
        np.testing.assert_allclose(w, wl, rtol=1e-9, atol=1e-12 * abs(wl).max())
        np.testing.assert_allclose(v.T @ v, np.eye(3), atol=1e-12)
        np.testing.assert_allclose(c @ v, v * w, atol=1e-9 * abs(wl).max())
    w, v = cpu.eig3(np.zeros((3, 3)))
    assert (w == 0).all() and (v == np.eye(3)).all()          # Eigen: lambda = 0, V = I
    w32, _ = cpu.eig3(np.diag([3.0, 1.0, 2.0]).astype(np.float32))
    np.testing.assert_allclose(w32, [1, 2, 3])


@pytest.mark.parametrize("order", ["literal", "documented"])
def test_cpp_port_matches_numpy_restatement(order):
    rng = np.random.default_rng(5)
    xyz = rng.uniform(0, 200, (2000, 3)).astype(np.float32)
    idx, _ = cpu.knn_search(xyz, xyz, 30)
    nn, nn_ptr = knn_csr(idx)
    np.testing.assert_allclose(cpu.compute_features(xyz, nn, nn_ptr, 1, order), rn.compute_features(xyz, nn, nn_ptr, 1, order), atol=1e-10)
    np.testing.assert_allclose(cpu.compute_features_multiscale(xyz, nn, nn_ptr, [5, 10, 30], order),
                               rn.compute_features_multiscale(xyz, nn, nn_ptr, [5, 10, 30], order), atol=1e-10)
    a, ma = cpu.compute_features_optimal(xyz, nn, nn_ptr, 1, 2, 8, order, return_margin=True)
    b, mb = rn.compute_features_optimal(xyz, nn, nn_ptr, 1, 2, 8, order, return_margin=True)
    np.testing.assert_allclose(a, b, atol=1e-10)
    np.testing.assert_allclose(ma, mb, atol=1e-12)
    # float arithmetic of the reference stays inside the stated tolerance of the float64 evaluation
    f32 = cpu.compute_features(xyz, nn, nn_ptr, 1, order, f64=False)
    assert np.abs(f32 - cpu.compute_features(xyz, nn, nn_ptr, 1, order)).max() < 1e-4


def test_upstream_multiscale_self_consistency():
    """reference tests/test_pgeof.py:30-46 on the oracle (10000 pts -> 3000 here, k=50)."""
    xyz = np.random.default_rng(6).uniform(0.0, 200.0, size=(3000, 3)).astype(np.float32)
    nn, nn_ptr = knn_csr(KDTree(xyz).query(xyz, k=50, workers=-1)[1])
    multi = cpu.compute_features_multiscale(xyz, nn, nn_ptr, np.flip(np.array([50, 20])), f64=False)
    simple = cpu.compute_features(xyz, nn, nn_ptr, 50, f64=False)
    multi_simple = cpu.compute_features_multiscale(xyz, nn, nn_ptr, [20], f64=False)
    np.testing.assert_allclose(multi[:, 0], multi_simple[:, 0], 1e-1, 1e-5)
    np.testing.assert_allclose(multi[:, 1], simple, 1e-1, 1e-5)


def test_known_answers():
    """SURVEY.md A.6."""
    xyz = np.array([[1, 2, 3], [1, 2, 3], [1, 2, 3], [4, 4, 4]], np.float32)
    nn = np.array([0, 0, 1, 2, 3], np.uint32)
    nn_ptr = np.array([0, 1, 4, 4, 5], np.uint32)
    expect = [0, 0, 0, 0, 0, 0, 1, 0, 1e-3, 1e-3, 0]
    for impl in (rn, cpu):
        f = impl.compute_features(xyz, nn, nn_ptr, 1)
        np.testing.assert_allclose(f[0], expect, atol=1e-12)          # k = 1
        np.testing.assert_allclose(f[1], expect, atol=1e-12)          # coincident neighbours
        assert (f[2] == 0).all()                                       # empty row < k_min
        assert (impl.compute_features(xyz, nn, nn_ptr, 2)[0] == 0).all()   # k_min gate
    # exact plane z = c (documented order): normal (0,0,1), scattering 0, curvature 0
    rng = np.random.default_rng(7)
    plane = np.c_[rng.uniform(0, 1, (40, 2)), np.full(40, 0.5)].astype(np.float32)
    f = cpu.compute_features(plane, np.arange(40, dtype=np.uint32), np.array([0, 40], np.uint32), 1, "documented")[0]
    np.testing.assert_allclose(f[4:7], [0, 0, 1], atol=1e-9)
    assert abs(f[2]) < 1e-9 and abs(f[10]) < 1e-9
    # collinear along x (documented order): planarity 0, linearity = s0 / (s0 + eps)
    line = np.c_[np.linspace(0, 1, 30), np.zeros(30), np.zeros(30)].astype(np.float32)
    f = cpu.compute_features(line, np.arange(30, dtype=np.uint32), np.array([0, 30], np.uint32), 1, "documented")[0]
    assert abs(f[1]) < 1e-9 and abs(f[0] - f[7] / (f[7] + 1e-3)) < 1e-9


def test_error_paths_of_the_oracle():
    xyz = np.zeros((5, 3), np.float32)
    with pytest.raises(ValueError):
        cpu.knn_search(xyz, xyz, 6)                      # nn_search.hpp:37
    with pytest.raises(ValueError):
        cpu.radius_search(xyz, xyz, 1.0, 6)              # nn_search.hpp:92-95
    nn, p = np.zeros(5, np.uint32), np.array([0, 5], np.uint32)
    with pytest.raises(ValueError):
        cpu.compute_features(xyz, nn, p, 0)              # pgeof.hpp:81
    with pytest.raises(ValueError):
        cpu.compute_features_multiscale(xyz, nn, p, [5, 2])   # pgeof.hpp:165
    with pytest.raises(ValueError):
        cpu.compute_features_optimal(xyz, nn, p, 0, 1, 0)     # pgeof.hpp:250


def test_golden_fixture_still_reproduced():
    g = np.load(GOLD)
    xyz = g["xyz"]
    idx, d2 = cpu.knn_search(xyz, xyz, 20)
    np.testing.assert_equal(idx, g["knn_idx"]); np.testing.assert_equal(d2, g["knn_d2"])
    ridx, rd2 = cpu.radius_search(xyz, xyz, 0.2, 10)
    np.testing.assert_equal(ridx, g["radius_idx"]); np.testing.assert_equal(rd2, g["radius_d2"])
    nn, nn_ptr = knn_csr(idx)
    for order in ("literal", "documented"):
        np.testing.assert_allclose(cpu.compute_features(xyz, nn, nn_ptr, 1, order), g["features_" + order], atol=1e-9)
        np.testing.assert_allclose(cpu.compute_features_multiscale(xyz, nn, nn_ptr, [5, 10, 20], order), g["multiscale_" + order], atol=1e-9)
        np.testing.assert_allclose(cpu.compute_features_optimal(xyz, nn, nn_ptr, 1, 1, 5, order), g["optimal_" + order], atol=1e-9)
        ids = g["selected_ids"]
        s32 = cpu.compute_features_selected(xyz, 0.15, 12, ids, order)
        s64 = cpu.compute_features_selected(xyz.astype(np.float64) * 1.000000123, 0.15, 12, ids, order)
        # conditioning aware: rank-deficient 2-3 point balls amplify even float64 round-off 1000x
        # through 1/(s0 + 1e-3) in the literal order
        rnn, rptr = radius_csr(cpu.radius_search(xyz, xyz, 0.15, 12)[0])
        ev = row_eigvals(xyz, rnn, rptr)
        compare_features(s64, g["selected_f64_" + order], ev, order, "selected f64", ids, max_weak=0.05, max_degenerate=0.2, max_ill=0.2)
        stats = compare_features(s32, g["selected_f32_" + order], ev, order, "selected f32", ids, max_weak=0.05, max_degenerate=0.2, max_ill=0.2)
        assert stats["degenerate_rows"] < 0.2 * len(xyz)
    # README glue for a radius result (README.md:157-163)
    nn_r, ptr_r = radius_csr(ridx)
    assert ptr_r[-1] == len(nn_r) == (ridx >= 0).sum()
