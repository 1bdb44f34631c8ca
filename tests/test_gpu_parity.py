"""GPU parity tests: the CUDA path, called through the drop-in Python surface and the C ABI, against
the CPU oracle on the same seeded inputs.  Bars (BASELINE.json:north_star): neighbour indices and
squared distances bit-exact with (d2, index) tie-breaking; optimal-k exact; the 11 features within
1e-4 absolute / 1e-3 relative (tests/helpers.py states the conditioning-aware exemptions)."""
import ctypes
import os

import numpy as np
import pytest
from scipy.spatial import KDTree

import pgeof
import point_geometric_features_b200 as b200
from oracle import cpu, ref_numpy as rn
from point_geometric_features_b200 import synth
from tests.helpers import compare_features, knn_csr, radius_csr, row_eigvals, row_eigvals_dense

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "readme_600.npz")
ORDERS = ("literal", "documented")
# bounds on the rows compare_features may treat with a relaxed tolerance (fractions of the rows; tests/helpers.py):
# uniform clouds have essentially none, the LiDAR-like scene has thin walls / ground (near rank-deficient) and 2-3 point balls
UNIFORM = dict(max_weak=2e-3, max_degenerate=0.0, max_ill=1e-3)
LIDAR = dict(max_weak=0.02, max_degenerate=0.2, max_ill=0.15)


def dense_lidar(n, seed):
    """LiDAR-like scene with the footprint shrunk so that n points have the density of the 10 M-point config (C3)."""
    return synth.lidar_like_cloud(n, seed=seed, extent=140.0 * (n / 1e7) ** 0.5)


@pytest.fixture(autouse=True)
def _literal_order():
    b200.set_eig_order("literal")
    yield
    b200.set_eig_order("env")


def _assert_search_equal(got, ref):
    np.testing.assert_array_equal(got[0], ref[0])
    np.testing.assert_array_equal(got[1].view(np.uint32), ref[1].view(np.uint32))   # bit-exact float32


# --------------------------------------------------------------------------------------------
# the reference's own three tests (tests/test_pgeof.py:8-46), seeded, against the new module
# --------------------------------------------------------------------------------------------
def test_upstream_knn():
    xyz = np.random.default_rng(0).uniform(0.0, 200.0, size=(1000, 3)).astype(np.float32)
    _, k_legacy = KDTree(xyz).query(xyz, k=10, workers=-1)
    k_new, _ = pgeof.knn_search(xyz, xyz, 10)
    assert k_new.dtype == np.uint32 and k_new.shape == (1000, 10)
    np.testing.assert_equal(k_legacy, k_new)


def test_upstream_radius_search():
    xyz = np.random.default_rng(0).random(size=(1000, 3), dtype=np.float32)
    _, k_legacy = KDTree(xyz).query(xyz, k=10, distance_upper_bound=0.2, workers=-1)
    k_legacy[k_legacy == xyz.shape[0]] = -1
    k_new, d2 = pgeof.radius_search(xyz, xyz, 0.2, 10)
    assert k_new.dtype == np.int32 and d2.dtype == np.float32
    np.testing.assert_equal(k_legacy, k_new)


def test_upstream_multiscale():
    xyz = np.random.default_rng(0).uniform(0.0, 200.0, size=(10000, 3)).astype(np.float32)
    kneigh = KDTree(xyz).query(xyz, k=50, workers=-1)
    nn_ptr = (np.arange(10000 + 1) * 50).astype("uint32")
    nn = np.ascontiguousarray(kneigh[1].flatten().astype("uint32"))
    multi = pgeof.compute_features_multiscale(xyz, nn, nn_ptr, np.flip(np.array([50, 20])), False)
    simple = pgeof.compute_features(xyz, nn, nn_ptr, 50, False)
    multi_simple = pgeof.compute_features_multiscale(xyz, nn, nn_ptr, [20], False)
    assert multi.shape == (10000, 2, 11) and simple.shape == (10000, 11)
    np.testing.assert_allclose(multi[:, 0], multi_simple[:, 0], 1e-1, 1e-5)
    np.testing.assert_allclose(multi[:, 1], simple, 1e-1, 1e-5)


def test_golden_fixture():
    g = np.load(GOLD)
    xyz = g["xyz"]
    _assert_search_equal(pgeof.knn_search(xyz, xyz, 20), (g["knn_idx"], g["knn_d2"]))
    _assert_search_equal(pgeof.radius_search(xyz, xyz, 0.2, 10), (g["radius_idx"], g["radius_d2"]))
    nn, nn_ptr = knn_csr(g["knn_idx"])
    ev = row_eigvals_dense(xyz, g["knn_idx"])
    for order in ORDERS:
        b200.set_eig_order(order)
        compare_features(pgeof.compute_features(xyz, nn, nn_ptr), g["features_" + order], ev, order, **UNIFORM)
        ms = pgeof.compute_features_multiscale(xyz, nn, nn_ptr, [5, 10, 20])
        for s, ks in enumerate((5, 10, 20)):
            compare_features(ms[:, s], g["multiscale_" + order][:, s], row_eigvals_dense(xyz, g["knn_idx"][:, :ks]), order, "scale %d" % ks,
                             max_weak=0.02, max_degenerate=0.0, max_ill=1e-3)     # 5-point neighbourhoods: a few near-planar ones
        opt = pgeof.compute_features_optimal(xyz, nn, nn_ptr, 1, 1, 5)
        sure = g["optimal_margin_" + order] > 1e-9
        np.testing.assert_array_equal(opt[sure, 11], g["optimal_" + order][sure, 11])


# --------------------------------------------------------------------------------------------
# neighbour search: bit-exact against the (d2, index) oracle
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k", [1, 2, 10, 20, 32, 33, 50, 64, 65, 100, 128, 129, 200, 256, 300, 512, 513, 1000, 3000])
def test_knn_uniform_bitexact(k):
    xyz = synth.uniform_cloud(20000, seed=k)
    _assert_search_equal(pgeof.knn_search(xyz, xyz, k), cpu.knn_search(xyz, xyz, k))


def test_knn_c2_subsample_bitexact():
    """BASELINE config 2 (1M uniform pts, k=50): 1M on the GPU, oracle on the full cloud for 20k sampled queries."""
    xyz = synth.uniform_cloud(1_000_000, seed=0)
    idx, d2 = pgeof.knn_search(xyz, xyz, 50)
    rows = np.random.default_rng(1).choice(len(xyz), 20000, replace=False)
    ref = cpu.knn_search(xyz, xyz[rows], 50)
    _assert_search_equal((idx[rows], d2[rows]), ref)
    assert (idx[:, 0] == np.arange(len(xyz))).all() and (d2[:, 0] == 0).all()      # self first: d2 = 0, lowest index
    assert (np.diff(d2, axis=1) >= 0).all()                                       # rows ascending


def test_knn_query_differs_from_data_and_leaves_bbox():
    rng = np.random.default_rng(2)
    data = synth.uniform_cloud(30000, seed=3, extent=50.0)
    query = np.concatenate([rng.uniform(0, 50, (2000, 3)), rng.uniform(-40, 90, (1500, 3)), rng.uniform(-1e4, 1e4, (50, 3)),
                            data[:100]]).astype(np.float32)
    for k in (1, 16, 50):
        _assert_search_equal(pgeof.knn_search(data, query, k), cpu.knn_search(data, query, k))


def test_knn_ties_duplicates_lattice_and_degenerate_shapes():
    rng = np.random.default_rng(4)
    g = np.stack(np.meshgrid(*[np.arange(12)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    lattice = np.concatenate([g, g[rng.integers(0, len(g), 500)]])[rng.permutation(len(g) + 500)]     # exact ties + duplicates
    heavy_dup = np.concatenate([np.tile(np.float32([[1, 2, 3]]), (700, 1)), rng.uniform(0, 5, (300, 3)).astype(np.float32)])
    plane = np.c_[rng.uniform(0, 10, (5000, 2)), np.zeros(5000)].astype(np.float32)                   # zero z extent
    line = np.c_[rng.uniform(0, 10, 3000), np.full(3000, 2.0), np.full(3000, -1.0)].astype(np.float32)
    single = np.float32([[3, 3, 3]])
    for name, xyz, ks in (("lattice", lattice, (1, 7, 27, 50)), ("heavy_dup", heavy_dup, (5, 50, 300)), ("plane", plane, (8, 50)),
                          ("line", line, (8, 50)), ("single", single, (1,)), ("tiny", lattice[:5], (5,))):
        for k in ks:
            got, ref = pgeof.knn_search(xyz, xyz, k), cpu.knn_search(xyz, xyz, k, brute=len(xyz) < 4000)
            np.testing.assert_array_equal(got[0], ref[0], err_msg="%s k=%d" % (name, k))
            np.testing.assert_array_equal(got[1], ref[1], err_msg="%s k=%d" % (name, k))


def test_knn_lidar_like_bitexact():
    xyz = synth.lidar_like_cloud(300000, seed=0)
    idx, d2 = pgeof.knn_search(xyz, xyz, 30)
    rows = np.random.default_rng(5).choice(len(xyz), 20000, replace=False)
    _assert_search_equal((idx[rows], d2[rows]), cpu.knn_search(xyz, xyz[rows], 30))


@pytest.mark.parametrize("r,max_knn", [(0.2, 10), (0.05, 10), (0.2, 1), (0.35, 64), (0.3, 100), (0.0, 4), (2.0, 32)])
def test_radius_unit_cube_bitexact(r, max_knn):
    xyz = np.random.default_rng(6).random((3000, 3), dtype=np.float32)
    _assert_search_equal(pgeof.radius_search(xyz, xyz, r, max_knn), cpu.radius_search(xyz, xyz, r, max_knn))


def test_radius_lidar_like_c3_subsample():
    """BASELINE config 3 shape (LiDAR-like, r=0.2, max_k=64) at 400k points."""
    xyz = synth.lidar_like_cloud(400000, seed=1)
    idx, d2 = pgeof.radius_search(xyz, xyz, 0.2, 64)
    rows = np.random.default_rng(7).choice(len(xyz), 20000, replace=False)
    _assert_search_equal((idx[rows], d2[rows]), cpu.radius_search(xyz, xyz[rows], 0.2, 64))
    # CSR emitted directly == README glue on the padded result (README.md:157-163)
    nn, nn_ptr = b200.radius_search_csr(xyz, xyz, 0.2, 64)
    nn_ref, ptr_ref = radius_csr(idx)
    np.testing.assert_array_equal(nn_ptr, ptr_ref)
    np.testing.assert_array_equal(nn, nn_ref)


def test_radius_dense_ball_overflows_candidate_buffer():
    """> 512 points inside the ball: the threshold is bisected by rescans, result still exact."""
    rng = np.random.default_rng(8)
    xyz = np.concatenate([rng.normal(0, 0.05, (3000, 3)), rng.uniform(-2, 2, (2000, 3))]).astype(np.float32)
    q = xyz[:400]
    for max_knn in (16, 64, 200):
        _assert_search_equal(pgeof.radius_search(xyz, q, 0.5, max_knn), cpu.radius_search(xyz, q, 0.5, max_knn))
    _assert_search_equal(pgeof.knn_search(xyz, q, 40), cpu.knn_search(xyz, q, 40))


def test_search_edge_cases():
    xyz = synth.uniform_cloud(100, seed=9)
    idx, d2 = pgeof.knn_search(xyz, xyz[:0], 5)
    assert idx.shape == (0, 5) and d2.shape == (0, 5)
    idx, d2 = pgeof.knn_search(xyz, xyz, 0)
    assert idx.shape == (100, 0)
    idx, d2 = pgeof.radius_search(xyz, xyz, 1e9, 100)                               # everything, max_knn == n
    _assert_search_equal((idx, d2), cpu.radius_search(xyz, xyz, 1e9, 100))
    _assert_search_equal(pgeof.knn_search(xyz, xyz, 100), cpu.knn_search(xyz, xyz, 100))   # knn == n
    with pytest.raises(ValueError):
        pgeof.knn_search(xyz, xyz, 101)
    strided = np.zeros((100, 4), np.float32)[:, :3]                                  # row stride 16 B, unit inner stride
    strided[:] = xyz
    _assert_search_equal(pgeof.knn_search(strided, strided, 7), cpu.knn_search(xyz, xyz, 7))


# --------------------------------------------------------------------------------------------
# features
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("order", ORDERS)
def test_features_c1_readme_example(order):
    """BASELINE config 1: 10k uniform [0,1)^3 float32 points, knn_search k=20 -> CSR -> compute_features."""
    b200.set_eig_order(order)
    xyz = np.random.default_rng(10).random((10000, 3)).astype("float32")
    knn, _ = pgeof.knn_search(xyz, xyz, 20)
    nn_ptr = (np.arange(10000 + 1) * 20).astype("uint32")
    nn = knn.flatten().astype("uint32")
    f = pgeof.compute_features(xyz, nn, nn_ptr)
    assert f.dtype == np.float32 and f.shape == (10000, 11)
    compare_features(f, cpu.compute_features(xyz, nn, nn_ptr, 1, order), row_eigvals_dense(xyz, knn), order, **UNIFORM)


@pytest.mark.parametrize("order", ORDERS)
def test_features_c2_subsample(order):
    b200.set_eig_order(order)
    xyz = synth.uniform_cloud(200000, seed=11)
    idx, _ = pgeof.knn_search(xyz, xyz, 50)
    nn, nn_ptr = knn_csr(idx)
    f = pgeof.compute_features(xyz, nn, nn_ptr)
    stats = compare_features(f, cpu.compute_features(xyz, nn, nn_ptr, 1, order), row_eigvals_dense(xyz, idx), order, **UNIFORM)
    assert stats["ill_conditioned_rows"] < 0.001 * len(xyz) and stats["degenerate_rows"] == 0 and stats["weak_rows"] == 0


@pytest.mark.parametrize("order", ORDERS)
def test_features_ragged_lidar_rows_and_kmin(order):
    b200.set_eig_order(order)
    xyz = dense_lidar(120000, seed=2)                                               # ~40 neighbours within 0.3, as in C3
    idx, _ = pgeof.radius_search(xyz, xyz, 0.3, 48)
    nn, nn_ptr = radius_csr(idx)
    ev = row_eigvals(xyz, nn, nn_ptr)
    for k_min in (1, 5, 30):
        f = pgeof.compute_features(xyz, nn, nn_ptr, k_min)
        ref = cpu.compute_features(xyz, nn, nn_ptr, k_min, order)
        short = np.diff(nn_ptr.astype(np.int64)) < k_min
        assert (f[short] == 0).all()                                                # calloc semantics, pgeof.hpp:88,103
        compare_features(f, ref, ev, order, "k_min=%d" % k_min, **LIDAR)
    # the sparse version of the scene: nearly every ball holds 1-3 points (rank deficient by construction)
    xyz = synth.lidar_like_cloud(60000, seed=2)
    nn, nn_ptr = radius_csr(pgeof.radius_search(xyz, xyz, 0.3, 48)[0])
    compare_features(pgeof.compute_features(xyz, nn, nn_ptr), cpu.compute_features(xyz, nn, nn_ptr, 1, order), row_eigvals(xyz, nn, nn_ptr), order,
                     "sparse scene", max_weak=0.05, max_degenerate=1.0, max_ill=1.0)


def test_features_known_answers_and_errors():
    xyz = np.array([[1, 2, 3], [1, 2, 3], [1, 2, 3], [4, 4, 4]], np.float32)
    nn = np.array([0, 0, 1, 2, 3], np.uint32)
    nn_ptr = np.array([0, 1, 4, 4, 5], np.uint32)
    f = pgeof.compute_features(xyz, nn, nn_ptr)
    expect = np.float32([0, 0, 0, 0, 0, 0, 1, 0, 1e-3, 1e-3, 0])
    np.testing.assert_allclose(f[0], expect, atol=1e-7)
    np.testing.assert_allclose(f[1], expect, atol=1e-7)
    assert (f[2] == 0).all()
    assert (pgeof.compute_features(xyz, nn, nn_ptr, 2)[0] == 0).all()
    b200.set_eig_order("documented")
    rng = np.random.default_rng(12)
    plane = np.c_[rng.uniform(0, 1, (40, 2)), np.full(40, 0.5)].astype(np.float32)
    f = pgeof.compute_features(plane, np.arange(40, dtype=np.uint32), np.array([0, 40], np.uint32))[0]
    np.testing.assert_allclose(f[4:7], [0, 0, 1], atol=1e-6)
    assert abs(f[2]) < 1e-6 and abs(f[10]) < 1e-6
    with pytest.raises(IndexError):                                                 # reference: UB (pca.hpp:124-126)
        pgeof.compute_features(xyz, np.array([0, 9], np.uint32), np.array([0, 2], np.uint32))
    with pytest.raises(IndexError):
        pgeof.compute_features(xyz, nn, np.array([0, 9], np.uint32))
    assert pgeof.compute_features(xyz, nn[:0], np.zeros(1, np.uint32)).shape == (0, 11)


@pytest.mark.parametrize("order", ORDERS)
def test_multiscale_c4_shape(order):
    """BASELINE config 4 shape: k_scales = [10, 20, 50, 100] from one k = 100 kNN pass (60k points here)."""
    b200.set_eig_order(order)
    xyz = synth.uniform_cloud(60000, seed=13)
    idx, _ = pgeof.knn_search(xyz, xyz, 100)
    nn, nn_ptr = knn_csr(idx)
    scales = [10, 20, 50, 100]
    ms = pgeof.compute_features_multiscale(xyz, nn, nn_ptr, scales)
    ref = cpu.compute_features_multiscale(xyz, nn, nn_ptr, scales, order)
    assert ms.shape == (60000, 4, 11)
    for s, ks in enumerate(scales):
        compare_features(ms[:, s], ref[:, s], row_eigvals_dense(xyz, idx[:, :ks]), order, "scale %d" % ks, **UNIFORM)
    # the last scale is the whole row: bitwise the same code path as compute_features
    np.testing.assert_array_equal(ms[:, 3], pgeof.compute_features(xyz, nn, nn_ptr))


def test_multiscale_ragged_rows_many_scales_and_inputs():
    xyz = dense_lidar(50000, seed=3)
    idx, _ = pgeof.radius_search(xyz, xyz, 0.4, 40)
    nn, nn_ptr = radius_csr(idx)
    scales = [1, 2, 3, 5, 8, 8, 13, 21, 30, 40, 64]                                  # > 8 scales: two passes; a duplicate; one never reached
    ms = pgeof.compute_features_multiscale(xyz, nn, nn_ptr, np.asarray(scales, np.int64))
    ref = cpu.compute_features_multiscale(xyz, nn, nn_ptr, scales)
    lens = np.diff(nn_ptr.astype(np.int64))
    for s, ks in enumerate(scales):
        assert (ms[lens < ks, s] == 0).all()                                         # early break, pgeof.hpp:193
        bounds = dict(max_weak=0.05, max_degenerate=1.0, max_ill=1.0) if ks <= 5 else dict(max_weak=0.05, max_degenerate=0.25, max_ill=0.2)
        compare_features(ms[:, s], ref[:, s], row_eigvals(xyz, nn, nn_ptr, ks), "literal", "scale %d" % ks, **bounds)   # k <= 3: rank deficient by construction
    assert (ms[:, -1] == 0).all()


@pytest.mark.parametrize("k_min,k_step,k_min_search", [(1, 1, 10), (1, 1, 1), (5, 3, 10), (20, 7, 4), (1, 200, 10)])
def test_optimal_k_exact(k_min, k_step, k_min_search):
    """BASELINE config 5 shape (k = 100, scan from k_min_search) on 30k points: k_opt must be exact."""
    xyz = synth.uniform_cloud(30000, seed=14)
    idx, _ = pgeof.knn_search(xyz, xyz, 100)
    nn, nn_ptr = knn_csr(idx)
    opt = pgeof.compute_features_optimal(xyz, nn, nn_ptr, k_min, k_step, k_min_search)
    ref, margin = cpu.compute_features_optimal(xyz, nn, nn_ptr, k_min, k_step, k_min_search, return_margin=True)
    assert opt.shape == (30000, 12) and opt.dtype == np.float32
    sure = margin > 1e-9        # below that two neighbourhood sizes tie in float64 itself
    assert sure.mean() > 0.999
    np.testing.assert_array_equal(opt[sure, 11], ref[sure, 11])
    rows = np.nonzero(sure)[0]
    kopt = ref[rows, 11].astype(int)
    ev = np.stack([np.linalg.eigvalsh(np.cov(xyz[idx[r, :k]].astype(np.float64).T, bias=True)) for r, k in zip(rows[:3000], kopt[:3000])])
    # a scan that starts below 5 neighbours mostly ends on 1-3 point neighbourhoods: rank deficient by construction
    bounds = dict(max_weak=1.0, max_degenerate=1.0, max_ill=1.0) if max(k_min, k_min_search) < 5 else dict(max_weak=0.05, max_degenerate=0.0, max_ill=2e-3)
    compare_features(opt[rows[:3000], :11], ref[rows[:3000], :11], ev, "literal", "optimal features", **bounds)


def test_optimal_ragged_rows_and_gates():
    xyz = dense_lidar(40000, seed=4)
    idx, _ = pgeof.radius_search(xyz, xyz, 0.35, 60)
    nn, nn_ptr = radius_csr(idx)
    opt = pgeof.compute_features_optimal(xyz, nn, nn_ptr, 3, 2, 6)
    ref, margin = cpu.compute_features_optimal(xyz, nn, nn_ptr, 3, 2, 6, return_margin=True)
    lens = np.diff(nn_ptr.astype(np.int64))
    assert (opt[lens < 6] == 0).all()                                                # pgeof.hpp:272
    sure = margin > 1e-9
    np.testing.assert_array_equal(opt[sure, 11], ref[sure, 11])


@pytest.mark.parametrize("shape", ["plane", "line", "blobs", "lattice", "tiny", "far", "needle"])
def test_optimal_k_exact_on_degenerate_and_scaled_neighbourhoods(shape):
    """The optimal-k scan filters with closed-form float eigenvalues, which lose digits exactly where eigenvalues coincide or
    vanish; its a-posteriori bound must hand every such case to the double evaluation: k_opt exact (pgeof.hpp:286-289) on
    planes / lines (zero eigenvalues), isotropic blobs and a lattice (coincident eigenvalues), and on clouds scaled or
    shifted by orders of magnitude."""
    rng = np.random.default_rng(33)
    n = 20000
    if shape == "plane":
        xyz = np.c_[rng.uniform(0, 30, (n, 2)), np.zeros(n)]
    elif shape == "line":
        xyz = np.c_[rng.uniform(0, 3000, n), np.full(n, 2.0), np.full(n, -1.0)]
    elif shape == "blobs":
        xyz = rng.normal(0, 1, (n, 3)) + rng.integers(0, 8, (n, 1)) * 10.0
    elif shape == "lattice":
        g = np.arange(28, dtype=np.float64)
        xyz = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)[:n]
    elif shape == "tiny":
        xyz = rng.uniform(0, 1, (n, 3)) * 1e-4
    elif shape == "far":
        xyz = rng.uniform(0, 40, (n, 3)) + 3000.0
    else:                                                                            # thin needles: one large, two small close eigenvalues
        xyz = np.c_[rng.uniform(0, 2000, n), rng.normal(0, 1e-2, n), rng.normal(0, 1e-2, n)]
    xyz = np.ascontiguousarray(xyz, np.float32)
    idx, _ = pgeof.knn_search(xyz, xyz, 60)
    nn, nn_ptr = knn_csr(idx)
    for k_min_search in (3, 10):
        opt = pgeof.compute_features_optimal(xyz, nn, nn_ptr, 1, 1, k_min_search)
        ref, margin = cpu.compute_features_optimal(xyz, nn, nn_ptr, 1, 1, k_min_search, return_margin=True)
        sure = margin > 1e-9
        print("optimal-k %s (k_min_search %d): %.4f of the rows decided in float64 by more than 1e-9" % (shape, k_min_search, sure.mean()))
        np.testing.assert_array_equal(opt[sure, 11], ref[sure, 11])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("order", ORDERS)
def test_selected_fused_radius_features(dtype, order):
    b200.set_eig_order(order)
    E = pgeof.EFeatureID
    ids = [E.Verticality, E.Curvature, E.Eigentropy, E.K_optimal, E.Linearity, E.Planarity, E.Scattering, E.VerticalityPGEOF,
           E.Normal_x, E.Normal_y, E.Normal_z, E.Length, E.Surface, E.Volume]
    xyz = (np.random.default_rng(15).random((20000, 3)) * (1 + 1e-9)).astype(dtype)
    for r, max_knn in ((0.06, 50), (0.08, 12), (0.02, 30)):                          # full balls, truncated balls, mostly < 2 points
        out = pgeof.compute_features_selected(xyz, r, max_knn, ids)
        assert out.dtype == dtype and out.shape == (20000, len(ids))
        ref = cpu.compute_features_selected(xyz, r, max_knn, [int(i) for i in ids], order)
        # neighbourhoods (in the metric of the dtype) for the conditioning masks
        tree_idx, _ = cpu.radius_search(xyz.astype(np.float32), xyz.astype(np.float32), r, max_knn)
        nn, nn_ptr = radius_csr(tree_idx)
        lone = np.diff(nn_ptr.astype(np.int64)) < 2
        if dtype == np.float32:
            assert (out[lone] == 0).all()                                            # pgeof.hpp:355
        bounds = dict(max_weak=0.01, max_degenerate=1.0, max_ill=1.0) if r < 0.05 else dict(max_weak=0.01, max_degenerate=2e-3, max_ill=1e-3)
        compare_features(out, ref, row_eigvals(xyz, nn, nn_ptr), order, "selected r=%g" % r, [int(i) for i in ids], **bounds)


def test_selected_bench_jakteristics_shape_and_ties():
    """tests/bench_jakteristics.py:12-45 workload (10k float64 pts in [0,200)^3, r=5, max_knn=50, Verticality)."""
    xyz = np.random.default_rng(16).uniform(0.0, 200.0, size=(10000, 3))
    out = pgeof.compute_features_selected(xyz, 5.0, 50, [pgeof.EFeatureID.Verticality])
    ref = cpu.compute_features_selected(xyz, 5.0, 50, [12])
    # Verticality = 1 - |n_z| of unit eigenvectors in float64: every row to 1e-6, no exemptions (the eigen-gaps of 50-point balls are wide)
    assert out.shape == (10000, 1) and out.dtype == np.float64
    assert np.abs(out - ref).max() < 1e-6
    # lattice: exact-distance ties at the max_knn boundary must resolve by index like the oracle (eigenvalue columns)
    g = np.stack(np.meshgrid(*[np.arange(10)] * 3, indexing="ij"), -1).reshape(-1, 3)
    for dtype in (np.float32, np.float64):
        lat = g.astype(dtype)
        ids = [7, 8, 10, 13]
        out = pgeof.compute_features_selected(lat, 1.5, 9, ids)
        ref = cpu.compute_features_selected(lat, 1.5, 9, ids, "literal")
        np.testing.assert_allclose(out, ref, atol=2e-4, rtol=2e-3)


# --------------------------------------------------------------------------------------------
# device tensors (DLPack / torch CUDA) and the raw C ABI
# --------------------------------------------------------------------------------------------
def test_torch_cuda_tensors_match_numpy_path():
    import torch
    xyz = synth.uniform_cloud(50000, seed=17)
    t = torch.from_numpy(xyz).cuda()
    idx_np, d2_np = pgeof.knn_search(xyz, xyz, 24)
    idx_t, d2_t = pgeof.knn_search(t, t, 24)
    assert idx_t.is_cuda and idx_t.dtype == torch.uint32 and d2_t.dtype == torch.float32
    np.testing.assert_array_equal(idx_t.cpu().numpy(), idx_np)
    np.testing.assert_array_equal(d2_t.cpu().numpy(), d2_np)
    nn_ptr = (torch.arange(50001, device="cuda") * 24).to(torch.uint32)
    f_t = pgeof.compute_features(t, idx_t.view(-1), nn_ptr)                          # zero-copy CSR view
    nn, ptr = knn_csr(idx_np)
    np.testing.assert_array_equal(f_t.cpu().numpy(), pgeof.compute_features(xyz, nn, ptr))
    r_t = pgeof.radius_search(t, t[:1000], 9.0, 16)
    r_np = pgeof.radius_search(xyz, xyz[:1000], 9.0, 16)
    np.testing.assert_array_equal(r_t[0].cpu().numpy(), r_np[0])
    fused = b200.knn_features(t, 24)
    np.testing.assert_array_equal(fused.cpu().numpy(), f_t.cpu().numpy())
    with pytest.raises(TypeError):
        pgeof.knn_search(t, xyz, 4)                                                  # mixed memory spaces
    with pytest.raises(TypeError):
        pgeof.knn_search(t.double(), t.double(), 4)
    assert b200.launch_count() > 0


def test_c_abi_called_directly_through_ctypes():
    """What a non-Python FFI (cgo / JNI stub of INTEGRATION.md) would do: plain pointers and sizes."""
    lib = ctypes.CDLL(b200.LIBRARY_PATH)
    lib.pgeof_last_error.restype = ctypes.c_char_p
    xyz = synth.uniform_cloud(5000, seed=18)
    k = 16
    idx = np.empty((5000, k), np.uint32)
    d2 = np.empty((5000, k), np.float32)
    vp = ctypes.c_void_p
    rc = lib.pgeof_knn_search(xyz.ctypes.data_as(vp), ctypes.c_size_t(5000), xyz.ctypes.data_as(vp), ctypes.c_size_t(5000),
                              ctypes.c_uint32(k), idx.ctypes.data_as(vp), d2.ctypes.data_as(vp))
    assert rc == 0, lib.pgeof_last_error()
    _assert_search_equal((idx, d2), cpu.knn_search(xyz, xyz, k))
    nn, nn_ptr = knn_csr(idx)
    out = np.empty((5000, 11), np.float32)
    rc = lib.pgeof_compute_features(xyz.ctypes.data_as(vp), ctypes.c_size_t(5000), nn.ctypes.data_as(vp), ctypes.c_size_t(len(nn)),
                                    nn_ptr.ctypes.data_as(vp), ctypes.c_size_t(5000), ctypes.c_uint32(1), ctypes.c_int(0), out.ctypes.data_as(vp))
    assert rc == 0, lib.pgeof_last_error()
    compare_features(out, cpu.compute_features(xyz, nn, nn_ptr), row_eigvals_dense(xyz, idx), **UNIFORM)
    rc = lib.pgeof_knn_search(xyz.ctypes.data_as(vp), ctypes.c_size_t(5000), xyz.ctypes.data_as(vp), ctypes.c_size_t(5000),
                              ctypes.c_uint32(5001), idx.ctypes.data_as(vp), d2.ctypes.data_as(vp))
    assert rc == -1 and b"knn size" in lib.pgeof_last_error()                        # PGEOF_EINVAL


# --------------------------------------------------------------------------------------------
# paths added with the tile kernels: radius retries / adaptive grid, unaligned nn streams, slab shards
# --------------------------------------------------------------------------------------------
def test_knn_mixed_density_surfaces_clusters_bitexact():
    """A dense sheet, a sparse volume, tight clusters and a line in one mostly empty box: the uniform-density radius
    seed fails in every way here (group retries, region halving, adaptive cell edge, generic fallback)."""
    rng = np.random.default_rng(11)
    sheet = np.c_[rng.uniform(0, 60, (400000, 2)), rng.normal(0, 0.01, 400000)]
    volume = rng.uniform(0, 60, (60000, 3)) * [1, 1, 0.5]
    centers = rng.uniform(5, 55, (300, 3)) * [1, 1, 0.4]
    clusters = (centers[rng.integers(0, 300, 150000)] + rng.normal(0, 0.15, (150000, 3)))
    line = np.c_[rng.uniform(0, 60, 20000), np.full(20000, 30.0), np.full(20000, 12.0)] + rng.normal(0, 0.002, (20000, 3))
    xyz = np.concatenate([sheet, volume, clusters, line]).astype(np.float32)[rng.permutation(630000)]
    rows = rng.choice(len(xyz), 12000, replace=False)
    for k in (20, 50):
        idx, d2 = pgeof.knn_search(xyz, xyz, k)
        _assert_search_equal((idx[rows], d2[rows]), cpu.knn_search(xyz, xyz[rows], k))
        assert (np.diff(d2, axis=1) >= 0).all() and (idx[:, 0] == np.arange(len(xyz))).sum() > 0.99 * len(xyz)


def test_features_unaligned_nn_stream_on_device():
    """nn as a view that starts 4 bytes into an allocation (the 256-bit stream loads need their own alignment head)."""
    import torch
    xyz = synth.uniform_cloud(40000, seed=21)
    idx, _ = cpu.knn_search(xyz, xyz, 23)                                      # odd row length: every alignment occurs
    nn, nn_ptr = knn_csr(idx)
    ref = pgeof.compute_features(xyz, nn, nn_ptr)
    t = torch.from_numpy(xyz).cuda()
    for shift in (1, 2, 3, 5):
        big = torch.zeros(len(nn) + shift, dtype=torch.int32, device="cuda")
        big[shift:] = torch.from_numpy(nn.view(np.int32)).cuda()
        got = pgeof.compute_features(t, big[shift:].view(torch.uint32), torch.from_numpy(nn_ptr.view(np.int32)).cuda().view(torch.uint32))
        np.testing.assert_array_equal(got.cpu().numpy(), ref)
        ms = pgeof.compute_features_multiscale(t, big[shift:].view(torch.uint32), torch.from_numpy(nn_ptr.view(np.int32)).cuda().view(torch.uint32), [5, 23])
        np.testing.assert_array_equal(ms.cpu().numpy()[:, 1], ref)


def test_slab_shards_reproduce_the_single_device_rows():
    import torch
    from point_geometric_features_b200 import shard
    xyz = synth.uniform_cloud(200000, seed=31)
    t = torch.from_numpy(xyz).cuda()
    full_idx, full_d2 = pgeof.knn_search(t, t, 50)
    seen = torch.zeros(len(xyz), dtype=torch.bool, device="cuda")
    for r in range(4):
        rows, idx, d2, feats = shard.knn_features_shard(t, 50, r, 4)
        assert bool((idx.view(torch.int32) == full_idx.view(torch.int32)[rows]).all()) and bool((d2 == full_d2[rows]).all())
        seen[rows] = True
    assert bool(seen.all())


def test_local_queries_use_a_clipped_grid_and_fall_back_when_a_ball_outgrows_it():
    """Queries confined to a small box make the library index only the neighbourhood of that box (GridClip).  Queries in
    a dense part stay on the clipped grid; queries inside an empty hole need balls far larger than the halo and must
    be re-run on the full grid.  Both have to match the oracle bit for bit."""
    rng = np.random.default_rng(41)
    xyz = rng.uniform(0, 100, (300000, 3)).astype(np.float32)
    hole = np.linalg.norm(xyz - np.float32([50, 50, 50]), axis=1) < 30
    data = xyz[~hole]
    in_hole = (np.float32([50, 50, 50]) + rng.uniform(-4, 4, (700, 3))).astype(np.float32)        # nearest data ~26 away
    dense = data[(np.abs(data - np.float32([12, 12, 12])) < 6).all(1)][:3000]                     # a corner blob of data points
    for k in (8, 50):
        for q in (in_hole, dense, np.concatenate([in_hole[:50], dense[:50]])):
            _assert_search_equal(pgeof.knn_search(data, q, k), cpu.knn_search(data, q, k, brute=True))
    ridx, rd2 = pgeof.radius_search(data, dense, 1.5, 40)
    ref = cpu.radius_search(data, dense, 1.5, 40)
    np.testing.assert_array_equal(ridx, ref[0])
    np.testing.assert_array_equal(rd2, ref[1])


@pytest.mark.parametrize("order", ORDERS)
def test_fused_knn_features_is_bit_identical_to_the_two_calls(order):
    """knn_features without return_neighbors never writes the (idx, d2) rows: the tile kernel accumulates the moments
    from its staged candidates in the order compute_features walks them, so every row must match bit for bit --
    uniform and LiDAR-like data (the latter sends a third of the queries through the generic search + feature pass)."""
    import torch
    b200.set_eig_order(order)
    for xyz, k in ((synth.uniform_cloud(400000, seed=51), 50), (synth.uniform_cloud(60000, seed=52), 7),
                   (synth.lidar_like_cloud(300000, seed=2), 32)):
        t = torch.from_numpy(xyz).cuda()
        idx, _ = pgeof.knn_search(t, t, k)
        ptr = (torch.arange(len(xyz) + 1, device="cuda") * k).to(torch.uint32)
        for k_min in (1, k + 1):
            ref = pgeof.compute_features(t, idx.view(-1), ptr, k_min)
            fused = b200.knn_features(t, k, k_min)
            assert bool((fused.view(torch.int32) == ref.view(torch.int32)).all())
    b200.set_eig_order("literal")


def test_randomised_search_cases_against_brute_force():
    """tools/fuzz_search.py for a few seconds: random cloud shapes / sizes / k / radius, query == data or not, checked
    bit for bit against a brute-force evaluation of the defined metric on the device (2800 large cases ran clean in r1)."""
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_search.py"), "8"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "0 mismatches" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]


# --------------------------------------------------------------------------------------------
# BASELINE full size (10M points, k = 50): size-independent properties + sampled oracle rows
# --------------------------------------------------------------------------------------------
def test_metric_workload_full_size_properties():
    import torch
    n, k = 10_000_000, 50
    xyz = synth.uniform_cloud(n, seed=0)
    t = torch.from_numpy(xyz).cuda()
    idx, d2 = pgeof.knn_search(t, t, k)
    nn_ptr = (torch.arange(n + 1, device="cuda", dtype=torch.int64) * k).to(torch.uint32)
    feats = pgeof.compute_features(t, idx.view(-1), nn_ptr)
    torch.cuda.synchronize()
    # sortedness, self-first, index range -- on the device
    d2f = d2.view(n, k)
    assert bool((d2f[:, 1:] >= d2f[:, :-1]).all())
    idx64 = idx.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    assert bool((idx64[:, 0] == torch.arange(n, device="cuda")).all()) and bool((d2f[:, 0] == 0).all())
    assert int(idx64.max()) < n
    # checksum of recomputed distances: d2 must be exactly the defined float32 metric of the returned indices
    rows = torch.from_numpy(np.random.default_rng(3).choice(n, 200000, replace=False)).cuda()
    p = t[idx64[rows]]                                  # (m, k, 3)
    q = t[rows][:, None, :]
    diff = q - p
    re = (diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1]) + diff[..., 2] * diff[..., 2]
    assert bool((re == d2f[rows]).all())
    # sampled rows against the oracle run over the full 10M cloud
    sample = np.random.default_rng(4).choice(n, 512, replace=False)
    ref = cpu.knn_search(xyz, xyz[sample], k, brute=True)
    s = torch.from_numpy(sample).cuda()
    _assert_search_equal((idx64[s].cpu().numpy().astype(np.uint32), d2f[s].cpu().numpy()), ref)
    nn = ref[0].reshape(-1)
    fref = cpu.compute_features(xyz, nn, (np.arange(513) * k).astype(np.uint32))
    compare_features(feats[s].cpu().numpy(), fref, row_eigvals_dense(xyz, ref[0]), max_weak=0.0, max_degenerate=0.0, max_ill=0.01)
    assert bool(torch.isfinite(feats).all())


def _recomputed_d2(t, q, idx64):
    """the defined float32 metric fl(fl(dx^2 + dy^2) + dz^2) of the returned indices, on the device"""
    diff = q[:, None, :] - t[idx64]
    return (diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1]) + diff[..., 2] * diff[..., 2]


def test_c3_full_size_radius_csr_features():
    """BASELINE config 3 at its stated size: 10 M LiDAR-like points, radius_search r = 0.2, max_k = 64 -> CSR ->
    compute_features.  Size-independent properties on every row + sampled rows against the oracle over the full cloud."""
    import torch
    n, r, mk = 10_000_000, 0.2, 64
    xyz = synth.lidar_like_cloud(n, seed=0)
    t = torch.from_numpy(xyz).cuda()
    ridx, rd2 = pgeof.radius_search(t, t, r, mk)
    nn, nn_ptr = b200.radius_search_csr(t, t, r, mk)
    feats = pgeof.compute_features(t, nn, nn_ptr)
    torch.cuda.synchronize()
    valid = ridx >= 0
    cnt = valid.sum(1)
    # padding is a suffix of -1 / 0, kept entries are strictly inside the ball and sorted, the query itself comes first
    assert bool((valid[:, 1:] <= valid[:, :-1]).all()) and bool((rd2[~valid] == 0).all())
    r2 = float(np.float32(r) * np.float32(r))
    assert bool((rd2[valid] < r2).all())
    big = torch.where(valid, rd2, torch.full_like(rd2, 3e38))
    assert bool((big[:, 1:] >= big[:, :-1]).all())
    assert bool((rd2[:, 0] == 0).all()) and int(cnt.min()) >= 1
    # the CSR entry point emits exactly the compacted table
    ptr64 = nn_ptr.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    assert bool((ptr64[1:] - ptr64[:-1] == cnt).all()) and int(ptr64[-1]) == nn.shape[0]
    assert bool((nn.view(torch.int32) == ridx[valid]).all())
    # sampled rows: distances recomputed on the device, rows and features against the oracle over the full cloud
    sample = np.sort(np.random.default_rng(5).choice(n, 2000, replace=False))
    sd = torch.from_numpy(sample).cuda()
    re = _recomputed_d2(t, t[sd], ridx[sd].clamp(min=0).to(torch.int64))
    assert bool((re[valid[sd]] == rd2[sd][valid[sd]]).all())
    ref = cpu.radius_search(xyz, xyz[sample], r, mk)
    _assert_search_equal((ridx[sd].cpu().numpy(), rd2[sd].cpu().numpy()), ref)
    rnn, rptr = radius_csr(ref[0])
    fref = cpu.compute_features(xyz, rnn, rptr)
    compare_features(feats[sd].cpu().numpy(), fref, row_eigvals(xyz, rnn, rptr), "literal", "C3 features", **LIDAR)
    assert bool(torch.isfinite(feats).all())


def test_c4_full_size_knn100_multiscale():
    """BASELINE config 4 at its stated size: 10 M uniform points, one k = 100 kNN pass, k_scales = [10, 20, 50, 100]."""
    import torch
    n, k, scales = 10_000_000, 100, [10, 20, 50, 100]
    xyz = synth.uniform_cloud(n, seed=0)
    t = torch.from_numpy(xyz).cuda()
    idx, d2 = pgeof.knn_search(t, t, k)
    nn_ptr = (torch.arange(n + 1, device="cuda", dtype=torch.int64) * k).to(torch.uint32)
    ms = pgeof.compute_features_multiscale(t, idx.view(-1), nn_ptr, scales)
    torch.cuda.synchronize()
    assert tuple(ms.shape) == (n, 4, 11) and bool(torch.isfinite(ms).all())
    assert bool((d2[:, 1:] >= d2[:, :-1]).all())
    idx64 = idx.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    assert bool((idx64[:, 0] == torch.arange(n, device="cuda")).all()) and int(idx64.max()) < n
    rows = torch.from_numpy(np.random.default_rng(3).choice(n, 100000, replace=False)).cuda()
    assert bool((_recomputed_d2(t, t[rows], idx64[rows]) == d2[rows]).all())
    sample = np.random.default_rng(4).choice(n, 256, replace=False)
    ref = cpu.knn_search(xyz, xyz[sample], k, brute=True)
    sd = torch.from_numpy(sample).cuda()
    _assert_search_equal((idx64[sd].cpu().numpy().astype(np.uint32), d2[sd].cpu().numpy()), ref)
    fref = cpu.compute_features_multiscale(xyz, ref[0].reshape(-1), (np.arange(257) * k).astype(np.uint32), scales)
    got = ms[sd].cpu().numpy()
    for s_, ks in enumerate(scales):
        compare_features(got[:, s_], fref[:, s_], row_eigvals_dense(xyz, ref[0][:, :ks]), "literal", "C4 scale %d" % ks, **UNIFORM)


def test_c5_full_size_shard_knn100_optimal():
    """BASELINE config 5 at its stated size: a 50 M-point cloud, one rank's slab of the query-sharded run (1 / 25 of the
    rows, shard-local CSR offsets as pgeof.hpp:83 allows), kNN k = 100 -> compute_features_optimal(k_min_search = 10)."""
    import torch
    from point_geometric_features_b200 import shard
    n, k = 50_000_000, 100
    xyz = synth.uniform_cloud(n, seed=0)
    t = torch.from_numpy(xyz).cuda()
    rows, q = shard.slab_queries(t, 7, 25)
    m = int(q.shape[0])
    assert abs(m - n // 25) < n // 250
    idx, d2 = pgeof.knn_search(t, q, k)
    nn_ptr = (torch.arange(m + 1, device="cuda", dtype=torch.int64) * k).to(torch.uint32)
    opt = pgeof.compute_features_optimal(t, idx.view(-1), nn_ptr, 1, 1, 10)
    torch.cuda.synchronize()
    assert tuple(opt.shape) == (m, 12) and bool(torch.isfinite(opt).all())
    assert bool((opt[:, 11] >= 10).all()) and bool((opt[:, 11] <= k).all())
    assert bool((d2[:, 1:] >= d2[:, :-1]).all())
    idx64 = idx.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    assert bool((idx64[:, 0] == rows).all())
    pick = np.random.default_rng(8).choice(m, 96, replace=False)
    pd_ = torch.from_numpy(pick).cuda()
    qs = q[pd_].cpu().numpy()
    ref = cpu.knn_search(xyz, qs, k, brute=True)
    _assert_search_equal((idx64[pd_].cpu().numpy().astype(np.uint32), d2[pd_].cpu().numpy()), ref)
    fref, margin = cpu.compute_features_optimal(xyz, ref[0].reshape(-1), (np.arange(97) * k).astype(np.uint32), 1, 1, 10, return_margin=True)
    got = opt[pd_].cpu().numpy()
    sure = margin > 1e-9
    np.testing.assert_array_equal(got[sure, 11], fref[sure, 11])
    kopt = fref[:, 11].astype(int)
    ev = np.stack([np.linalg.eigvalsh(np.cov(xyz[ref[0][i, :kk]].astype(np.float64).T, bias=True)) for i, kk in enumerate(kopt)])
    compare_features(got[sure, :11], fref[sure, :11], ev[sure], "literal", "C5 optimal features", max_weak=0.05, max_degenerate=0.0, max_ill=0.02)


# --------------------------------------------------------------------------------------------
# round 2: any k, |r|, the reference's float32 optimal-k arithmetic, every PGEOF_* switch, fused path vs the oracle
# --------------------------------------------------------------------------------------------
def test_knn_and_radius_beyond_512_neighbours():
    """The reference takes any knn <= len(data) (nn_search.hpp:37,92): block-per-query kernel with keys in global scratch."""
    rng = np.random.default_rng(61)
    xyz = np.concatenate([rng.normal(0, 0.3, (4000, 3)), rng.uniform(-3, 3, (3000, 3))]).astype(np.float32)   # a dense core: balls overflow and bisect
    q = xyz[::7]
    for k in (600, 2048, len(xyz)):                                                  # knn == n: every point, sorted
        _assert_search_equal(pgeof.knn_search(xyz, q, k), cpu.knn_search(xyz, q, k, brute=True))
    for r, max_knn in ((0.6, 700), (0.2, 1500), (50.0, len(xyz))):                   # partly filled, mostly padding, everything
        _assert_search_equal(pgeof.radius_search(xyz, q, r, max_knn), cpu.radius_search(xyz, q, r, max_knn))
    nn, nn_ptr = b200.radius_search_csr(xyz, q, 0.6, 700)
    nn_ref, ptr_ref = radius_csr(cpu.radius_search(xyz, q, 0.6, 700)[0])
    np.testing.assert_array_equal(nn_ptr, ptr_ref)
    np.testing.assert_array_equal(nn, nn_ref)


def test_negative_radius_is_its_absolute_value():
    """The reference only ever uses r * r (nn_search.hpp:98, pgeof.hpp:336)."""
    xyz = np.random.default_rng(62).random((4000, 3), dtype=np.float32)
    _assert_search_equal(pgeof.radius_search(xyz, xyz, -0.1, 12), pgeof.radius_search(xyz, xyz, 0.1, 12))
    ids = [pgeof.EFeatureID.Length, pgeof.EFeatureID.Surface, pgeof.EFeatureID.Verticality]
    a = pgeof.compute_features_selected(xyz.astype(np.float64), -0.1, 30, ids)
    b = pgeof.compute_features_selected(xyz.astype(np.float64), 0.1, 30, ids)
    np.testing.assert_allclose(a, b, atol=1e-9)          # (the moments are summed in grid order: equal up to float64 round-off)
    assert np.abs(a).max() > 0
    with pytest.raises(ValueError):
        pgeof.radius_search(xyz, xyz, float("nan"), 4)


def test_optimal_k_against_the_float32_arithmetic_of_the_reference():
    """pgeof.hpp:286-289 compares FLOAT32 eigenentropies of float32 two-pass PCAs; the kernel decides in float64.  The two may
    only differ where the float32 scan itself cannot tell the two neighbourhood sizes apart: report the agreement rate and
    require every disagreement to sit on a float32 entropy margin below 1e-5."""
    for xyz, k, what in ((synth.uniform_cloud(30000, seed=14), 100, "uniform"), (dense_lidar(30000, seed=5), 60, "lidar")):
        idx, _ = pgeof.knn_search(xyz, xyz, k)
        nn, nn_ptr = knn_csr(idx)
        opt = pgeof.compute_features_optimal(xyz, nn, nn_ptr, 1, 1, 10)
        ref32, margin32 = cpu.compute_features_optimal(xyz, nn, nn_ptr, 1, 1, 10, f64=False, return_margin=True)
        differ = opt[:, 11] != ref32[:, 11]
        rate = 1.0 - differ.mean()
        print("optimal-k agreement with the float32 oracle (%s): %.5f, largest margin among disagreements %.3g"
              % (what, rate, margin32[differ].max() if differ.any() else 0.0))
        assert rate > 0.99
        assert (margin32[differ] < 1e-5).all()


SWITCHES = [("PGEOF_KNN_TILE", "0"), ("PGEOF_RADIUS_TILE", "0"), ("PGEOF_GRID_CLIP", "0"), ("PGEOF_KNN_WARPS", "2"), ("PGEOF_KNN_FUSED", "0"), ("PGEOF_KNN_LOCK", "0"), ("PGEOF_KNN_PAIR", "1"), ("PGEOF_KNN_ROLLED", "1"), ("PGEOF_KNN_TILE128", "1"), ("PGEOF_KNN_TWO_LEVEL", "0"),
            ("PGEOF_FEATURES_SORT", "0"), ("PGEOF_FEATURES_CTA", "128"), ("PGEOF_FEATURES_CTA", "256"), ("PGEOF_FEATURES_SORT_MIN_ROWS", "0"),
            ("PGEOF_GRID_SCATTER", "1"), ("PGEOF_GRID_SCATTER", "2"), ("PGEOF_RADIUS_CSR_TWO_PASS", "1"), ("PGEOF_MULTISCALE_SPLIT", "0"), ("PGEOF_MULTISCALE_CHUNK_ROWS", "7001"), ("PGEOF_FEATURES_TEX", "0"), ("PGEOF_FEATURES_CARVEOUT", "-1"), ("PGEOF_FEATURES_PHASES", "1"), ("PGEOF_OPTIMAL_SCAN", "0"), ("PGEOF_OPTIMAL_SCAN", "1"), ("PGEOF_OPTIMAL_CTAS", "6")]


@pytest.mark.parametrize("name,value", SWITCHES)
def test_every_switch_leaves_the_results_unchanged(name, value, monkeypatch):
    """the PGEOF_* switches select code paths, never results."""
    import torch
    xyz = synth.uniform_cloud(120000, seed=71)
    t = torch.from_numpy(xyz).cuda()
    q = t[(t[:, 2] < 60)]                                                            # local queries: the clipped-grid path
    base = {}
    for phase in ("default", "switched"):
        if phase == "switched":
            monkeypatch.setenv(name, value)
        idx, d2 = pgeof.knn_search(t, t, 50)
        qi, qd = pgeof.knn_search(t, q, 20)
        wi, wd = pgeof.knn_search(t, t[:30000], 100)                                 # 64 < k <= 128 (PGEOF_KNN_TILE128)
        ri, rd = pgeof.radius_search(t, t, 6.0, 40)
        ptr = (torch.arange(len(xyz) + 1, device="cuda") * 50).to(torch.uint32)
        f = pgeof.compute_features(t, idx.view(-1), ptr)
        ms = pgeof.compute_features_multiscale(t, idx.view(-1), ptr, [10, 13, 20, 50])
        op = pgeof.compute_features_optimal(t, idx.view(-1), ptr, 1, 1, 10)
        fu = b200.knn_features(t, 50)
        cn, cp = b200.radius_search_csr(t, t, 6.0, 40)                               # single-search pair (PGEOF_RADIUS_CSR_TWO_PASS, PGEOF_RADIUS_TILE)
        got = [x.cpu().numpy().view(np.uint32) for x in (idx, d2, qi, qd, wi, wd, ri, rd, f, ms, op, fu, cn, cp)]
        if phase == "default":
            base = got
            ref = cpu.knn_search(xyz, xyz[:3000], 50)
            _assert_search_equal((idx[:3000].cpu().numpy(), d2[:3000].cpu().numpy()), ref)
        else:
            for a, b in zip(base, got):
                np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("order", ORDERS)
def test_fused_knn_features_against_the_oracle(order):
    """knn_features without the neighbour lists, checked against the CPU oracle itself (not against the two-call CUDA path)."""
    b200.set_eig_order(order)
    for xyz, k, bounds in ((synth.uniform_cloud(150000, seed=81), 50, UNIFORM), (dense_lidar(100000, seed=6), 32, LIDAR)):
        fused = b200.knn_features(xyz, k)
        idx, _ = cpu.knn_search(xyz, xyz, k)
        nn, nn_ptr = knn_csr(idx)
        compare_features(fused, cpu.compute_features(xyz, nn, nn_ptr, 1, order), row_eigvals_dense(xyz, idx), order, "fused k=%d" % k, **bounds)


def test_slab_shards_from_the_library_match_the_host_restatement():
    """pgeof_slab_plan_dev / _fill_dev against shard.slab_queries on the host tensor: same rows, same order, same coordinates."""
    import torch
    from point_geometric_features_b200 import shard
    for xyz in (synth.uniform_cloud(300007, seed=91), dense_lidar(200000, seed=7),
                np.c_[np.random.default_rng(92).uniform(0, 9, (5000, 2)), np.full(5000, 3.0)].astype(np.float32)):   # zero extent along z
        t = torch.from_numpy(xyz)
        for world in (1, 2, 3, 8):
            total = 0
            for r in range(world):
                rows_h, q_h = shard.slab_queries(t, r, world)
                rows_d, q_d = shard.slab_queries(t.cuda(), r, world)
                assert rows_d.dtype == torch.int64 and torch.equal(rows_d.cpu(), rows_h) and torch.equal(q_d.cpu(), q_h)
                total += rows_h.shape[0]
            assert total == len(xyz)


def test_csr_native_outputs_and_wide_offsets(monkeypatch):
    """knn_search_csr / radius_search_csr emit what the README glue builds; nn_ptr may be 64-bit."""
    import torch
    xyz = dense_lidar(150000, seed=8)
    idx, _ = cpu.knn_search(xyz, xyz, 17)
    nn_ref, ptr_ref = knn_csr(idx)
    for wide in (False, True):
        nn, ptr = b200.knn_search_csr(xyz, xyz, 17, wide)
        assert nn.dtype == np.uint32 and ptr.dtype == (np.uint64 if wide else np.uint32)
        np.testing.assert_array_equal(nn, nn_ref)
        np.testing.assert_array_equal(ptr.astype(np.uint64), ptr_ref.astype(np.uint64))
        t = torch.from_numpy(xyz).cuda()
        nn_t, ptr_t = b200.knn_search_csr(t, t, 17, wide)
        assert ptr_t.dtype == (torch.int64 if wide else torch.uint32)
        np.testing.assert_array_equal(nn_t.cpu().numpy(), nn_ref)
        # the feature functions take either width and give bit-identical rows
        f32 = pgeof.compute_features(xyz, nn_ref, ptr_ref)
        np.testing.assert_array_equal(pgeof.compute_features(xyz, nn, ptr), f32)
        np.testing.assert_array_equal(pgeof.compute_features(t, nn_t, ptr_t).cpu().numpy(), f32)
        np.testing.assert_array_equal(pgeof.compute_features_multiscale(xyz, nn, ptr, [5, 17]), pgeof.compute_features_multiscale(xyz, nn_ref, ptr_ref, [5, 17]))
        np.testing.assert_array_equal(pgeof.compute_features_optimal(xyz, nn, ptr, 1, 1, 4), pgeof.compute_features_optimal(xyz, nn_ref, ptr_ref, 1, 1, 4))
    # radius CSR: one search for the call pair (parked padded table) == the two-pass count + fill == README glue on the padded result
    ridx, _ = cpu.radius_search(xyz, xyz[:40000], 0.25, 48)
    nn_ref, ptr_ref = radius_csr(ridx)
    for two_pass in (False, True):
        if two_pass:
            monkeypatch.setenv("PGEOF_RADIUS_CSR_TWO_PASS", "1")
        for a, q in ((xyz, xyz[:40000]), (torch.from_numpy(xyz).cuda(), torch.from_numpy(xyz[:40000]).cuda())):
            nn, ptr = b200.radius_search_csr(a, q, 0.25, 48)
            nn, ptr = (nn.cpu().numpy(), ptr.cpu().numpy()) if hasattr(nn, "cpu") else (nn, ptr)
            np.testing.assert_array_equal(ptr, ptr_ref)
            np.testing.assert_array_equal(nn, nn_ref)
    with pytest.raises(IndexError):
        pgeof.compute_features(xyz, nn_ref, np.array([0, 2 ** 33], np.uint64))


def test_torch_adaptor_frnn_and_spt_helpers_against_brute_force():
    """FRNN-shaped batched fixed-radius search and SuperPoint-Transformer style knn_1 / knn_2 (nn_search.hpp:72)."""
    import torch
    from point_geometric_features_b200 import torch_adaptor as ta
    g = torch.Generator().manual_seed(5)
    N, P1, P2, K, r = 3, 700, 900, 12, 0.17
    p1, p2 = torch.rand(N, P1, 3, generator=g).cuda(), torch.rand(N, P2, 3, generator=g).cuda()
    l1, l2 = torch.tensor([700, 433, 1]).cuda(), torch.tensor([900, 10, 512]).cuda()
    dists, idxs, nn, grid = ta.frnn_grid_points(p1, p2, l1, l2, K=K, r=r, return_nn=True)
    assert dists.shape == (N, P1, K) and idxs.dtype == torch.int64 and nn.shape == (N, P1, K, 3) and grid is None
    for n in range(N):
        a, b = p1[n, :l1[n]], p2[n, :l2[n]]
        diff = a[:, None, :] - b[None, :, :]
        d2 = (diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1]) + diff[..., 2] * diff[..., 2]      # the defined float32 metric
        d2 = torch.where(d2 < torch.tensor(r, dtype=torch.float32).cuda() ** 2, d2, torch.full_like(d2, float("inf")))
        order = torch.argsort(d2, dim=1, stable=True)[:, :K]                                                # ties by index
        ref_d = torch.gather(d2, 1, order)
        kk = order.shape[1]
        ref_i = torch.where(torch.isinf(ref_d), torch.full_like(order, -1), order)
        ref_d = torch.where(torch.isinf(ref_d), torch.full_like(ref_d, -1.0), ref_d)
        assert torch.equal(idxs[n, :l1[n], :kk], ref_i) and torch.equal(dists[n, :l1[n], :kk], ref_d)
        assert bool((idxs[n, :l1[n], kk:] == -1).all()) and bool((idxs[n, l1[n]:] == -1).all()) and bool((dists[n, l1[n]:] == -1).all())
        hit = idxs[n, :l1[n]] >= 0
        assert torch.equal(nn[n, :l1[n]][hit], b[idxs[n, :l1[n]][hit]])
    # SPT helpers on a flat cloud with a sorted batch vector
    xyz = torch.cat([p2[0], p2[2, :512] + 5.0])
    batch = torch.cat([torch.zeros(900, dtype=torch.int64), torch.ones(512, dtype=torch.int64)]).cuda()
    nb, ds = ta.knn_1(xyz, 8, r_max=0.2, batch=batch)
    assert nb.shape == (1412, 8) and ds.shape == (1412, 8)
    assert bool((nb[:900][nb[:900] >= 0] < 900).all()) and bool((nb[900:][nb[900:] >= 0] >= 900).all())    # no neighbour across clouds
    assert bool((nb != torch.arange(1412, device="cuda")[:, None]).all())                                   # self dropped
    full = torch.cdist(xyz[:900].double(), xyz[:900].double())
    full.fill_diagonal_(float("inf"))
    ref = torch.sort(full, dim=1).values[:, :8]
    got = ds[:900].double()
    ok = got >= 0
    assert torch.allclose(got[ok], ref[ok], atol=1e-6) and bool((ref[~ok] >= 0.2 - 1e-6).all())
    nb2, ds2 = ta.knn_2(xyz, xyz[:50], 5, r_max=10.0)
    assert bool((nb2[:, 0] == torch.arange(50, device="cuda")).all()) and bool((ds2[:, 0] == 0).all())


def test_chunked_host_pipeline_matches_the_serial_flavour(monkeypatch):
    """Host-buffer feature calls upload nn in chunks beside compute + download (capi.cu csr_pipeline): same bits as the serial
    flavour, ragged rows and uint64 offsets included; a corrupt nn_ptr is refused either way."""
    xyz = dense_lidar(150000, seed=31)
    ridx, _ = pgeof.radius_search(xyz, xyz, 0.25, 48)
    nn, nn_ptr = radius_csr(ridx)
    kidx, _ = pgeof.knn_search(xyz, xyz, 40)
    knn, kptr = knn_csr(kidx)
    results = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("PGEOF_HOST_CHUNK_MB", mode)            # 0: serial, 1: 1 MB slices -> 16 chunks
        results[mode] = [pgeof.compute_features(xyz, nn, nn_ptr, 3),
                         pgeof.compute_features(xyz, knn, kptr),
                         pgeof.compute_features_multiscale(xyz, knn, kptr, [5, 20, 40]),
                         pgeof.compute_features_optimal(xyz, knn, kptr, 1, 1, 10),
                         pgeof.compute_features(xyz, nn, nn_ptr.astype(np.uint64), 3)]                 # 64-bit offsets (extension)
        bad = nn_ptr.copy()
        bad[len(bad) // 2] = bad[len(bad) // 2 - 1] - 1 if bad[len(bad) // 2 - 1] > 0 else bad[-1] + 5     # non-monotonic offsets
        with pytest.raises((IndexError, ValueError)):
            pgeof.compute_features(xyz, nn, bad)
    for a, b in zip(results["0"], results["1"]):
        np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32))
    assert np.abs(results["1"][0]).max() > 0


def test_knn_sparse_queued_queries_on_the_coarser_grid(monkeypatch):
    """Non-uniform cloud: the queued queries of the sparse parts (scatter blobs, poles) are deferred to a grid with 4x the cell
    edge (search.cu: two-level handling).  Bit-exact with and without the second level, for both tile kernels."""
    monkeypatch.setenv("PGEOF_KNN_COARSE_MIN", "1")              # build the second grid however few queries were deferred
    xyz = dense_lidar(300000, seed=41)
    rows = np.random.default_rng(42).choice(len(xyz), 4000, replace=False)
    for k in (20, 50):
        idx, d2 = pgeof.knn_search(xyz, xyz, k)
        _assert_search_equal((idx[rows], d2[rows]), cpu.knn_search(xyz, xyz[rows], k))
        monkeypatch.setenv("PGEOF_KNN_TWO_LEVEL", "0")
        idx0, d20 = pgeof.knn_search(xyz, xyz, k)
        monkeypatch.delenv("PGEOF_KNN_TWO_LEVEL")
        np.testing.assert_array_equal(idx, idx0)
        np.testing.assert_array_equal(d2.view(np.uint32), d20.view(np.uint32))
