"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/pgeof_b200.h declares, and the host-side logic of the binding (argument names, defaults,
strict dtypes, error types of src/pgeof_ext.cpp + the three std::invalid_argument sites) behaves
like the reference.  No compute call is made: there is no GPU here."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest

import point_geometric_features_b200 as b200

HEADER = b200.HEADER_PATH


def _declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"PGEOF_API\s+[\w\s\*]+?\b(pgeof_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(b200.LIBRARY_PATH)
    names = _declared_symbols()
    assert len(names) >= 25, names
    for name in names:
        assert hasattr(lib, name), "libpgeof_b200.so does not export %s" % name
    lib.pgeof_abi_version.restype = ctypes.c_int
    assert lib.pgeof_abi_version() == b200.pgeof_ext.abi_version == 1
    lib.pgeof_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.pgeof_last_error(), bytes)


def test_library_is_sm100a_only_and_uses_bulk_copies():
    """The product is sm_100a code: one cubin arch, no PTX for other targets."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", b200.LIBRARY_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_reference_surface_names_defaults_and_enum():
    import pgeof
    for name in ("EFeatureID", "compute_features", "compute_features_multiscale", "compute_features_optimal",
                 "knn_search", "radius_search", "compute_features_selected"):      # src/pgeof/__init__.py:1-9
        assert hasattr(pgeof, name)
    names = ["Linearity", "Planarity", "Scattering", "VerticalityPGEOF", "Normal_x", "Normal_y", "Normal_z", "Length",
             "Surface", "Volume", "Curvature", "K_optimal", "Verticality", "Eigentropy"]     # pca.hpp:47-63
    for value, name in enumerate(names):
        assert int(getattr(pgeof.EFeatureID, name)) == value
        assert getattr(pgeof.pgeof_ext, name) == getattr(pgeof.EFeatureID, name)              # export_values()
    doc = pgeof.compute_features.__doc__
    assert "xyz" in doc and "nn_ptr" in doc and "k_min: typing.SupportsInt = 1" in doc.replace("int = 1", "typing.SupportsInt = 1") or "k_min" in doc
    assert "k_scales" in pgeof.compute_features_multiscale.__doc__                            # pgeof_ext.cpp:61
    assert "k_min_search" in pgeof.compute_features_optimal.__doc__ and "k_step" in pgeof.compute_features_optimal.__doc__
    assert "search_radius" in pgeof.radius_search.__doc__ and "max_knn" in pgeof.radius_search.__doc__
    assert "selected_features" in pgeof.compute_features_selected.__doc__


def test_strict_dtypes_raise_typeerror():
    import pgeof
    xyz = np.zeros((8, 3), np.float32)
    nn, ptr = np.zeros(8, np.uint32), np.array([0, 8], np.uint32)
    with pytest.raises(TypeError):
        pgeof.knn_search(xyz.astype(np.float64), xyz, 2)                 # "data"_a.noconvert()
    with pytest.raises(TypeError):
        pgeof.knn_search(xyz, xyz.astype(np.float64), 2)
    with pytest.raises(TypeError):
        pgeof.radius_search(xyz, np.zeros((8, 2), np.float32), 1.0, 2)   # (n, 3) only
    with pytest.raises(TypeError):
        pgeof.compute_features(xyz, nn.astype(np.int64), ptr)            # uint32 only (pgeof.hpp:78-79)
    with pytest.raises(TypeError):
        pgeof.compute_features(xyz, nn, ptr.astype(np.int32))
    with pytest.raises(TypeError):
        pgeof.compute_features(xyz.astype(np.float64), nn, ptr)          # float-only binding (pgeof_ext.cpp:35)
    with pytest.raises(TypeError):
        pgeof.compute_features(np.asfortranarray(xyz), nn, ptr)
    with pytest.raises(TypeError):
        pgeof.compute_features_selected(xyz.astype(np.int32), 1.0, 4, [pgeof.EFeatureID.Verticality])
    with pytest.raises(TypeError):
        pgeof.knn_search([[0.0, 0.0, 0.0]], xyz, 1)                      # not an array


def test_invalid_argument_sites_raise_valueerror():
    import pgeof
    xyz = np.zeros((8, 3), np.float32)
    nn, ptr = np.zeros(8, np.uint32), np.array([0, 8], np.uint32)
    with pytest.raises(ValueError, match="knn size"):
        pgeof.knn_search(xyz, xyz, 9)                                    # nn_search.hpp:37
    with pytest.raises(ValueError, match="max knn size"):
        pgeof.radius_search(xyz, xyz, 1.0, 9)                            # nn_search.hpp:92-95
    with pytest.raises(ValueError, match="k_min"):
        pgeof.compute_features(xyz, nn, ptr, 0)                          # pgeof.hpp:81
    with pytest.raises(ValueError, match="k_scales"):
        pgeof.compute_features_multiscale(xyz, nn, ptr, [50, 20])        # pgeof.hpp:165 (test commented out upstream)
    with pytest.raises(ValueError, match="k_scales"):
        pgeof.compute_features_multiscale(xyz, nn, ptr, np.array([0, 20]))
    with pytest.raises(ValueError, match="k_min"):
        pgeof.compute_features_optimal(xyz, nn, ptr, 0, 1, 0)            # pgeof.hpp:250 ('&&' kept)
    with pytest.raises(ValueError, match="k_step"):
        pgeof.compute_features_optimal(xyz, nn, ptr, 1, 0, 1)            # reference: modulo by zero


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point fails loudly (never computes on the CPU)."""
    import pgeof
    if b200.device_count() > 0:
        pytest.skip("a GPU is present")
    xyz = np.random.default_rng(0).random((16, 3)).astype(np.float32)
    nn, ptr = np.arange(16, dtype=np.uint32), np.array([0, 8, 16], np.uint32)
    calls = [
        lambda: pgeof.knn_search(xyz, xyz, 4),
        lambda: pgeof.radius_search(xyz, xyz, 0.5, 4),
        lambda: pgeof.compute_features(xyz, nn, ptr),
        lambda: pgeof.compute_features_multiscale(xyz, nn, ptr, [2, 4]),
        lambda: pgeof.compute_features_optimal(xyz, nn, ptr),
        lambda: pgeof.compute_features_selected(xyz, 0.5, 4, [pgeof.EFeatureID.Verticality]),
        lambda: pgeof.compute_features_selected(xyz.astype(np.float64), 0.5, 4, [pgeof.EFeatureID.Verticality]),
    ]
    for call in calls:
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call()
    lib = ctypes.CDLL(b200.LIBRARY_PATH)
    lib.pgeof_last_error.restype = ctypes.c_char_p
    rc = lib.pgeof_knn_search(xyz.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(16), xyz.ctypes.data_as(ctypes.c_void_p),
                              ctypes.c_size_t(16), ctypes.c_uint32(4), None, None)
    assert rc == -2 and b"no CPU fallback" in lib.pgeof_last_error()     # PGEOF_ECUDA


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing in the package may reference it."""
    pkg = os.path.dirname(b200.__file__)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(root, f), errors="ignore").read()
                assert "oracle" not in text.lower().replace("oracle takes", ""), os.path.join(root, f)
    assert "oracle" not in inspect.getsource(b200)


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU port on a bounded sample): exactly one JSON line on stdout with the contract's keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-sample", "20000"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "config", "cpu_baseline", "e2e"):
        assert key in d
    assert d["impl"] == "reference" and d["unit"] == "points/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0


@pytest.mark.parametrize("config", ["C3", "C4", "C5"])
def test_bench_reference_arm_runs_every_config(config):
    """The other BASELINE.json configurations behind --config: same one-line contract, bounded CPU sample."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--config", config, "--steps", "1", "--warmup", "0",
                        "--cpu-sample", "5000"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    d = json.loads(p.stdout.strip())
    assert d["impl"] == "reference" and d["value"] > 0 and d["config"]["sample_points"] == 5000


def test_reference_arm_maps_no_product_library():
    """The CPU arm must not load the CUDA library or the binding: run it under an import hook that fails on them."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, runpy\n"
            "sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0', '--cpu-sample', '3000']\n"
            "runpy.run_path(%r, run_name='__main__')\n"
            "maps = open('/proc/self/maps').read()\n"
            "assert 'libpgeof_b200' not in maps and 'pgeof_ext' not in maps, 'product library mapped by the reference arm'\n"
            "assert 'point_geometric_features_b200' not in sys.modules\n") % os.path.join(root, "bench.py")
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]


def test_type_stub_and_marker_match_the_module():
    """Packaging parity (reference CMakeLists.txt:34-41 ships pgeof_ext.pyi + py.typed): the hand-written stub names exactly
    what the compiled module exports."""
    import ast
    import point_geometric_features_b200.pgeof_ext as ext
    pkg = os.path.dirname(ext.__file__)
    root = os.path.dirname(pkg)
    assert os.path.exists(os.path.join(pkg, "py.typed")) and os.path.exists(os.path.join(root, "pgeof", "py.typed"))
    tree = ast.parse(open(os.path.join(pkg, "pgeof_ext.pyi")).read())
    names = {n.name for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef))} | {n.target.id for n in tree.body if isinstance(n, ast.AnnAssign)}
    assert names == {n for n in dir(ext) if not n.startswith("_")}
    enum_cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "EFeatureID"][0]
    members = {n.targets[0].id: n.value.value for n in enum_cls.body if isinstance(n, ast.Assign)}
    assert members == {k: int(v) for k, v in ext.EFeatureID.__members__.items()}
