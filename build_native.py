"""In-tree build of the CUDA library and the Python binding (no JIT cache, no pip).

    python build_native.py [--force] [--verbose]

* ``libpgeof_b200.so``  -- csrc/*.cu compiled by nvcc for sm_100a only
  (``-gencode arch=compute_100a,code=sm_100a -lineinfo``), exporting the C ABI of
  ``include/pgeof_b200.h``.
* ``pgeof_ext.*.so``    -- binding/pgeof_ext.cpp (pybind11; nanobind is not in this image,
  SURVEY.md F2) linked against the library with ``$ORIGIN`` rpath.

Both land inside the package directory so that they travel with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
HERE = os.path.join(ROOT, "point_geometric_features_b200")
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libpgeof_b200.so")
EXT = os.path.join(HERE, "pgeof_ext" + sysconfig.get_config_var("EXT_SUFFIX"))

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
CXX = os.environ.get("CXX") or shutil.which("g++") or "g++"

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC,-fvisibility=hidden,-O3",
    "-Xptxas", "-v" if os.environ.get("PGEOF_PTXAS_V") else "-O3",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode != 0 or (verbose and p.stdout.strip()):
        print(p.stdout, flush=True)
    if p.returncode != 0:
        raise RuntimeError("build step failed: " + " ".join(cmd))
    return p.stdout


def build_library(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "pgeof_b200.h"))
    cus = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    jobs, objs = [], []
    for f in cus:
        src = os.path.join(CSRC, f)
        obj = os.path.join(OBJ, f[:-3] + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + headers):
            jobs.append([NVCC] + NVCC_FLAGS + ["-c", src, "-o", obj])
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(lambda c: _run(c, verbose), jobs))
    if force or jobs or _newer(LIB, objs):
        _run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-cudart", "static"], verbose)
    return LIB


def build_binding(force=False, verbose=False):
    import pybind11

    src = os.path.join(HERE, "binding", "pgeof_ext.cpp")
    if not (force or _newer(EXT, [src, LIB, os.path.join(ROOT, "include", "pgeof_b200.h")])):
        return EXT
    cmd = [CXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden",
           "-I" + pybind11.get_include(), "-I" + sysconfig.get_paths()["include"], "-I" + os.path.join(ROOT, "include"),
           src, "-o", EXT, "-L" + HERE, "-lpgeof_b200", "-Wl,-rpath,$ORIGIN"]
    _run(cmd, verbose)
    return EXT


def build(force=False, verbose=False):
    return build_library(force, verbose), build_binding(force, verbose)


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv)
    print("built", LIB, "and", EXT)
