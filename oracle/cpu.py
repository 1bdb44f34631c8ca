"""ctypes front-end of ``oracle/cpu_ref.cpp`` -- TEST INFRASTRUCTURE ONLY.

Same import rule as ``ref_numpy.py``: tests, ``smoke()`` and ``bench.py``'s CPU
legs only.  ``_f32`` entry points run the reference's float arithmetic (the timed
CPU baseline, kind "port"); ``_f64`` ones evaluate the same algorithm in double.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libpgeof_oracle.so")
_lib = None

_c = ctypes
_pf, _pd = _c.POINTER(_c.c_float), _c.POINTER(_c.c_double)
_pu, _pi = _c.POINTER(_c.c_uint32), _c.POINTER(_c.c_int32)


def build(force=False):
    src = os.path.join(_HERE, "cpu_ref.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B" if force else "-s", "_build/libpgeof_oracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.oracle_hardware_threads.restype = _c.c_int
    return _lib


def hardware_threads():
    return int(lib().oracle_hardware_threads())


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _p(a, t):
    return a.ctypes.data_as(t)


def knn_search(data, query, k, nthreads=0, brute=False):
    data, query = _f32(data), _f32(query)
    if k > data.shape[0]:
        raise ValueError("knn size is greater than the data point cloud size")
    nq = query.shape[0]
    idx = np.empty((nq, k), np.uint32)
    d2 = np.empty((nq, k), np.float32)
    fn = lib().oracle_knn_brute_f32 if brute else lib().oracle_knn_f32
    fn(_p(data, _pf), _c.c_size_t(data.shape[0]), _p(query, _pf), _c.c_size_t(nq), _c.c_uint32(k),
       _p(idx, _pu), _p(d2, _pf), _c.c_int(nthreads))
    return idx, d2


def radius_search(data, query, r, max_knn, nthreads=0):
    data, query = _f32(data), _f32(query)
    if max_knn > data.shape[0]:
        raise ValueError("max knn size is greater than the data point cloud size")
    nq = query.shape[0]
    idx = np.empty((nq, max_knn), np.int32)
    d2 = np.empty((nq, max_knn), np.float32)
    lib().oracle_radius_f32(_p(data, _pf), _c.c_size_t(data.shape[0]), _p(query, _pf), _c.c_size_t(nq),
                            _c.c_float(r), _c.c_uint32(max_knn), _p(idx, _pi), _p(d2, _pf), _c.c_int(nthreads))
    return idx, d2


def _order(eig_order):
    return {"literal": 0, "documented": 1}[eig_order]


def compute_features(xyz, nn, nn_ptr, k_min=1, eig_order="literal", f64=True, nthreads=0):
    if k_min < 1:
        raise ValueError("k_min should be > 1")
    xyz, nn, nn_ptr = _f32(xyz), _u32(nn), _u32(nn_ptr)
    n = len(nn_ptr) - 1
    out = np.empty((n, 11), np.float64 if f64 else np.float32)
    fn = lib().oracle_features_f64 if f64 else lib().oracle_features_f32
    fn(_p(xyz, _pf), _p(nn, _pu), _p(nn_ptr, _pu), _c.c_size_t(n), _c.c_uint32(k_min), _c.c_int(_order(eig_order)),
       _p(out, _pd if f64 else _pf), _c.c_int(nthreads))
    return out


def compute_features_multiscale(xyz, nn, nn_ptr, k_scales, eig_order="literal", f64=True, nthreads=0):
    scales = [int(s) for s in k_scales]
    prev = 1
    for s in scales:
        if s < prev:
            raise ValueError("k_scales should be > 1 and sorted in ascending order")
        prev = s
    xyz, nn, nn_ptr = _f32(xyz), _u32(nn), _u32(nn_ptr)
    sc = np.asarray(scales, np.uint32)
    n = len(nn_ptr) - 1
    out = np.empty((n, len(scales), 11), np.float64 if f64 else np.float32)
    fn = lib().oracle_multiscale_f64 if f64 else lib().oracle_multiscale_f32
    fn(_p(xyz, _pf), _p(nn, _pu), _p(nn_ptr, _pu), _c.c_size_t(n), _p(sc, _pu), _c.c_size_t(len(scales)),
       _c.c_int(_order(eig_order)), _p(out, _pd if f64 else _pf), _c.c_int(nthreads))
    return out


def compute_features_optimal(xyz, nn, nn_ptr, k_min=1, k_step=1, k_min_search=1, eig_order="literal", f64=True,
                             nthreads=0, return_margin=False):
    if k_min < 1 and k_min_search < 1:
        raise ValueError("k_min and k_min_search should be > 1")
    if k_step < 1:
        raise ValueError("k_step should be >= 1")
    xyz, nn, nn_ptr = _f32(xyz), _u32(nn), _u32(nn_ptr)
    n = len(nn_ptr) - 1
    dt = np.float64 if f64 else np.float32
    out = np.empty((n, 12), dt)
    margin = np.empty(n, dt)
    fn = lib().oracle_optimal_f64 if f64 else lib().oracle_optimal_f32
    pt = _pd if f64 else _pf
    fn(_p(xyz, _pf), _p(nn, _pu), _p(nn_ptr, _pu), _c.c_size_t(n), _c.c_uint32(k_min), _c.c_uint32(k_step),
       _c.c_uint32(k_min_search), _c.c_int(_order(eig_order)), _p(out, pt), _p(margin, pt), _c.c_int(nthreads))
    return (out, margin) if return_margin else out


def compute_features_selected(xyz, r, max_knn, ids, eig_order="literal", nthreads=0):
    xyz = np.ascontiguousarray(xyz)
    ids = np.asarray([int(i) for i in ids], np.int32)
    n = xyz.shape[0]
    if xyz.dtype == np.float32:
        out = np.empty((n, len(ids)), np.float32)
        lib().oracle_selected_f32(_p(xyz, _pf), _c.c_size_t(n), _c.c_float(r), _c.c_uint32(max_knn), _p(ids, _pi),
                                  _c.c_size_t(len(ids)), _c.c_int(_order(eig_order)), _p(out, _pf), _c.c_int(nthreads))
    else:
        xyz = np.ascontiguousarray(xyz, np.float64)
        out = np.empty((n, len(ids)), np.float64)
        lib().oracle_selected_f64(_p(xyz, _pd), _c.c_size_t(n), _c.c_double(r), _c.c_uint32(max_knn), _p(ids, _pi),
                                  _c.c_size_t(len(ids)), _c.c_int(_order(eig_order)), _p(out, _pd), _c.c_int(nthreads))
    return out


def eig3(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.float32:
        w, v = np.empty(3, np.float32), np.empty((3, 3), np.float32)
        lib().oracle_eig3_f32(_p(a, _pf), _p(w, _pf), _p(v, _pf))
    else:
        a = np.ascontiguousarray(a, np.float64)
        w, v = np.empty(3, np.float64), np.empty((3, 3), np.float64)
        lib().oracle_eig3_f64(_p(a, _pd), _p(w, _pd), _p(v, _pd))
    return w, v
