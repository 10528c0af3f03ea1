"""ctypes wrapper of oracle/cpu_ref.c — TEST INFRASTRUCTURE / TIMED CPU BASELINE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libcpu_ref.so")


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def _load():
    if not os.path.exists(_LIB):
        build()
    lib = C.CDLL(_LIB)
    p = C.c_void_p
    lib.cpu_ref_hashagg_sum_count.restype = C.c_int64
    lib.cpu_ref_hashagg_sum_count.argtypes = [p, p, p, p, C.c_int64, C.c_int, p, p, p, p, p, C.c_int64]
    lib.cpu_ref_filter_project.restype = C.c_int64
    lib.cpu_ref_filter_project.argtypes = [p, p, p, p, C.c_int64, C.c_int64, C.c_int, p, p, p]
    lib.cpu_ref_q1_filter_agg.restype = C.c_int64
    lib.cpu_ref_q1_filter_agg.argtypes = [p, p, p, p, C.c_int64, C.c_int64, C.c_int64, C.c_int, p, p, p, C.c_int64]
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _bits(valid):
    """bool array -> Arrow validity bitmap (or None when all valid)"""
    if valid is None or bool(np.all(valid)):
        return None
    return np.packbits(np.asarray(valid, bool), bitorder="little")


def hashagg_sum_count(k, v, k_valid=None, v_valid=None, nthreads=1, max_groups=None):
    """SUM(v), COUNT(v) GROUP BY k -> dict(k, k_valid, sum, sum_valid, count) (dense, insertion order)"""
    k = np.ascontiguousarray(k, np.int64); v = np.ascontiguousarray(v, np.int64)
    n = len(k)
    cap = int(max_groups if max_groups is not None else n)
    kb, vb = _bits(k_valid), _bits(v_valid)
    ok = np.empty(cap, np.int64); okv = np.empty(cap, np.uint8); os_ = np.empty(cap, np.int64); osv = np.empty(cap, np.uint8); oc = np.empty(cap, np.int64)
    g = lib().cpu_ref_hashagg_sum_count(_ptr(k), _ptr(kb), _ptr(v), _ptr(vb), n, nthreads, _ptr(ok), _ptr(okv), _ptr(os_), _ptr(osv), _ptr(oc), cap)
    if g > cap:
        raise ValueError(f"{g} groups exceed max_groups={cap}")
    return dict(k=ok[:g], k_valid=okv[:g].astype(bool), sum=os_[:g], sum_valid=osv[:g].astype(bool), count=oc[:g])


def hashagg_time_only(k, v, nthreads):
    """run the aggregation without materialising outputs (timing)"""
    return lib().cpu_ref_hashagg_sum_count(_ptr(k), None, _ptr(v), None, len(k), nthreads, None, None, None, None, None, 0)


def filter_project(a, b, thr, a_valid=None, b_valid=None, nthreads=1):
    """Filter[a < thr] -> Project[a, a + b] -> (a_out, c_out, c_valid)"""
    a = np.ascontiguousarray(a, np.int64); b = np.ascontiguousarray(b, np.int64)
    n = len(a)
    oa = np.empty(n, np.int64); oc = np.empty(n, np.int64); ocv = np.empty(n, np.uint8)
    m = lib().cpu_ref_filter_project(_ptr(a), _ptr(_bits(a_valid)), _ptr(b), _ptr(_bits(b_valid)), n, int(thr), nthreads, _ptr(oa), _ptr(oc), _ptr(ocv))
    return oa[:m], oc[:m], ocv[:m].astype(bool)


def q1_filter_agg(f, k1, k2, v, lo, hi, nthreads=1, max_groups=None):
    """Filter[lo <= f <= hi] -> SUM(v) GROUP BY k1, k2 -> dict(k1, k2, sum)"""
    f, k1, k2, v = (np.ascontiguousarray(a, np.int64) for a in (f, k1, k2, v))
    n = len(f)
    cap = int(max_groups if max_groups is not None else n)
    o1 = np.empty(cap, np.int64); o2 = np.empty(cap, np.int64); os_ = np.empty(cap, np.int64)
    g = lib().cpu_ref_q1_filter_agg(_ptr(f), _ptr(k1), _ptr(k2), _ptr(v), n, int(lo), int(hi), nthreads, _ptr(o1), _ptr(o2), _ptr(os_), cap)
    if g > cap:
        raise ValueError(f"{g} groups exceed max_groups={cap}")
    return dict(k1=o1[:g], k2=o2[:g], sum=os_[:g])


def q1_time_only(f, k1, k2, v, lo, hi, nthreads):
    return lib().cpu_ref_q1_filter_agg(_ptr(f), _ptr(k1), _ptr(k2), _ptr(v), len(f), int(lo), int(hi), nthreads, None, None, None, 0)


def filter_project_time_only(a, b, thr, nthreads, scratch=None):
    """M0 timing: outputs go to caller-provided (or fresh) scratch arrays"""
    n = len(a)
    oa, oc = scratch if scratch is not None else (np.empty(n, np.int64), np.empty(n, np.int64))
    return lib().cpu_ref_filter_project(_ptr(a), None, _ptr(b), None, n, int(thr), nthreads, _ptr(oa), _ptr(oc), None)
