"""ctypes loader for oracle/per_oracle.c (TEST INFRASTRUCTURE -- see the header of that file).

PARITY UNPINNED: restates torchrl's published segment-tree / PrioritizedSampler algorithm; torchrl is
absent from the reference tree and from this image, and the reference has no golden vectors for it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libper_oracle.so")
_lib = None


def build():
    src = os.path.join(_HERE, "per_oracle.c")
    if os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(src):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-Wall", "-ffp-contract=off", "-fno-fast-math",
                           "-o", _SO, src, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.po_create.restype = C.c_void_p
        L.po_create.argtypes = [C.c_int64, C.c_float, C.c_double, C.c_int, C.c_int, C.c_int]
        L.po_destroy.argtypes = [C.c_void_p]
        for name in ("po_capacity", "po_len", "po_cursor"):
            getattr(L, name).restype = C.c_int64
            getattr(L, name).argtypes = [C.c_void_p]
        L.po_max_priority.restype = C.c_double
        L.po_max_priority.argtypes = [C.c_void_p]
        L.po_sum_ptr.restype = C.POINTER(C.c_float)
        L.po_sum_ptr.argtypes = [C.c_void_p]
        L.po_min_ptr.restype = C.POINTER(C.c_float)
        L.po_min_ptr.argtypes = [C.c_void_p]
        L.po_set_len.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
        L.po_set_max_priority.argtypes = [C.c_void_p, C.c_double]
        L.po_update_leaves.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.po_query_sum.restype = C.c_float
        L.po_query_sum.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
        L.po_query_min.restype = C.c_float
        L.po_query_min.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
        L.po_scan_lower_bound.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.po_sample.restype = C.c_int
        L.po_sample.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p]
        L.po_default_priority.restype = C.c_float
        L.po_default_priority.argtypes = [C.c_void_p]
        L.po_extend.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.po_update_priority.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.po_build.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.po_nstep.argtypes = [C.c_int64, C.c_int, C.c_double] + [C.c_void_p] * 5 + [C.c_int64] + [C.c_void_p] * 6
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleTree:
    """Single-threaded CPU sum/min segment tree + sampler with torchrl's semantics."""

    def __init__(self, size, alpha=0.5, eps=1e-8, strict_pow2=False, weight_eps_in_denominator=False,
                 default_priority_fp64=True):
        self.L = lib()
        self.size = int(size)
        self.h = self.L.po_create(self.size, alpha, eps, int(strict_pow2), int(weight_eps_in_denominator),
                                  int(default_priority_fp64))
        if not self.h:
            raise MemoryError
        self.capacity = self.L.po_capacity(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.po_destroy(self.h)
            self.h = None

    @property
    def sum(self):
        return np.ctypeslib.as_array(self.L.po_sum_ptr(self.h), shape=(2 * self.capacity,))

    @property
    def min(self):
        return np.ctypeslib.as_array(self.L.po_min_ptr(self.h), shape=(2 * self.capacity,))

    def __len__(self):
        return self.L.po_len(self.h)

    @property
    def cursor(self):
        return self.L.po_cursor(self.h)

    @property
    def max_priority(self):
        return self.L.po_max_priority(self.h)

    def set_len(self, n, cursor=None):
        self.L.po_set_len(self.h, int(n), int(n % self.size if cursor is None else cursor))

    def build(self, leaves):
        leaves = np.ascontiguousarray(leaves, dtype=np.float32)
        self.L.po_build(self.h, leaves.size, _p(leaves))

    def update_leaves(self, idx, vals):
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        vals = np.ascontiguousarray(vals, dtype=np.float32)
        self.L.po_update_leaves(self.h, idx.size, _p(idx), _p(vals))

    def update_priority(self, idx, prio):
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        prio = np.ascontiguousarray(prio, dtype=np.float32)
        self.L.po_update_priority(self.h, idx.size, _p(idx), _p(prio))

    def extend(self, n):
        out = np.empty(n, dtype=np.int64)
        self.L.po_extend(self.h, int(n), _p(out))
        return out

    def default_priority(self):
        return self.L.po_default_priority(self.h)

    def query_sum(self, lo, hi):
        return self.L.po_query_sum(self.h, int(lo), int(hi))

    def query_min(self, lo, hi):
        return self.L.po_query_min(self.h, int(lo), int(hi))

    def scan(self, mass):
        mass = np.ascontiguousarray(mass, dtype=np.float32)
        out = np.empty(mass.size, dtype=np.int64)
        self.L.po_scan_lower_bound(self.h, mass.size, _p(mass), _p(out))
        return out

    def sample(self, u, beta=0.5, mode=0):
        """u: fp64 uniforms in [0,1).  Returns (idx int64, weight fp32, mass fp32, p_sum, p_min)."""
        u = np.ascontiguousarray(u, dtype=np.float64)
        idx = np.empty(u.size, dtype=np.int64)
        w = np.empty(u.size, dtype=np.float32)
        mass = np.empty(u.size, dtype=np.float32)
        ps, pm = C.c_float(0), C.c_float(0)
        rc = self.L.po_sample(self.h, u.size, _p(u), int(mode), float(beta), _p(idx), _p(w), _p(mass),
                              C.byref(ps), C.byref(pm))
        if rc == -1:
            raise RuntimeError("Cannot sample from an empty storage.")
        if rc == -2:
            raise RuntimeError("non-positive p_sum")
        if rc == -3:
            raise RuntimeError("non-positive p_min")
        return idx, w, mass, ps.value, pm.value


def nstep_arrays(size, n_step, gamma, slot_seq, next_link, reward, done, trunc, start):
    """Array restatement of _compute_n_step over the SoA ring (see po_nstep in per_oracle.c)."""
    L = lib()
    start = np.ascontiguousarray(start, dtype=np.int64)
    n = start.size
    ret = np.empty(n, np.float32); gam = np.empty(n, np.float32); dn = np.empty(n, np.uint8)
    last = np.empty(n, np.int64); succ = np.empty(n, np.int64)
    a = [np.ascontiguousarray(slot_seq, np.int64), np.ascontiguousarray(next_link, np.int64),
         np.ascontiguousarray(reward, np.float32), np.ascontiguousarray(done, np.uint8),
         np.ascontiguousarray(trunc, np.uint8)]
    L.po_nstep(int(size), int(n_step), float(gamma), *[_p(x) for x in a], n, _p(start), _p(ret), _p(gam), _p(dn),
               _p(last), _p(succ))
    return ret, gam, dn, last, succ
