"""ctypes access to tests/hostsim/libhostsim.so (TEST INFRASTRUCTURE): the device logic of
csrc/imomd_core.cuh compiled for the host with a one-thread Cta policy."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")
SO = os.path.join(ROOT, "tests", "hostsim", "libhostsim.so")
CSRC = os.path.join(ROOT, "imomd-rrtstar_b200", "csrc")

u32p = np.ctypeslib.ndpointer(np.uint32, flags="C")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C")


class Report(C.Structure):
    _fields_ = [("status", C.c_int32), ("active_trees", C.c_int32), ("iterations", C.c_int64),
                ("expansions", C.c_int64), ("tree_vertices", C.c_int64),
                ("words_consumed", C.c_int64), ("connected", C.c_int32), ("dm_updated", C.c_int32),
                ("near_ties", C.c_int64), ("unresolved_ties", C.c_int64), ("log_entries", C.c_int64),
                ("certified_ties", C.c_int64), ("tree_hash", C.c_uint64), ("busy_cycles", C.c_uint64)]


def _stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("imomd_core.cuh", "imomd_layout.h", "imomd_setup.h")]
    return any(os.path.getmtime(d) > t for d in deps)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if _stale():
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                                   "-Wno-unused-function", "-o", SO, SRC])
        L = C.CDLL(SO)
        vp = C.c_void_p
        L.hs_create.restype = vp
        L.hs_create.argtypes = [C.c_uint32, u32p, u32p, f64p, f64p, f64p, C.c_int, C.c_int, u32p, f64p,
                                C.c_double, C.c_int, C.c_int, C.c_int, C.c_double, C.c_char_p, C.c_int]
        L.hs_destroy.argtypes = [vp]
        L.hs_grid_dim.argtypes = [vp]
        L.hs_feed_words.argtypes = [vp, u32p, C.c_size_t]
        L.hs_expand.argtypes = [vp, C.c_int, C.POINTER(Report)]
        L.hs_expand_batch.argtypes = [vp, C.c_int, C.c_int, C.POINTER(Report)]
        L.hs_get_tree.argtypes = [vp, C.c_int, C.c_int, u32p, f64p]
        L.hs_get_state.argtypes = [vp, C.c_int, f64p, u32p, f64p, u32p, f64p, i32p, i32p, i32p]
        L.hs_get_frontier.restype = C.c_uint32
        L.hs_get_frontier.argtypes = [vp, C.c_int, C.c_int, u32p]
        L.hs_check_hash.argtypes = [vp, C.c_int, C.c_int]
        L.hs_nearest.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_double,
                                 C.POINTER(C.c_uint32), C.POINTER(C.c_double), i64p]
        L.hs_attach_list.argtypes = [vp, C.c_int, C.c_int, u32p, C.c_uint32, u32p]
        L.hs_visit_order.restype = C.c_uint32
        L.hs_visit_order.argtypes = [vp, C.c_int, C.c_int, u32p]
        L.hs_treehash.argtypes = [vp, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.hs_certified.restype = C.c_uint64
        L.hs_certified.argtypes = [vp]
        L.hs_set_certify.argtypes = [vp, C.c_int]
        L.hs_connection_log.restype = C.c_uint32
        L.hs_connection_log.argtypes = [vp, C.c_int, u32p, u32p, C.c_uint32]
        _lib = L
    return _lib


class HostSimPlanner:
    def __init__(self, g, dests, probs, goal_bias=1.0, pseudo=False, grid_dim=0, slots=0, tie_band=0.0):
        """dests: list of per-query destination lists; probs: list of KxK matrices."""
        L = lib()
        self.L, self.g = L, g
        self.Q, self.K = len(dests), len(dests[0])
        d = np.ascontiguousarray(np.asarray(dests, np.uint32).ravel())
        pr = np.ascontiguousarray(np.asarray(probs, np.float64).ravel())
        err = C.create_string_buffer(256)
        self.h = L.hs_create(g.n, g.row_ptr, g.col, g.w, g.lat, g.lon, self.Q, self.K, d, pr,
                             goal_bias, int(pseudo), grid_dim, slots, tie_band, err, 256)
        if not self.h:
            raise ValueError(err.value.decode())

    def close(self):
        self.L.hs_destroy(self.h)

    def feed(self, words):
        w = np.ascontiguousarray(words, np.uint32)
        self.L.hs_feed_words(self.h, w, w.size)

    def expand(self, iters):
        rep = (Report * self.Q)()
        self.L.hs_expand(self.h, iters, rep)
        return list(rep)

    def expand_batch(self, steps, B):
        rep = (Report * self.Q)()
        self.L.hs_expand_batch(self.h, steps, B, rep)
        return list(rep)

    def tree(self, k, q=0):
        par = np.empty(self.g.n, np.uint32); cost = np.empty(self.g.n)
        self.L.hs_get_tree(self.h, q, k, par, cost)
        return par, cost

    def state(self, q=0):
        K = self.K
        D = np.empty(K * K); cn = np.empty(K * K, np.uint32); P = np.empty(K * K)
        mi = np.empty(K * K, np.uint32); mv = np.empty(K * K)
        done = np.empty(K, np.int32); ds = np.empty(K, np.int32); fl = np.empty(2, np.int32)
        self.L.hs_get_state(self.h, q, D, cn, P, mi, mv, done, ds, fl)
        return dict(D=D.reshape(K, K), conn=cn.reshape(K, K), P=P.reshape(K, K),
                    mh_id=mi.reshape(K, K), mh_val=mv.reshape(K, K), is_done=done, ds_parent=ds,
                    connected=int(fl[0]), dm_updated=int(fl[1]))

    def frontier(self, k, q=0):
        out = np.empty(self.g.n, np.uint32)
        c = self.L.hs_get_frontier(self.h, q, k, out)
        return np.sort(out[:c]).astype(np.uint64)

    def treehash(self, q=0):
        h, sm = C.c_uint64(), C.c_uint64()
        self.L.hs_treehash(self.h, q, C.byref(h), C.byref(sm))
        return h.value, sm.value

    def certified(self):
        return int(self.L.hs_certified(self.h))

    def set_certify(self, enable):
        self.L.hs_set_certify(self.h, int(enable))

    def check_hash(self, k, q=0):
        return self.L.hs_check_hash(self.h, q, k)

    def nearest(self, k, lat, lon, q=0):
        idx = C.c_uint32(); d = C.c_double(); st = np.zeros(3, np.int64)
        self.L.hs_nearest(self.h, q, k, lat, lon, C.byref(idx), C.byref(d), st)
        return idx.value, d.value, st

    def attach_list(self, k, vertices, q=0):
        v = np.ascontiguousarray(vertices, np.uint32)
        out = np.empty(v.size, np.uint32)
        self.L.hs_attach_list(self.h, q, k, v, v.size, out)
        return out

    def visit_order(self, k, q=0):
        out = np.empty(self.g.n, np.uint32)
        c = self.L.hs_visit_order(self.h, q, k, out)
        return out[:c].copy()

    def connection_log(self, q=0):
        cap = 1 << 20
        si = np.empty(cap, np.uint32); vx = np.empty(cap, np.uint32)
        c = self.L.hs_connection_log(self.h, q, si, vx, cap)
        return si[:c].copy(), vx[:c].copy()
