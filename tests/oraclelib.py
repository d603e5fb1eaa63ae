"""ctypes access to the two checkers (TEST INFRASTRUCTURE):

* ``oracle/liboracle.so``          -- the CPU restatement (oracle/imomd_oracle.cpp)
* ``oracle/_ref/libimomd_ref.so``  -- the unmodified reference behind oracle/ref_shim.cpp

Only tests, ``__graft_entry__.smoke()`` and ``bench.py``'s baseline legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libimomd_ref.so")
REF_PROBE = os.path.join(ORACLE_DIR, "_ref", "ref_probe")

u32p = np.ctypeslib.ndpointer(np.uint32, flags="C")
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C")


class Trace(C.Structure):
    _fields_ = [
        ("tree", C.c_int32), ("skipped", C.c_int32), ("rand_id", C.c_uint64),
        ("rand_lat", C.c_double), ("rand_lon", C.c_double), ("x_nearest", C.c_uint64),
        ("nearest_dist", C.c_double), ("x_new", C.c_uint64), ("parent_of_new", C.c_uint64),
        ("cost_of_new", C.c_double), ("tree_size_after", C.c_uint64),
        ("frontier_after", C.c_uint64),
    ]

    def astuple(self):
        return tuple(getattr(self, f) for f, _ in self._fields_)


def build_oracle(force: bool = False) -> None:
    if force or not os.path.exists(ORACLE_SO) or (
        os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(ORACLE_DIR, "imomd_oracle.cpp"))
    ):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)


def have_ref() -> bool:
    return os.path.exists(REF_SO) and os.path.exists(REF_PROBE)


def _sig(fn, res, args):
    fn.restype = res
    fn.argtypes = args
    return fn


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        build_oracle()
        L = C.CDLL(ORACLE_SO)
        vp = C.c_void_p
        _sig(L.orc_haversine, C.c_double, [C.c_double] * 4)
        _sig(L.orc_bearing, C.c_double, [C.c_double] * 4)
        _sig(L.orc_mt19937_words, None, [C.c_uint32, C.c_uint64, u32p])
        _sig(L.orc_uniform_real, None, [C.c_uint32, C.c_double, C.c_double, C.c_uint64, f64p])
        _sig(L.orc_uniform_size_t, None, [C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, u64p])
        _sig(L.orc_uniform_int, None, [C.c_uint32, C.c_int, C.c_int, C.c_uint64, i32p])
        _sig(L.orc_graph_create, vp, [C.c_uint32, u32p, u32p, f64p, f64p, f64p])
        _sig(L.orc_graph_destroy, None, [vp])
        _sig(L.orc_native_csr_from_edges, None,
             [C.c_uint32, C.c_uint64, u64p, u64p, f64p, u32p, u32p, f64p])
        _sig(L.orc_planner_create, vp, [vp, C.c_int, u64p, C.c_double, C.c_int])
        _sig(L.orc_planner_destroy, None, [vp])
        _sig(L.orc_planner_iterate, C.c_uint64, [vp, C.c_int])
        _sig(L.orc_planner_iterate_batched, C.c_uint64, [vp, C.c_int, C.c_int])
        _sig(L.orc_planner_expand_traced, C.c_int, [vp, C.c_int, C.POINTER(Trace)])
        _sig(L.orc_planner_get_tree, None, [vp, C.c_int, u32p, f64p])
        _sig(L.orc_planner_tree_size, C.c_uint64, [vp, C.c_int])
        _sig(L.orc_planner_children, C.c_int, [vp, C.c_int, C.c_uint64, u64p, C.c_int])
        _sig(L.orc_planner_treehash, C.c_uint64, [vp])
        _sig(L.orc_planner_get_matrices, None, [vp, f64p, u64p, f64p, u64p, f64p])
        _sig(L.orc_planner_get_flags, None, [vp, i32p, i32p, i32p, i32p])
        _sig(L.orc_planner_frontier, C.c_uint64, [vp, C.c_int, u64p, C.c_uint64])
        _sig(L.orc_planner_nearest, C.c_int,
             [vp, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_uint64), C.POINTER(C.c_double),
              C.POINTER(C.c_int32), C.POINTER(C.c_double)])
        _sig(L.orc_planner_tie_count, C.c_uint64, [vp])
        _sig(L.orc_planner_connection_log, C.c_uint64, [vp, u32p, u64p, C.c_uint64])
        _sig(L.orc_solver_create, vp, [C.c_int] * 5)
        _sig(L.orc_solver_destroy, None, [vp])
        _sig(L.orc_solver_solve, C.c_int,
             [vp, C.c_int, f64p, C.c_int, C.c_int, C.POINTER(C.c_double), i32p, C.c_int])
        _sig(L.orc_path_costs, None, [C.c_int, f64p, C.c_int, C.c_int, i32p, i32p, f64p])
        _oracle = L
    return _oracle


def ref():
    global _ref
    if _ref is None:
        L = C.CDLL(REF_SO)
        vp = C.c_void_p
        _sig(L.ref_graph_from_edges, vp, [C.c_uint64, f64p, f64p, C.c_uint64, u64p, u64p, C.c_void_p])
        _sig(L.ref_graph_grid, vp, [C.c_int])
        _sig(L.ref_graph_from_osm, vp, [C.c_char_p])
        _sig(L.ref_graph_fake, vp, [C.c_int])
        _sig(L.ref_graph_num_nodes, C.c_uint64, [vp])
        _sig(L.ref_graph_num_arcs, C.c_uint64, [vp])
        _sig(L.ref_graph_export, None, [vp, f64p, f64p, u64p, u32p, u32p, f64p])
        _sig(L.ref_haversine, C.c_double, [C.c_double] * 4)
        _sig(L.ref_bearing, C.c_double, [C.c_double] * 4)
        _sig(L.ref_pick_destinations, None, [vp, C.c_uint32, C.c_int, u64p])
        _sig(L.ref_graph_components, C.c_uint64, [vp, u32p])
        _sig(L.ref_planner_create, vp,
             [vp, C.c_uint64, C.c_uint64, u64p, C.c_int, C.c_double] + [C.c_int] * 8)
        _sig(L.ref_planner_K, C.c_int, [vp])
        _sig(L.ref_planner_iterate, C.c_uint64, [vp, C.c_int])
        _sig(L.ref_planner_expand_traced, C.c_int, [vp, C.c_int, C.POINTER(Trace)])
        _sig(L.ref_planner_get_tree, None, [vp, C.c_int, u32p, f64p])
        _sig(L.ref_planner_tree_size, C.c_uint64, [vp, C.c_int])
        _sig(L.ref_planner_children, C.c_int, [vp, C.c_int, C.c_uint64, u64p, C.c_int])
        _sig(L.ref_planner_treehash, C.c_uint64, [vp])
        _sig(L.ref_planner_get_matrices, None, [vp, f64p, u64p, f64p, u64p, f64p])
        _sig(L.ref_planner_get_flags, None, [vp, i32p, i32p, i32p, i32p])
        _sig(L.ref_planner_frontier, C.c_uint64, [vp, C.c_int, u64p, C.c_uint64])
        _sig(L.ref_planner_connection_set, C.c_uint64, [vp, C.c_int, u64p, C.c_uint64])
        _sig(L.ref_planner_nearest, C.c_int,
             [vp, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_uint64), C.POINTER(C.c_double),
              C.POINTER(C.c_int32), C.POINTER(C.c_double)])
        _sig(L.ref_planner_attach, None, [vp, C.c_int, C.c_uint64])
        _sig(L.ref_planner_parent_order, C.c_uint64, [vp, C.c_int, u64p, C.c_uint64])
        _sig(L.ref_planner_solve_rtsp, C.c_int, [vp])
        _sig(L.ref_planner_path_cost, C.c_double, [vp])
        _sig(L.ref_planner_path, C.c_uint64, [vp, u64p, C.c_uint64])
        _sig(L.ref_planner_sequence, C.c_int, [vp, i32p, C.c_int])
        _sig(L.ref_solver_create, vp, [C.c_int] * 5)
        _sig(L.ref_solver_solve, C.c_int,
             [vp, C.c_int, f64p, C.c_int, C.c_int, C.POINTER(C.c_double), i32p, C.c_int])
        _sig(L.ref_path_costs, None, [C.c_int, f64p, C.c_int, C.c_int, i32p, i32p, f64p])
        _sig(L.ref_mt19937_words, None, [C.c_uint32, C.c_uint64, u32p])
        _sig(L.ref_uniform_real, None, [C.c_uint32, C.c_double, C.c_double, C.c_uint64, f64p])
        _sig(L.ref_uniform_size_t, None, [C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, u64p])
        _sig(L.ref_uniform_int, None, [C.c_uint32, C.c_int, C.c_int, C.c_uint64, i32p])
        _sig(L.ref_toolchain, C.c_char_p, [])
        _ref = L
    return _ref


# ------------------------------------------------------------------ helpers ---

def ref_export_graph(gh, name="ref"):
    """RoadGraph (native-order CSR) from a reference graph handle."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "imomd-rrtstar_b200"))
    from imomd_b200.graphs import RoadGraph
    R = ref()
    n = R.ref_graph_num_nodes(gh)
    m = R.ref_graph_num_arcs(gh)
    lat = np.empty(n); lon = np.empty(n); ids = np.empty(n, np.uint64)
    row_ptr = np.empty(n + 1, np.uint32); col = np.empty(m, np.uint32); w = np.empty(m)
    R.ref_graph_export(gh, lat, lon, ids, row_ptr, col, w)
    assert (ids == np.arange(n, dtype=np.uint64)).all()
    return RoadGraph(lat=lat, lon=lon, row_ptr=row_ptr, col=col, w=w, name=name)


class OraclePlanner:
    """Restatement planner on a RoadGraph."""

    def __init__(self, g, dest, goal_bias=1.0, pseudo=False):
        L = oracle()
        self.L, self.g, self.K = L, g, len(dest)
        self.gh = L.orc_graph_create(g.n, g.row_ptr, g.col, g.w, g.lat, g.lon)
        self.ph = L.orc_planner_create(self.gh, self.K, np.asarray(dest, np.uint64), goal_bias,
                                       int(pseudo))

    def close(self):
        self.L.orc_planner_destroy(self.ph)
        self.L.orc_graph_destroy(self.gh)

    def iterate(self, iters):
        return self.L.orc_planner_iterate(self.ph, iters)

    def iterate_batched(self, steps, B):
        """the batched mode's semantics (imomd_expand_batch): steps x (B expansions per tree)"""
        return self.L.orc_planner_iterate_batched(self.ph, steps, B)

    def expand_traced(self, k):
        t = Trace()
        self.L.orc_planner_expand_traced(self.ph, k, C.byref(t))
        return t

    def tree(self, k):
        par = np.empty(self.g.n, np.uint32); cost = np.empty(self.g.n)
        self.L.orc_planner_get_tree(self.ph, k, par, cost)
        return par, cost

    def treehash(self):
        return self.L.orc_planner_treehash(self.ph)

    def matrices(self):
        K = self.K
        D = np.empty(K * K); cn = np.empty(K * K, np.uint64); P = np.empty(K * K)
        mi = np.empty(K * K, np.uint64); mv = np.empty(K * K)
        self.L.orc_planner_get_matrices(self.ph, D, cn, P, mi, mv)
        return dict(D=D.reshape(K, K), conn=cn.reshape(K, K), P=P.reshape(K, K),
                    mh_id=mi.reshape(K, K), mh_val=mv.reshape(K, K))

    def flags(self):
        K = self.K
        done = np.empty(K, np.int32); ds = np.empty(K, np.int32)
        c = np.empty(1, np.int32); u = np.empty(1, np.int32)
        self.L.orc_planner_get_flags(self.ph, done, c, u, ds)
        return dict(is_done=done, connected=int(c[0]), dm_updated=int(u[0]), ds_parent=ds)

    def frontier(self, k):
        out = np.empty(self.g.n, np.uint64)
        c = self.L.orc_planner_frontier(self.ph, k, out, self.g.n)
        return np.sort(out[:c])

    def nearest(self, k, lat, lon):
        idx = C.c_uint64(); d = C.c_double(); t = C.c_int32(); s = C.c_double()
        ok = self.L.orc_planner_nearest(self.ph, k, lat, lon, C.byref(idx), C.byref(d),
                                        C.byref(t), C.byref(s))
        return ok, idx.value, d.value, t.value, s.value

    def ties(self):
        return self.L.orc_planner_tie_count(self.ph)

    def connection_log(self):
        cap = 1 << 20
        si = np.empty(cap, np.uint32); vx = np.empty(cap, np.uint64)
        c = self.L.orc_planner_connection_log(self.ph, si, vx, cap)
        return si[:c].copy(), vx[:c].copy()


class RefPlanner:
    """The unmodified reference planner on a reference graph handle.

    NOTE: goal_bias < 1 uses a function-local static distribution whose range is
    frozen by the first planner of the process -- use ref_probe (a subprocess) or
    keep one graph size per process for such runs.
    """

    def __init__(self, gh, dest, goal_bias=1.0, pseudo=False, max_iter=1000000, max_time=60,
                 ga=(10000, 1000, 5)):
        R = ref()
        self.R, self.gh, self.K = R, gh, len(dest)
        self.n = R.ref_graph_num_nodes(gh)
        obj = np.asarray(dest[1:-1], np.uint64)
        if obj.size == 0:
            obj = np.zeros(1, np.uint64)
        self.ph = R.ref_planner_create(gh, dest[0], dest[-1], obj, len(dest) - 2, goal_bias,
                                       int(pseudo), max_iter, max_time, 1, 1, *ga)

    def iterate(self, iters):
        return self.R.ref_planner_iterate(self.ph, iters)

    def expand_traced(self, k):
        t = Trace()
        self.R.ref_planner_expand_traced(self.ph, k, C.byref(t))
        return t

    def tree(self, k):
        par = np.empty(self.n, np.uint32); cost = np.empty(self.n)
        self.R.ref_planner_get_tree(self.ph, k, par, cost)
        return par, cost

    def treehash(self):
        return self.R.ref_planner_treehash(self.ph)

    def matrices(self):
        K = self.K
        D = np.empty(K * K); cn = np.empty(K * K, np.uint64); P = np.empty(K * K)
        mi = np.empty(K * K, np.uint64); mv = np.empty(K * K)
        self.R.ref_planner_get_matrices(self.ph, D, cn, P, mi, mv)
        return dict(D=D.reshape(K, K), conn=cn.reshape(K, K), P=P.reshape(K, K),
                    mh_id=mi.reshape(K, K), mh_val=mv.reshape(K, K))

    def flags(self):
        K = self.K
        done = np.empty(K, np.int32); ds = np.empty(K, np.int32)
        c = np.empty(1, np.int32); u = np.empty(1, np.int32)
        self.R.ref_planner_get_flags(self.ph, done, c, u, ds)
        return dict(is_done=done, connected=int(c[0]), dm_updated=int(u[0]), ds_parent=ds)

    def frontier(self, k):
        out = np.empty(self.n, np.uint64)
        c = self.R.ref_planner_frontier(self.ph, k, out, self.n)
        return np.sort(out[:c])

    def nearest(self, k, lat, lon):
        idx = C.c_uint64(); d = C.c_double(); t = C.c_int32(); s = C.c_double()
        ok = self.R.ref_planner_nearest(self.ph, k, lat, lon, C.byref(idx), C.byref(d),
                                        C.byref(t), C.byref(s))
        return ok, idx.value, d.value, t.value, s.value

    def children(self, k, v):
        out = np.empty(64, np.uint64)
        c = self.R.ref_planner_children(self.ph, k, v, out, 64)
        return None if c < 0 else [int(x) for x in out[:c]]

    def connection_set(self, idx):
        out = np.empty(self.n, np.uint64)
        c = self.R.ref_planner_connection_set(self.ph, idx, out, self.n)
        return [int(x) for x in out[:c]]

    def attach(self, k, x):
        self.R.ref_planner_attach(self.ph, k, int(x))

    def parent_order(self, k):
        out = np.empty(self.n, np.uint64)
        c = self.R.ref_planner_parent_order(self.ph, k, out, self.n)
        return out[:c].copy()

    def solve_rtsp(self):
        return self.R.ref_planner_solve_rtsp(self.ph)

    def path_cost(self):
        return self.R.ref_planner_path_cost(self.ph)
