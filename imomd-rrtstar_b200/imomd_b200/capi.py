"""ctypes binding of the C ABI (include/imomd_b200.h) and of the host library.

This is plumbing: numpy arrays in, numpy arrays out, every call goes through
libimomd_b200.so (CUDA).  Loading fails loudly when the library is missing and every
compute call fails loudly (ImomdError) without a CUDA device -- there is no fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import time
import numpy as np

PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# IMOMD_B200_LIB: developer override (kernel variants built by tools/); the product path is lib/
LIB_CUDA = os.environ.get("IMOMD_B200_LIB") or os.path.join(PKG, "lib", "libimomd_b200.so")
LIB_HOST = os.path.join(PKG, "lib", "libimomd_host.so")

NONE = 0xFFFFFFFF
KMAX = 32

u32p = np.ctypeslib.ndpointer(np.uint32, flags="C")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C")


class ImomdError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [("goal_bias", C.c_double), ("pseudo_mode", C.c_int32), ("max_resident_queries", C.c_int32),
                ("frontier_chunks", C.c_uint32), ("arena_entries", C.c_uint32),
                ("log_entries", C.c_uint32), ("reserved", C.c_uint32), ("tie_band", C.c_double)]


class Report(C.Structure):
    _fields_ = [("status", C.c_int32), ("active_trees", C.c_int32), ("iterations", C.c_int64),
                ("expansions", C.c_int64), ("tree_vertices", C.c_int64),
                ("words_consumed", C.c_int64), ("connected", C.c_int32), ("dm_updated", C.c_int32),
                ("near_ties", C.c_int64), ("unresolved_ties", C.c_int64), ("log_entries", C.c_int64),
                ("certified_ties", C.c_int64), ("tree_hash", C.c_uint64), ("busy_cycles", C.c_uint64)]


class NNQuery(C.Structure):
    _fields_ = [("query", C.c_int32), ("tree", C.c_int32), ("lat", C.c_double), ("lon", C.c_double)]


NN_DTYPE = np.dtype([("query", np.int32), ("tree", np.int32), ("lat", np.float64), ("lon", np.float64)])

_cuda = None
_host = None


def cuda_lib():
    global _cuda
    if _cuda is None:
        if not os.path.exists(LIB_CUDA):
            raise ImomdError(f"{LIB_CUDA} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`"
                             " (there is no CPU fallback)")
        L = C.CDLL(LIB_CUDA)
        vp, vpp = C.c_void_p, C.POINTER(C.c_void_p)
        L.imomd_last_error.restype = C.c_char_p
        L.imomd_version.restype = C.c_char_p
        L.imomd_sizeof_query_state.restype = C.c_size_t
        L.imomd_graph_create.argtypes = [C.c_int, C.c_uint32, u32p, u32p, f64p, f64p, f64p, vpp]
        L.imomd_graph_create_ex.argtypes = [C.c_int, C.c_uint32, u32p, u32p, f64p, f64p, f64p, C.c_int, vpp]
        L.imomd_graph_destroy.argtypes = [vp]
        L.imomd_graph_info.argtypes = [vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                       C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.imomd_planner_create.argtypes = [vp, C.c_int, C.c_int, u32p, f64p, C.POINTER(Params), vpp]
        L.imomd_planner_destroy.argtypes = [vp]
        L.imomd_planner_reset.argtypes = [vp]
        L.imomd_planner_set_queries.argtypes = [vp, u32p, f64p]
        L.imomd_planner_clear_words.argtypes = [vp]
        L.imomd_get_counters.argtypes = [vp, np.ctypeslib.ndpointer(np.uint64, flags="C")]
        L.imomd_get_tie_info.argtypes = [vp, C.c_int, f64p]
        L.imomd_planner_feed_words.argtypes = [vp, u32p, C.c_size_t]
        L.imomd_planner_words_available.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.imomd_expand.argtypes = [vp, C.c_int, C.POINTER(Report)]
        L.imomd_expand_async.argtypes = [vp, C.c_int]
        L.imomd_expand_batch.argtypes = [vp, C.c_int, C.c_int, C.POINTER(Report)]
        L.imomd_expand_batch_async.argtypes = [vp, C.c_int, C.c_int]
        L.imomd_set_time_budget.argtypes = [vp, C.c_double]
        L.imomd_sync.argtypes = [vp, C.POINTER(Report)]
        L.imomd_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_float)]
        L.imomd_get_state.argtypes = [vp, C.c_int, f64p, u32p, f64p, u32p, f64p, i32p, i32p]
        L.imomd_get_states.argtypes = [vp, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.imomd_get_treehash.argtypes = [vp, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.imomd_planner_info.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                         C.POINTER(C.c_uint64)]
        L.imomd_clear_dm_updated.argtypes = [vp, C.c_int]
        L.imomd_get_visit_order.argtypes = [vp, C.c_int, C.c_int, u32p, C.c_uint32, C.POINTER(C.c_uint32)]
        L.imomd_set_profiling.argtypes = [vp, C.c_int]
        L.imomd_get_profile.argtypes = [vp, C.c_int, np.ctypeslib.ndpointer(np.uint64, flags="C"),
                                        np.ctypeslib.ndpointer(np.uint64, flags="C")]
        L.imomd_get_phase_cycles.argtypes = [vp, C.c_int, np.ctypeslib.ndpointer(np.uint64, flags="C"),
                                             np.ctypeslib.ndpointer(np.uint64, flags="C")]
        L.imomd_get_flags.argtypes = [vp, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.imomd_get_tree.argtypes = [vp, C.c_int, C.c_int, u32p, f64p]
        L.imomd_get_frontier.argtypes = [vp, C.c_int, C.c_int, u32p, C.c_uint32, C.POINTER(C.c_uint32)]
        L.imomd_walk_to_root.argtypes = [vp, C.c_int, C.c_int, C.c_uint32, u32p, C.c_uint32,
                                         C.POINTER(C.c_uint32)]
        L.imomd_get_connection_log.argtypes = [vp, C.c_int, u32p, u32p, C.c_uint32, C.POINTER(C.c_uint32)]
        L.imomd_nearest.argtypes = [vp, C.c_int, C.c_void_p, u32p, f64p]
        L.imomd_nearest_batch.argtypes = [vp, C.c_int, C.c_void_p, u32p, f64p, C.POINTER(C.c_float),
                                          C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
        L.imomd_nearest_batch_last_timing.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                                      C.POINTER(C.c_float), C.POINTER(C.c_uint32)]
        L.imomd_ctx_create.argtypes = [C.c_int, vpp]
        L.imomd_ctx_destroy.argtypes = [vp]
        L.imomd_ga_eval.argtypes = [vp, C.c_int, f64p, C.c_int, C.c_int, i32p, i32p, f64p]
        _cuda = L
    return _cuda


def host_lib():
    global _host
    if _host is None:
        if not os.path.exists(LIB_HOST):
            raise ImomdError(f"{LIB_HOST} is missing: run __graft_entry__.build()")
        L = C.CDLL(LIB_HOST)
        L.imomd_host_selection_probabilities.argtypes = [C.c_int, f64p, f64p, f64p]
        L.imomd_host_great_circle_m.restype = C.c_double
        L.imomd_host_great_circle_m.argtypes = [C.c_double] * 4
        u64p = np.ctypeslib.ndpointer(np.uint64, flags="C")
        L.imomd_host_run.argtypes = [C.c_int, C.c_uint64, f64p, f64p, C.c_uint64, u64p, u64p, f64p,
                                     C.c_int, u64p, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int), u64p, C.c_uint64,
                                     C.POINTER(C.c_uint64), i32p, C.c_int, C.POINTER(C.c_int),
                                     C.POINTER(C.c_double), C.POINTER(C.c_longlong),
                                     C.POINTER(C.c_longlong), C.c_char_p, C.c_int]
        L.imomd_host_planner_create.restype = C.c_void_p
        L.imomd_host_planner_create.argtypes = [C.c_int, C.c_uint64, f64p, f64p, C.c_uint64, u64p, u64p,
                                                f64p, C.c_int, u64p, C.c_double, C.c_int, C.c_int,
                                                C.c_int, C.c_int, C.c_char_p, C.c_int]
        L.imomd_host_planner_destroy.argtypes = [C.c_void_p]
        L.imomd_host_planner_step.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_int]
        L.imomd_host_planner_native.restype = C.c_void_p
        L.imomd_host_planner_native.argtypes = [C.c_void_p]
        L.imomd_host_planner_cost.restype = C.c_double
        L.imomd_host_planner_cost.argtypes = [C.c_void_p]
        L.imomd_host_cache_roundtrip.argtypes = [C.c_char_p, C.c_uint64, f64p, f64p, C.c_uint64, u64p, u64p,
                                                 f64p, u32p, u32p, f64p, f64p, f64p, C.c_char_p, C.c_int]
        L.imomd_host_solver_create.restype = C.c_void_p
        L.imomd_host_solver_create.argtypes = [C.c_int] * 6
        L.imomd_host_solver_destroy.argtypes = [C.c_void_p]
        L.imomd_host_solver_solve.argtypes = [C.c_void_p, C.c_int, f64p, C.c_int, C.c_int,
                                              C.POINTER(C.c_double), i32p, C.c_int,
                                              C.POINTER(C.c_uint64), C.c_char_p, C.c_int]
        L.imomd_host_island_begin.argtypes = [C.c_void_p, C.c_int, f64p, C.c_int, C.c_int, C.c_uint,
                                              C.c_char_p, C.c_int]
        L.imomd_host_island_step.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.imomd_host_island_best.argtypes = [C.c_void_p, C.POINTER(C.c_double), i32p, C.c_int]
        L.imomd_host_island_accept.argtypes = [C.c_void_p, C.c_double, i32p, C.c_int]
        L.imomd_host_island_evaluated.restype = C.c_uint64
        L.imomd_host_island_evaluated.argtypes = [C.c_void_p]
        _host = L
    return _host


def sizeof_query_state() -> int:
    return int(cuda_lib().imomd_sizeof_query_state())


def _check(rc):
    if rc != 0:
        raise ImomdError(f"status {rc}: {cuda_lib().imomd_last_error().decode()}")


def selection_probabilities(g, dest):
    """K x K matrix of the planner constructor (host facade, glibc math)."""
    d = np.asarray(dest, np.int64)
    lat = np.ascontiguousarray(g.lat[d]); lon = np.ascontiguousarray(g.lon[d])
    P = np.empty(len(d) * len(d))
    host_lib().imomd_host_selection_probabilities(len(d), lat, lon, P)
    return P.reshape(len(d), len(d))


def mt19937_words(seed: int, n: int) -> np.ndarray:
    """Raw std::mt19937 words (numpy's MT19937 is the same generator; init_genrand seeding)."""
    bg = np.random.MT19937()
    st = bg.state
    key = np.empty(624, np.uint32)
    key[0] = seed & 0xFFFFFFFF
    for i in range(1, 624):
        key[i] = (1812433253 * (int(key[i - 1]) ^ (int(key[i - 1]) >> 30)) + i) & 0xFFFFFFFF
    st["state"]["key"] = key
    st["state"]["pos"] = 624
    bg.state = st
    return bg.random_raw(n).astype(np.uint32)


class DeviceGraph:
    def __init__(self, g, device=0, grid_dim=0):
        L = cuda_lib()
        self.g = g
        self.h = C.c_void_p()
        _check(L.imomd_graph_create_ex(device, g.n, g.row_ptr, g.col, g.w, g.lat, g.lon, grid_dim,
                                       C.byref(self.h)))

    def info(self):
        n, a, G, d = C.c_uint32(), C.c_uint32(), C.c_int32(), C.c_int32()
        _check(cuda_lib().imomd_graph_info(self.h, C.byref(n), C.byref(a), C.byref(G), C.byref(d)))
        return dict(n=n.value, arcs=a.value, grid_dim=G.value, device=d.value)

    def close(self):
        if self.h:
            cuda_lib().imomd_graph_destroy(self.h)
            self.h = None


class DevicePlanner:
    """Q independent queries (K trees each) on one device graph."""

    def __init__(self, dg: DeviceGraph, dests, probs=None, goal_bias=1.0, pseudo=False,
                 frontier_chunks=0, arena_entries=0, log_entries=0, max_resident=0, tie_band=0.0):
        """max_resident: 0 = every query keeps its trees in HBM; n > 0 = at most n state slots (more
        queries than that stream through them); -1 = automatic."""
        L = cuda_lib()
        self.dg, self.g = dg, dg.g
        self.Q, self.K = len(dests), len(dests[0])
        if probs is None:
            probs = [selection_probabilities(self.g, d) for d in dests]
        d = np.ascontiguousarray(np.asarray(dests, np.uint32).ravel())
        pr = np.ascontiguousarray(np.asarray(probs, np.float64).ravel())
        par = Params(goal_bias, int(pseudo), int(max_resident), frontier_chunks, arena_entries, log_entries, 0,
                     float(tie_band))
        self.h = C.c_void_p()
        _check(L.imomd_planner_create(dg.h, self.Q, self.K, d, pr, C.byref(par), C.byref(self.h)))

    def close(self):
        if self.h:
            cuda_lib().imomd_planner_destroy(self.h)
            self.h = None

    def reset(self):
        _check(cuda_lib().imomd_planner_reset(self.h))

    def set_queries(self, dests_u32, probs_f64):
        """dests_u32: contiguous u32[Q*K]; probs_f64: contiguous f64[Q*K*K] (host buffers)."""
        _check(cuda_lib().imomd_planner_set_queries(self.h, dests_u32, probs_f64))

    def tie_info(self, q=0):
        out = np.zeros(8)
        _check(cuda_lib().imomd_get_tie_info(self.h, q, out))
        return out

    def counters(self):
        out = np.zeros((self.Q, 8), np.uint64)
        _check(cuda_lib().imomd_get_counters(self.h, out))
        return out

    def clear_words(self):
        _check(cuda_lib().imomd_planner_clear_words(self.h))

    def feed(self, words):
        w = np.ascontiguousarray(words, np.uint32)
        _check(cuda_lib().imomd_planner_feed_words(self.h, w, w.size))

    def expand(self, iters):
        rep = (Report * self.Q)()
        _check(cuda_lib().imomd_expand(self.h, iters, rep))
        return list(rep)

    def expand_batch(self, steps, B):
        """batched mode: `steps` steps of B expansions per tree, trees expanded concurrently"""
        rep = (Report * self.Q)()
        _check(cuda_lib().imomd_expand_batch(self.h, steps, B, rep))
        return list(rep)

    def set_time_budget(self, seconds):
        _check(cuda_lib().imomd_set_time_budget(self.h, seconds))

    def expand_async(self, iters):
        _check(cuda_lib().imomd_expand_async(self.h, iters))

    def sync(self):
        rep = (Report * self.Q)()
        _check(cuda_lib().imomd_sync(self.h, rep))
        return list(rep)

    def last_kernel_ms(self):
        ms = C.c_float()
        _check(cuda_lib().imomd_last_kernel_ms(self.h, C.byref(ms)))
        return ms.value

    def set_profiling(self, on=True):
        """Per-request cycle counters of profile() / phase_cycles(): off unless switched on."""
        _check(cuda_lib().imomd_set_profiling(self.h, 1 if on else 0))

    def profile(self, q=0):
        cyc = np.zeros(8, np.uint64); cnt = np.zeros(8, np.uint64)
        _check(cuda_lib().imomd_get_profile(self.h, q, cyc, cnt))
        names = ["-", "nn", "rescan", "mhins", "conn", "bfs", "check", "serial"]
        d = {names[i]: (int(cyc[i]), int(cnt[i])) for i in range(1, 8)}
        d["bfs_vertices_levels"] = (int(cyc[0]), int(cnt[0]))
        return d

    def phase_cycles(self, q=0):
        cyc = np.zeros(8, np.uint64); cnt = np.zeros(8, np.uint64)
        _check(cuda_lib().imomd_get_phase_cycles(self.h, q, cyc, cnt))
        return cyc, cnt

    def tree(self, k, q=0):
        par = np.empty(self.g.n, np.uint32); cost = np.empty(self.g.n)
        _check(cuda_lib().imomd_get_tree(self.h, q, k, par, cost))
        return par, cost

    def state(self, q=0):
        K = self.K
        D = np.empty(K * K); cn = np.empty(K * K, np.uint32); P = np.empty(K * K)
        mi = np.empty(K * K, np.uint32); mv = np.empty(K * K)
        done = np.empty(K, np.int32); ds = np.empty(K, np.int32)
        _check(cuda_lib().imomd_get_state(self.h, q, D, cn, P, mi, mv, done, ds))
        c, u = C.c_int32(), C.c_int32()
        _check(cuda_lib().imomd_get_flags(self.h, q, C.byref(c), C.byref(u)))
        return dict(D=D.reshape(K, K), conn=cn.reshape(K, K), P=P.reshape(K, K),
                    mh_id=mi.reshape(K, K), mh_val=mv.reshape(K, K), is_done=done, ds_parent=ds,
                    connected=c.value, dm_updated=u.value)

    def states(self, first=0, count=None):
        """K x K distance / connection-node / probability matrices and is_done flags of `count`
        consecutive queries in one device->host copy."""
        K = self.K
        count = self.Q - first if count is None else count
        D = np.empty((count, K, K)); cn = np.empty((count, K, K), np.uint32); P = np.empty((count, K, K))
        done = np.empty((count, K), np.int32)
        _check(cuda_lib().imomd_get_states(self.h, first, count, D.ctypes.data, cn.ctypes.data, P.ctypes.data,
                                           done.ctypes.data))
        return dict(D=D, conn=cn, P=P, is_done=done)

    def treehash(self, q=0):
        """(TREEHASH of SURVEY App. A, order-free checksum) of query q's trees, computed on the device."""
        h, sm = C.c_uint64(), C.c_uint64()
        _check(cuda_lib().imomd_get_treehash(self.h, q, C.byref(h), C.byref(sm)))
        return h.value, sm.value

    def info(self):
        nq, sl, st, cr = C.c_int32(), C.c_int32(), C.c_int32(), C.c_uint64()
        _check(cuda_lib().imomd_planner_info(self.h, C.byref(nq), C.byref(sl), C.byref(st), C.byref(cr)))
        return dict(queries=nq.value, slots=sl.value, streaming=bool(st.value), certified_requests=cr.value)

    def frontier(self, k, q=0):
        out = np.empty(self.g.n, np.uint32); c = C.c_uint32()
        _check(cuda_lib().imomd_get_frontier(self.h, q, k, out, self.g.n, C.byref(c)))
        return np.sort(out[:c.value]).astype(np.uint64)

    def visit_order(self, k, q=0):
        """pseudo-mode planners: the tree's vertices in the order they were visited (root first)."""
        out = np.empty(self.g.n, np.uint32); c = C.c_uint32()
        _check(cuda_lib().imomd_get_visit_order(self.h, q, k, out, self.g.n, C.byref(c)))
        return out[:c.value].copy()

    def walk_to_root(self, k, vertex, q=0, cap=1 << 20):
        out = np.empty(cap, np.uint32); ln = C.c_uint32()
        _check(cuda_lib().imomd_walk_to_root(self.h, q, k, vertex, out, cap, C.byref(ln)))
        return out[:min(ln.value, cap)].copy()

    def connection_log(self, q=0, cap=1 << 20):
        si = np.empty(cap, np.uint32); vx = np.empty(cap, np.uint32); c = C.c_uint32()
        _check(cuda_lib().imomd_get_connection_log(self.h, q, si, vx, cap, C.byref(c)))
        n = min(c.value, cap)
        return si[:n].copy(), vx[:n].copy()

    def nearest(self, lookups):
        """lookups: structured array NN_DTYPE -> (idx u32[n], dist f64[n])."""
        lk = np.ascontiguousarray(lookups, NN_DTYPE)
        n = lk.shape[0]
        idx = np.empty(n, np.uint32); dist = np.empty(n)
        _check(cuda_lib().imomd_nearest(self.h, n, lk.ctypes.data_as(C.c_void_p), idx, dist))
        return idx, dist


def nearest_batch(dp, lookups):
    """imomd_nearest_batch: (idx, dist, dict(kernel_ms, bytes, refined, prep_ms, merge_ms, segments, call_ms))."""
    lk = np.ascontiguousarray(lookups, NN_DTYPE)
    n = lk.shape[0]
    idx = np.empty(n, np.uint32); dist = np.empty(n)
    ms, by, nr = C.c_float(), C.c_uint64(), C.c_uint32()
    t0 = time.perf_counter()
    _check(cuda_lib().imomd_nearest_batch(dp.h, n, lk.ctypes.data_as(C.c_void_p), idx, dist,
                                          C.byref(ms), C.byref(by), C.byref(nr)))
    call_ms = (time.perf_counter() - t0) * 1e3          # host buffers in, host buffers out
    prep, tiles, merge, nseg = C.c_float(), C.c_float(), C.c_float(), C.c_uint32()
    _check(cuda_lib().imomd_nearest_batch_last_timing(dp.h, C.byref(prep), C.byref(tiles), C.byref(merge),
                                                      C.byref(nseg)))
    return idx, dist, dict(kernel_ms=ms.value, bytes=by.value, refined=nr.value, prep_ms=prep.value,
                           merge_ms=merge.value, segments=nseg.value, call_ms=call_ms)


class GaContext:
    def __init__(self, device=0):
        self.h = C.c_void_p()
        _check(cuda_lib().imomd_ctx_create(device, C.byref(self.h)))

    def close(self):
        if self.h:
            cuda_lib().imomd_ctx_destroy(self.h)
            self.h = None

    def eval(self, D, tours, lens):
        D = np.ascontiguousarray(D, np.float64)
        K = D.shape[0]
        tours = np.ascontiguousarray(tours, np.int32)
        lens = np.ascontiguousarray(lens, np.int32)
        out = np.empty(tours.shape[0])
        _check(cuda_lib().imomd_ga_eval(self.h, K, D.ravel(), tours.shape[0], tours.shape[1],
                                        tours.ravel(), lens, out))
        return out


class HostRow(C.Structure):
    _fields_ = [("time_s", C.c_double), ("path_cost", C.c_double), ("tree_size", C.c_longlong),
                ("iteration", C.c_longlong)]


def facade_run(g, dest, goal_bias=1.0, max_iter=1000000, max_time=60, ga=(10000, 1000, 5), device=0,
               pseudo=False):
    """One full planner run through the C++ facade (ImomdRRT + findShortestPath on a pthread)."""
    L = host_lib()
    eu, ev, ew = g.edge_list()
    eu = np.ascontiguousarray(eu, np.uint64); ev = np.ascontiguousarray(ev, np.uint64)
    ew = np.ascontiguousarray(ew, np.float64)
    rows = (HostRow * 4096)()
    n_rows = C.c_int(); path = np.empty(1 << 22, np.uint64); path_len = C.c_uint64()
    seq = np.empty(256, np.int32); seq_len = C.c_int(); cost = C.c_double()
    its = C.c_longlong(); exps = C.c_longlong(); err = C.create_string_buffer(512)
    rc = L.imomd_host_run(device, g.n, g.lat, g.lon, len(eu), eu, ev, ew, len(dest),
                          np.asarray(dest, np.uint64), goal_bias, int(pseudo), max_iter, max_time, ga[0], ga[1], ga[2],
                          C.cast(rows, C.c_void_p), 4096, C.byref(n_rows), path, path.size,
                          C.byref(path_len), seq, 256, C.byref(seq_len), C.byref(cost), C.byref(its),
                          C.byref(exps), err, 512)
    if rc != 0:
        raise ImomdError(err.value.decode())
    return dict(rows=[(r.time_s, r.path_cost, r.tree_size, r.iteration) for r in rows[:n_rows.value]],
                path=path[:path_len.value].copy(), sequence=seq[:seq_len.value].copy(),
                final_cost=cost.value, iterations=its.value, expansions=exps.value)


def graph_cache_roundtrip(g, path):
    """GraphCache::save + load of the containers built from g's edge list (C++ facade library); returns
    (row_ptr, col, w, lat, lon) of the LOADED containers in their iteration order."""
    L = host_lib()
    eu, ev, ew = g.edge_list()
    eu = np.ascontiguousarray(eu, np.uint64); ev = np.ascontiguousarray(ev, np.uint64)
    ew = np.ascontiguousarray(ew, np.float64)
    row_ptr = np.empty(g.n + 1, np.uint32); col = np.empty(2 * len(eu), np.uint32); w = np.empty(2 * len(eu))
    lat = np.empty(g.n); lon = np.empty(g.n)
    err = C.create_string_buffer(512)
    rc = L.imomd_host_cache_roundtrip(path.encode(), g.n, g.lat, g.lon, len(eu), eu, ev, ew, row_ptr, col, w,
                                      lat, lon, err, 512)
    if rc != 0:
        raise ImomdError(err.value.decode())
    return row_ptr, col[:row_ptr[-1]], w[:row_ptr[-1]], lat, lon


class FacadePlanner:
    """The C++ facade driven stepwise: `passes` x expandTreeLayers_, then optionally one
    solveRTSP_ (which, in pseudo mode, performs the tree merge).  `.dev` is a DevicePlanner view
    of the facade's C-ABI planner for state readback."""

    def __init__(self, g, dest, goal_bias=1.0, pseudo=False, ga=(10000, 1000, 5), device=0):
        L = host_lib()
        eu, ev, ew = g.edge_list()
        eu = np.ascontiguousarray(eu, np.uint64); ev = np.ascontiguousarray(ev, np.uint64)
        ew = np.ascontiguousarray(ew, np.float64)
        err = C.create_string_buffer(512)
        self.h = L.imomd_host_planner_create(device, g.n, g.lat, g.lon, len(eu), eu, ev, ew, len(dest),
                                             np.asarray(dest, np.uint64), goal_bias, int(pseudo),
                                             ga[0], ga[1], ga[2], err, 512)
        if not self.h:
            raise ImomdError(err.value.decode())
        dev = DevicePlanner.__new__(DevicePlanner)
        dev.h = C.c_void_p(L.imomd_host_planner_native(self.h))
        dev.g, dev.dg, dev.Q, dev.K = g, None, 1, len(dest)
        self.dev = dev

    def step(self, passes, solve_rtsp=True):
        err = C.create_string_buffer(512)
        if host_lib().imomd_host_planner_step(self.h, passes, int(solve_rtsp), err, 512) != 0:
            raise ImomdError(err.value.decode())

    def cost(self):
        return host_lib().imomd_host_planner_cost(self.h)

    def close(self):
        if self.h:
            host_lib().imomd_host_planner_destroy(self.h)
            self.h = None
            self.dev.h = None


class FacadeSolver:
    """EciGenSolver of the host facade (GA fitness evaluated on the GPU)."""

    def __init__(self, device=0, swapping=True, genetic=True, ga=(10000, 1000, 5)):
        self.h = host_lib().imomd_host_solver_create(device, int(swapping), int(genetic), *ga)

    def solve(self, D, source, target):
        D = np.ascontiguousarray(D, np.float64)
        K = D.shape[0]
        cost = C.c_double(); seq = np.empty(256, np.int32); ev = C.c_uint64()
        err = C.create_string_buffer(512)
        n = host_lib().imomd_host_solver_solve(self.h, K, D.ravel(), source, target, C.byref(cost), seq,
                                               256, C.byref(ev), err, 512)
        if n < 0:
            raise ImomdError(err.value.decode())
        return cost.value, seq[:n].copy(), ev.value

    def island_begin(self, D, source, target, seed):
        D = np.ascontiguousarray(D, np.float64)
        err = C.create_string_buffer(512)
        if host_lib().imomd_host_island_begin(self.h, D.shape[0], D.ravel(), source, target, seed, err, 512):
            raise ImomdError(err.value.decode())

    def island_step(self):
        err = C.create_string_buffer(512)
        rc = host_lib().imomd_host_island_step(self.h, err, 512)
        if rc < 0:
            raise ImomdError(err.value.decode())
        return rc == 1

    def island_best(self):
        cost = C.c_double(); seq = np.empty(256, np.int32)
        n = host_lib().imomd_host_island_best(self.h, C.byref(cost), seq, 256)
        return cost.value, [int(v) for v in seq[:n]]

    def island_accept(self, cost, tour):
        t = np.ascontiguousarray(tour, np.int32)
        host_lib().imomd_host_island_accept(self.h, cost, t, len(t))

    def island_evaluated(self):
        return int(host_lib().imomd_host_island_evaluated(self.h))

    def close(self):
        if self.h:
            host_lib().imomd_host_solver_destroy(self.h)
            self.h = None
