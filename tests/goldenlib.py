"""Loading the committed golden vectors (tests/golden/, generated from the unmodified
reference by tests/golden/make_golden.py)."""
import json
import os

import numpy as np

from imomd_b200.graphs import RoadGraph

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def graph(name):
    return RoadGraph.load(os.path.join(GOLDEN, "graphs", name + ".npz"))


def expansion_cases():
    return json.load(open(os.path.join(GOLDEN, "expansion.json")))["cases"]


def runs():
    return json.load(open(os.path.join(GOLDEN, "runs.json")))["runs"]


def runs_pseudo():
    return json.load(open(os.path.join(GOLDEN, "runs_pseudo.json")))["runs"]


def rtsp():
    return json.load(open(os.path.join(GOLDEN, "rtsp.json")))["solvers"]


def f64(int_bits):
    return np.asarray(int_bits, np.uint64).view(np.float64)


def treehash(trees):
    """FNV-style hash of SURVEY.md Appendix A step 6 over [(parent u32[N], cost f64[N]), ...]."""
    h = 1469598103934665603
    M = (1 << 64) - 1
    for par, cost in trees:
        vis = np.nonzero(par != 0xFFFFFFFF)[0]
        cb = np.ascontiguousarray(cost).view(np.uint64)
        for v in vis:
            h = ((h ^ int(v)) * 1099511628211) & M
            h = ((h ^ int(par[v])) * 1099511628211) & M
            h = ((h ^ int(cb[v])) * 1099511628211) & M
    return h


def check_checkpoint(impl_state, trees, cp, K, exact_heuristics):
    """impl_state: dict with D, conn, P, mh_id, mh_val, is_done, ds_parent, connected, dm_updated."""
    assert f"{treehash(trees):016x}" == cp["treehash"]
    assert [int((p != 0xFFFFFFFF).sum()) for p, _ in trees] == cp["tree_sizes"]
    D = np.asarray(impl_state["D"]).ravel()
    assert np.array_equal(D.view(np.uint64), np.asarray(cp["D"], np.uint64))
    conn = np.asarray(impl_state["conn"]).ravel().astype(np.uint64)
    conn[conn == 0xFFFFFFFF] = np.uint64(0xFFFFFFFFFFFFFFFF)
    assert np.array_equal(conn, np.asarray(cp["conn"], np.uint64))
    P = np.asarray(impl_state["P"]).ravel()
    assert np.array_equal(P.view(np.uint64), np.asarray(cp["P"], np.uint64))
    mi = np.asarray(impl_state["mh_id"]).ravel().astype(np.uint64)
    mi[mi == 0xFFFFFFFF] = np.uint64(0xFFFFFFFFFFFFFFFF)
    assert np.array_equal(mi, np.asarray(cp["mh_id"], np.uint64))
    mv, want = np.asarray(impl_state["mh_val"]).ravel(), f64(cp["mh_val"])
    if exact_heuristics:
        assert np.array_equal(mv.view(np.uint64), want.view(np.uint64))
    else:
        fin = np.isfinite(want)
        assert np.array_equal(np.isfinite(mv), fin)
        assert np.allclose(mv[fin], want[fin], rtol=1e-12, atol=0)
    assert list(impl_state["is_done"]) == cp["is_done"]
    assert list(impl_state["ds_parent"]) == cp["ds_parent"]
    assert impl_state["connected"] == cp["connected"] and impl_state["dm_updated"] == cp["dm_updated"]
