"""Pins the CPU restatement (oracle/imomd_oracle.cpp) against the UNMODIFIED reference
compiled into oracle/_ref (the reference ships no golden vectors of its own, SURVEY.md
section 4).  Runs only where oracle/_ref exists; the fixtures it implies are frozen in
tests/golden/ by tests/golden/make_golden.py for the GPU box.
"""
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest

import oraclelib as ol
from imomd_b200 import graphs

pytestmark = [pytest.mark.ref]

OSM = "/root/reference/osm_data/%s.osm"


def _full_compare(rp, op, K, n, check_children_of=()):
    for k in range(K):
        pr, cr = rp.tree(k)
        po, co = op.tree(k)
        assert np.array_equal(pr, po), f"tree {k}: parents differ at {np.nonzero(pr != po)[0][:5]}"
        assert np.array_equal(cr.view(np.uint64), co.view(np.uint64)), f"tree {k}: cost bits differ"
        assert np.array_equal(rp.frontier(k), op.frontier(k)), f"tree {k}: frontier differs"
    mr, mo = rp.matrices(), op.matrices()
    for key in ("D", "conn", "P", "mh_id", "mh_val"):
        a, b = mr[key], mo[key]
        if a.dtype == np.float64:
            assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), f"matrix {key} differs"
        else:
            assert np.array_equal(a, b), f"matrix {key} differs"
    fr, fo = rp.flags(), op.flags()
    assert np.array_equal(fr["is_done"], fo["is_done"])
    assert fr["connected"] == fo["connected"] and fr["dm_updated"] == fo["dm_updated"]
    assert np.array_equal(fr["ds_parent"], fo["ds_parent"])
    assert rp.treehash() == op.treehash()


@pytest.mark.refsrc
@pytest.mark.parametrize("name,K,iters", [("san_pablo", 7, 3000), ("FRB", 5, 1000),
                                          ("FRB-all", 10, 3000), ("FRB2", 5, 1000),
                                          ("san_pablo", 2, 3000), ("san_pablo", 3, 2000)])
def test_osm_goal_bias_one(name, K, iters):
    R = ol.ref()
    gh = R.ref_graph_from_osm((OSM % name).encode())
    g = ol.ref_export_graph(gh, name)
    dest = np.empty(K, np.uint64)
    R.ref_pick_destinations(gh, 7, K, dest)
    dest = [int(x) for x in dest]
    rp, op = ol.RefPlanner(gh, dest), ol.OraclePlanner(g, dest)
    # state right after construction (includes the half-built-trees quirk)
    _full_compare(rp, op, K, g.n)
    done = 0
    for chunk in (1, 9, 90, iters - 100):
        e1, e2 = rp.iterate(chunk), op.iterate(chunk)
        assert e1 == e2
        done += chunk
        _full_compare(rp, op, K, g.n)
    assert op.ties() == 0
    op.close()


@pytest.mark.refsrc
def test_children_order_matches_container():
    """The 13-bucket rule reproduces unordered_set<size_t> iteration for every vertex."""
    R = ol.ref()
    gh = R.ref_graph_from_osm((OSM % "san_pablo").encode())
    g = ol.ref_export_graph(gh, "san_pablo")
    dest = np.empty(7, np.uint64)
    R.ref_pick_destinations(gh, 7, 7, dest)
    dest = [int(x) for x in dest]
    rp, op = ol.RefPlanner(gh, dest), ol.OraclePlanner(g, dest)
    rp.iterate(1500), op.iterate(1500)
    O = ol.oracle()
    out = np.empty(64, np.uint64)
    checked = 0
    for k in range(7):
        par, _ = op.tree(k)
        for v in np.nonzero(par != 0xFFFFFFFF)[0]:
            c = O.orc_planner_children(op.ph, k, int(v), out, 64)
            mine = [int(x) for x in out[:c]]
            theirs = rp.children(k, int(v)) or []
            assert mine == theirs, (k, v, mine, theirs)
            checked += 1
    assert checked > 10000


def test_traces_grid():
    """Per-expansion traces (x_rand, x_nearest, x_new, parent, cost) on a jittered grid."""
    R = ol.ref()
    gh = R.ref_graph_grid(120)
    g = ol.ref_export_graph(gh, "grid120")
    K = 6
    dest = np.empty(K, np.uint64)
    R.ref_pick_destinations(gh, 7, K, dest)
    dest = [int(x) for x in dest]
    rp, op = ol.RefPlanner(gh, dest), ol.OraclePlanner(g, dest)
    for it in range(400):
        for k in range(K):
            a, b = rp.expand_traced(k), op.expand_traced(k)
            ta, tb = a.astuple(), b.astuple()
            for x, y in zip(ta, tb):
                if isinstance(x, float):
                    assert (np.isnan(x) and np.isnan(y)) or x == y, (it, k, ta, tb)
                else:
                    assert x == y, (it, k, ta, tb)
    _full_compare(rp, op, K, g.n)
    op.close()


def test_native_csr_emulation_matches_containers():
    """graphs.native_csr_from_edges and the oracle's C version give the order the real
    unordered_map containers iterate in."""
    R, O = ol.ref(), ol.oracle()
    g = graphs.jittered_grid(40)
    gh = R.ref_graph_from_edges(g.n, g.lat, g.lon, len(g.eu), g.eu, g.ev,
                                g.ew.ctypes.data_as(ol.C.c_void_p))
    gr = ol.ref_export_graph(gh)
    assert np.array_equal(gr.row_ptr, g.row_ptr) and np.array_equal(gr.col, g.col)
    assert np.array_equal(gr.w.view(np.uint64), g.w.view(np.uint64))
    rp = np.empty(g.n + 1, np.uint32); col = np.empty(g.arcs, np.uint32); w = np.empty(g.arcs)
    O.orc_native_csr_from_edges(g.n, len(g.eu), g.eu, g.ev, g.ew, rp, col, w)
    assert np.array_equal(rp, g.row_ptr) and np.array_equal(col, g.col)
    # a road-like graph has irregular degrees and ids
    g2 = graphs.road_like(3000, seed=3)
    gh2 = R.ref_graph_from_edges(g2.n, g2.lat, g2.lon, len(g2.eu), g2.eu, g2.ev,
                                 g2.ew.ctypes.data_as(ol.C.c_void_p))
    gr2 = ol.ref_export_graph(gh2)
    assert np.array_equal(gr2.row_ptr, g2.row_ptr) and np.array_equal(gr2.col, g2.col)


def _probe(args):
    out = subprocess.check_output([ol.REF_PROBE] + args, text=True)
    return json.loads(out.strip().splitlines()[-1])


def _load_dump(path):
    with open(path, "rb") as f:
        n, K = np.frombuffer(f.read(16), np.uint64)
        n, K = int(n), int(K)
        trees = []
        for _ in range(K):
            par = np.frombuffer(f.read(4 * n), np.uint32)
            cost = np.frombuffer(f.read(8 * n), np.float64)
            trees.append((par, cost))
        D = np.frombuffer(f.read(8 * K * K), np.float64).reshape(K, K)
        cn = np.frombuffer(f.read(8 * K * K), np.uint64).reshape(K, K)
    return trees, D, cn


@pytest.mark.parametrize("gb", [0.5, 0.0])
def test_goal_bias_below_one_via_probe(gb):
    """Uniform sampling branch (Lemire uniform_int) -- separate process because the
    reference's distributions are function-local statics."""
    g = graphs.jittered_grid(80, seed=9)
    dest = graphs.pick_destinations(g, 5, seed=3)
    with tempfile.TemporaryDirectory() as td:
        gp, dp = os.path.join(td, "g.bin"), os.path.join(td, "d.bin")
        g.write_probe_file(gp)
        res = _probe(["--graph", gp, "--dest", ",".join(map(str, dest)), "--iters", "1500",
                      "--goal-bias", str(gb), "--dump", dp])
        trees, D, cn = _load_dump(dp)
    op = ol.OraclePlanner(g, dest, gb)
    e = op.iterate(1500)
    assert e == res["expansions"]
    assert f"{op.treehash():016x}" == res["treehash"]
    for k in range(5):
        po, co = op.tree(k)
        assert np.array_equal(po, trees[k][0])
        assert np.array_equal(co.view(np.uint64), trees[k][1].view(np.uint64))
    m = op.matrices()
    assert np.array_equal(m["D"].view(np.uint64), D.view(np.uint64))
    assert np.array_equal(m["conn"], cn)
    op.close()


def test_fake_maps():
    R = ol.ref()
    for mtype, dest in ((-1, [0, 3, 2]), (-2, [0, 3, 4, 6])):
        gh = R.ref_graph_fake(mtype)
        g = ol.ref_export_graph(gh, f"fake{mtype}")
        rp, op = ol.RefPlanner(gh, dest), ol.OraclePlanner(g, dest)
        for _ in range(30):
            assert rp.iterate(1) == op.iterate(1)
            _full_compare(rp, op, len(dest), g.n)
        op.close()


def test_pseudo_mode_connection_sets():
    R = ol.ref()
    gh = R.ref_graph_grid(60)
    g = ol.ref_export_graph(gh, "grid60")
    K = 5
    dest = np.empty(K, np.uint64)
    R.ref_pick_destinations(gh, 11, K, dest)
    dest = [int(x) for x in dest]
    rp, op = ol.RefPlanner(gh, dest, pseudo=True), ol.OraclePlanner(g, dest, pseudo=True)
    rp.iterate(600), op.iterate(600)
    _full_compare(rp, op, K, g.n)
    si, vx = op.connection_log()
    nsets = K * (K - 1) // 2
    for idx in range(nsets):
        mine = set(int(v) for v in vx[si == idx])
        assert mine == set(rp.connection_set(idx)), idx
    op.close()


def test_rng_and_geometry_pins():
    R, O = ol.ref(), ol.oracle()
    a = np.empty(5000, np.uint32); b = np.empty(5000, np.uint32)
    for seed in (0, 7, 123):
        R.ref_mt19937_words(seed, 5000, a); O.orc_mt19937_words(seed, 5000, b)
        assert np.array_equal(a, b)
    x = np.empty(4000); y = np.empty(4000)
    for lo, hi in ((0.0, 1.0), (0.0, 0.5), (0.0, 3.25e-4), (-2e-5, 2e-5)):
        R.ref_uniform_real(0, lo, hi, 4000, x); O.orc_uniform_real(0, lo, hi, 4000, y)
        assert np.array_equal(x.view(np.uint64), y.view(np.uint64))
    u = np.empty(4000, np.uint64); v = np.empty(4000, np.uint64)
    for lo, hi in ((0, 6808), (0, 999999), (0, 9_999_999), (0, 1), (0, 4), (5, 5), (0, 2**32 - 2)):
        R.ref_uniform_size_t(3, lo, hi, 4000, u); O.orc_uniform_size_t(3, lo, hi, 4000, v)
        assert np.array_equal(u, v), (lo, hi)
    p = np.empty(4000, np.int32); q = np.empty(4000, np.int32)
    for lo, hi in ((0, 1), (0, 7), (0, 4), (1, 23), (-5, 17), (1, 1)):
        R.ref_uniform_int(0, lo, hi, 4000, p); O.orc_uniform_int(0, lo, hi, 4000, q)
        assert np.array_equal(p, q), (lo, hi)
    rng = np.random.default_rng(1)
    pts = np.column_stack((rng.uniform(-80, 80, 3000), rng.uniform(-179, 179, 3000),
                           rng.uniform(-80, 80, 3000), rng.uniform(-179, 179, 3000)))
    near = pts.copy(); near[:, 2:] = near[:, :2] + rng.uniform(-1e-3, 1e-3, (3000, 2))
    for row in np.vstack((pts, near)):
        assert R.ref_haversine(*row) == O.orc_haversine(*row)
        assert R.ref_bearing(*row) == O.orc_bearing(*row)
    assert O.orc_haversine(37.0, -122.0, 37.0, -122.0) == 0.0


def _random_pseudo_complete(K, rng, p_missing):
    D = rng.uniform(100.0, 5000.0, (K, K))
    D = (D + D.T) / 2
    miss = np.triu(rng.random((K, K)) < p_missing, 1)
    D[miss | miss.T] = np.inf
    # keep a spanning chain so the matrix is connected (the guard at
    # src/imomd_rrt_star.cpp:971 is load-bearing)
    perm = rng.permutation(K)
    for a, b in zip(perm[:-1], perm[1:]):
        if np.isinf(D[a, b]):
            D[a, b] = D[b, a] = rng.uniform(100.0, 5000.0)
    np.fill_diagonal(D, 0.0)
    return np.ascontiguousarray(D)


@pytest.mark.parametrize("K,p_missing", [(4, 0.0), (7, 0.3), (12, 0.5), (22, 0.6), (12, 0.0)])
def test_eci_gen_solver(K, p_missing):
    """solveRTSP on random pseudo-complete matrices; the solver-owned RNG persists across
    solves, so several matrices go through the same solver objects."""
    R, O = ol.ref(), ol.oracle()
    rs = R.ref_solver_create(1, 1, 2000, 300, 5)
    os_ = O.orc_solver_create(1, 1, 2000, 300, 5)
    rng = np.random.default_rng(K * 100 + int(p_missing * 10))
    for trial in range(6):
        D = _random_pseudo_complete(K, rng, p_missing)
        c1 = ol.C.c_double(); c2 = ol.C.c_double()
        s1 = np.empty(256, np.int32); s2 = np.empty(256, np.int32)
        n1 = R.ref_solver_solve(rs, K, D.ravel(), 0, K - 1, ol.C.byref(c1), s1, 256)
        n2 = O.orc_solver_solve(os_, K, D.ravel(), 0, K - 1, ol.C.byref(c2), s2, 256)
        assert n1 == n2 and np.array_equal(s1[:n1], s2[:n2]), (trial, s1[:n1], s2[:n2])
        assert c1.value == c2.value
    O.orc_solver_destroy(os_)


def test_path_costs_batch():
    R, O = ol.ref(), ol.oracle()
    rng = np.random.default_rng(4)
    K = 22
    D = _random_pseudo_complete(K, rng, 0.2)
    n, stride = 500, 64
    lens = rng.integers(2, 2 * K, n).astype(np.int32)
    tours = rng.integers(0, K, (n, stride)).astype(np.int32)
    a = np.empty(n); b = np.empty(n)
    R.ref_path_costs(K, D.ravel(), n, stride, tours.ravel(), lens, a)
    O.orc_path_costs(K, D.ravel(), n, stride, tours.ravel(), lens, b)
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64))


@pytest.mark.parametrize("K", [4, 9, 13, 14, 20, 29, 30, 32])
def test_eci_gen_solver_exact_ties(K):
    """Integer-valued symmetric matrices are full of exact cost ties: the candidate order of
    the cheapest insertion (an unordered_set<int> upstream) becomes visible."""
    R, O = ol.ref(), ol.oracle()
    rs = R.ref_solver_create(1, 1, 500, 100, 3)
    os_ = O.orc_solver_create(1, 1, 500, 100, 3)
    rng = np.random.default_rng(K)
    for trial in range(8):
        D = rng.integers(1, 4, (K, K)).astype(np.float64) * 100.0
        D = np.ascontiguousarray((D + D.T) / 2)
        np.fill_diagonal(D, 0.0)
        c1 = ol.C.c_double(); c2 = ol.C.c_double()
        s1 = np.empty(512, np.int32); s2 = np.empty(512, np.int32)
        n1 = R.ref_solver_solve(rs, K, D.ravel(), 0, K - 1, ol.C.byref(c1), s1, 512)
        n2 = O.orc_solver_solve(os_, K, D.ravel(), 0, K - 1, ol.C.byref(c2), s2, 512)
        assert n1 == n2 and np.array_equal(s1[:n1], s2[:n2]), (trial, s1[:n1], s2[:n2])
        assert c1.value == c2.value
    O.orc_solver_destroy(os_)


@pytest.mark.parametrize("hubs,spokes", [(10, 3), (14, 9)])
def test_high_degree_graph(hubs, spokes):
    """Degree up to 13: the adjacency-order emulation and the child-order rule at their limit,
    against the real containers."""
    R = ol.ref()
    g = graphs.hub_grid(40, seed=4, hubs=hubs, spokes=spokes)
    gh = R.ref_graph_from_edges(g.n, g.lat, g.lon, len(g.eu), g.eu, g.ev, g.ew.ctypes.data_as(ol.C.c_void_p))
    gr = ol.ref_export_graph(gh)
    assert np.array_equal(gr.row_ptr, g.row_ptr) and np.array_equal(gr.col, g.col)
    dest = graphs.pick_destinations(g, 5, seed=2)
    rp, op = ol.RefPlanner(gh, dest), ol.OraclePlanner(g, dest)
    for chunk in (100, 700, 1200):
        assert rp.iterate(chunk) == op.iterate(chunk)
        _full_compare(rp, op, 5, g.n)
    op.close()
