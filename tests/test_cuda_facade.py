"""The C++ host facade (reference library API: ImomdRRT constructor + findShortestPath pthread,
EciGenSolver::solveRTSP) on top of the CUDA path, against the unmodified reference run the
same way (oracle/_ref, travels prebuilt to the GPU box)."""
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest

import oraclelib as ol
from imomd_b200 import capi, graphs

pytestmark = pytest.mark.gpu


def _random_pseudo_complete(K, rng, p_missing):
    D = rng.uniform(100.0, 5000.0, (K, K)); D = (D + D.T) / 2
    miss = np.triu(rng.random((K, K)) < p_missing, 1)
    D[miss | miss.T] = np.inf
    perm = rng.permutation(K)
    for a, b in zip(perm[:-1], perm[1:]):
        if np.isinf(D[a, b]):
            D[a, b] = D[b, a] = rng.uniform(100.0, 5000.0)
    np.fill_diagonal(D, 0.0)
    return np.ascontiguousarray(D)


@pytest.mark.parametrize("K,p_missing", [(4, 0.0), (7, 0.3), (12, 0.5), (22, 0.6), (12, 0.0)])
def test_eci_gen_solver_with_gpu_fitness(K, p_missing):
    """Same (cost, tour) as the oracle restatement over a sequence of solves (the solver-owned
    RNG persists), with every GA fitness evaluated by the device kernel."""
    O = ol.oracle()
    os_ = O.orc_solver_create(1, 1, 2000, 300, 5)
    fs = capi.FacadeSolver(0, ga=(2000, 300, 5))
    rng = np.random.default_rng(K * 100 + int(p_missing * 10))
    for trial in range(6):
        D = _random_pseudo_complete(K, rng, p_missing)
        c2 = ol.C.c_double(); s2 = np.empty(256, np.int32)
        n2 = O.orc_solver_solve(os_, K, D.ravel(), 0, K - 1, ol.C.byref(c2), s2, 256)
        c1, s1, evaluated = fs.solve(D, 0, K - 1)
        assert np.array_equal(s1, s2[:n2]), (trial, s1, s2[:n2])
        assert c1 == c2.value
        assert evaluated > 0
    fs.close(); O.orc_solver_destroy(os_)


def _probe_run(g, dest, max_iter, ga, pseudo=False, goal_bias=1.0):
    with tempfile.TemporaryDirectory() as td:
        gp = os.path.join(td, "g.bin")
        g.write_probe_file(gp)
        out = subprocess.check_output(
            [ol.REF_PROBE, "--graph", gp, "--dest", ",".join(map(str, dest)), "--mode", "run",
             "--max-iter", str(max_iter), "--max-time", "600", "--ga", ",".join(map(str, ga))]
            + (["--pseudo"] if pseudo else []) + ["--goal-bias", repr(goal_bias)], text=True)
    return json.loads(out.strip().splitlines()[-1])


@pytest.mark.ref
@pytest.mark.parametrize("kind,K,max_iter", [("grid", 5, 1500), ("road", 4, 2500), ("grid", 8, 2500)])
def test_anytime_run_matches_reference(kind, K, max_iter):
    """Whole planner run: every logged improvement (iteration, path cost, tree size) and the
    final vertex path equal the reference's."""
    g = graphs.jittered_grid(70, seed=21) if kind == "grid" else graphs.road_like(5000, seed=4)
    dest = graphs.pick_destinations(g, K, seed=K + 1)
    ga = (3000, 400, 5)
    ref = _probe_run(g, dest, max_iter, ga)
    got = capi.facade_run(g, dest, max_iter=max_iter, max_time=600, ga=ga)
    assert got["iterations"] == ref["iterations"]
    assert got["expansions"] == ref["expansions"]
    assert len(ref["rows"]) > 0, "the reference found no path: pick another scenario"
    assert [(r[3], r[2]) for r in got["rows"]] == [(r[3], r[2]) for r in ref["rows"]]
    for a, b in zip(got["rows"], ref["rows"]):
        assert a[1] == b[1], (a, b)              # path cost, bit for bit
    assert got["final_cost"] == ref["final_cost"]
    assert list(got["path"]) == ref["path"]


@pytest.mark.ref
@pytest.mark.parametrize("kind,K,max_iter", [("grid", 5, 1500), ("road", 4, 2500), ("grid", 8, 2500),
                                             ("grid", 3, 1200)])
def test_pseudo_mode_run_matches_reference(kind, K, max_iter):
    """pseudo_mode: the trees of the pseudo destinations are merged into the source / target
    trees once everything is connected (mergePseudoTrees_, :744-916 upstream) and the planner
    keeps improving the merged pair.  Rows, final cost and path equal the reference's."""
    g = graphs.jittered_grid(70, seed=21) if kind == "grid" else graphs.road_like(5000, seed=4)
    dest = graphs.pick_destinations(g, K, seed=K + 1)
    ga = (3000, 400, 5)
    ref = _probe_run(g, dest, max_iter, ga, pseudo=True)
    got = capi.facade_run(g, dest, max_iter=max_iter, max_time=600, ga=ga, pseudo=True)
    assert len(ref["rows"]) > 0, "the reference found no path: pick another scenario"
    assert got["iterations"] == ref["iterations"]
    assert got["expansions"] == ref["expansions"]
    assert [(r[3], r[2]) for r in got["rows"]] == [(r[3], r[2]) for r in ref["rows"]]
    for a, b in zip(got["rows"], ref["rows"]):
        assert a[1] == b[1], (a, b)
    assert got["final_cost"] == ref["final_cost"]
    assert list(got["path"]) == ref["path"]


@pytest.mark.ref
@pytest.mark.parametrize("width,goal_bias,max_iter", [(60, 0.3, 3000), (90, 0.3, 4000), (60, 1.0, 3000)])
def test_bugtrap_pseudo_mode_matches_reference(width, goal_bias, max_iter):
    """BASELINE configs[1] (stand-in for the missing sanf.osm): source inside a
    U-shaped trap, target behind its closed side, 5 pseudo destinations around the opening,
    pseudo mode, goal_bias 0.3 as in the README recipe.  Rows, cost and path equal the
    reference's (run in its own process: goal_bias < 1 uses function-local static state)."""
    g, source, target, pseudo = graphs.bugtrap_grid(width, seed=3)
    dest = [source] + pseudo + [target]
    ga = (3000, 400, 5)
    ref = _probe_run(g, dest, max_iter, ga, pseudo=True, goal_bias=goal_bias)
    got = capi.facade_run(g, dest, goal_bias=goal_bias, max_iter=max_iter, max_time=600, ga=ga,
                          pseudo=True)
    assert len(ref["rows"]) > 0, "the reference found no path: pick another scenario"
    assert got["iterations"] == ref["iterations"]
    assert got["expansions"] == ref["expansions"]
    assert [(r[3], r[2]) for r in got["rows"]] == [(r[3], r[2]) for r in ref["rows"]]
    for a, b in zip(got["rows"], ref["rows"]):
        assert a[1] == b[1], (a, b)
    assert got["final_cost"] == ref["final_cost"]
    assert list(got["path"]) == ref["path"]


@pytest.mark.ref
@pytest.mark.parametrize("kind,K", [("grid", 5), ("road", 4), ("grid", 8)])
def test_pseudo_merge_state_matches_reference(kind, K):
    """Stepwise: 100 passes + solveRTSP_ at a time.  After the merge (the first solveRTSP_ on a
    connected graph) and after every later step, every tree (parents, cost bits, frontier), the
    matrices and the flags equal the reference's."""
    from parity import compare_with_oracle
    R = ol.ref()
    g = graphs.jittered_grid(70, seed=21) if kind == "grid" else graphs.road_like(5000, seed=4)
    dest = graphs.pick_destinations(g, K, seed=K + 1)
    gh = R.ref_graph_from_edges(g.n, g.lat, g.lon, len(g.eu), g.eu, g.ev,
                                g.ew.ctypes.data_as(ol.C.c_void_p))
    rp = ol.RefPlanner(gh, dest, pseudo=True, ga=(3000, 400, 5))
    fp = capi.FacadePlanner(g, dest, pseudo=True, ga=(3000, 400, 5))
    merged = False
    for step in range(8):
        rp.iterate(100); rp.solve_rtsp()
        fp.step(100, True)
        compare_with_oracle(fp.dev, rp, K, exact_heuristics=False)
        merged = merged or bool(rp.flags()["connected"])
        assert fp.cost() == rp.path_cost() or not merged
    assert merged, "the trees never connected: pick another scenario"
    fp.close()


@pytest.mark.parametrize("K", [4, 13, 14, 30])
def test_eci_gen_solver_exact_ties(K):
    O = ol.oracle()
    os_ = O.orc_solver_create(1, 1, 500, 100, 3)
    fs = capi.FacadeSolver(0, ga=(500, 100, 3))
    rng = np.random.default_rng(K)
    for trial in range(6):
        D = rng.integers(1, 4, (K, K)).astype(np.float64) * 100.0
        D = np.ascontiguousarray((D + D.T) / 2)
        np.fill_diagonal(D, 0.0)
        c2 = ol.C.c_double(); s2 = np.empty(512, np.int32)
        n2 = O.orc_solver_solve(os_, K, D.ravel(), 0, K - 1, ol.C.byref(c2), s2, 512)
        c1, s1, _ = fs.solve(D, 0, K - 1)
        assert np.array_equal(s1, s2[:n2]) and c1 == c2.value
    fs.close(); O.orc_solver_destroy(os_)


def test_ga_island_single_rank_matches_plain_solve():
    """One island with seed 0 and no migration partner is the reference's GA: same tour."""
    from imomd_b200 import sharding
    rng = np.random.default_rng(5)
    K = 14
    D = _random_pseudo_complete(K, rng, 0.3)
    fs = capi.FacadeSolver(0, ga=(2000, 300, 5))
    want_cost, want_seq, _ = fs.solve(D, 0, K - 1)
    fs.close()
    cost, tour, hist, evaluated = sharding.run_islands(D, 0, K - 1, 5, ga=(2000, 300, 5))
    assert cost == want_cost and tour == [int(v) for v in want_seq]
    assert evaluated > 0 and [h["best"] for h in hist] == sorted([h["best"] for h in hist], reverse=True)
