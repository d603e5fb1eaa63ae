"""GPU tests of the batched expansion mode (imomd_expand_batch): the K trees of a query are
expanded concurrently, one warp per tree, B expansions per tree per step, cross-tree effects
applied after the step.  The oracle restates exactly these semantics on the CPU
(orc_planner_iterate_batched), so parents, cost bits, frontiers, matrices and flags are compared
bit for bit, like in the exact mode; B = 1 per step is still NOT the reference's order (the
other trees are looked at after the pass), the exact mode stays imomd_expand."""
import numpy as np
import pytest

import oraclelib as ol
from imomd_b200 import capi, graphs
from parity import compare_with_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,K,gb,B", [("grid60", 6, 1.0, 1), ("grid60", 6, 1.0, 4), ("grid90", 12, 1.0, 8),
                                         ("road", 5, 0.7, 3), ("road", 9, 1.0, 16), ("grid150", 12, 0.5, 8),
                                         ("hub", 5, 1.0, 4)])
def test_batched_mode_matches_its_oracle(kind, K, gb, B):
    g = {"grid60": lambda: graphs.jittered_grid(60, seed=11), "grid90": lambda: graphs.jittered_grid(90, seed=3),
         "grid150": lambda: graphs.jittered_grid(150, seed=8), "road": lambda: graphs.road_like(6000, seed=2),
         "hub": lambda: graphs.hub_grid(40, seed=4, hubs=14, spokes=9)}[kind]()
    dest = graphs.pick_destinations(g, K, seed=5)
    op = ol.OraclePlanner(g, dest, gb)
    dg = capi.DeviceGraph(g)
    dp = capi.DevicePlanner(dg, [dest], None, gb)
    dp.feed(capi.mt19937_words(0, 1_000_000))
    for steps in (1, 5, 60, 200):
        rep = dp.expand_batch(steps, B)[0]
        e = op.iterate_batched(steps, B)
        assert rep.status == 0 and rep.iterations == steps * B and rep.expansions == e
        compare_with_oracle(dp, op, K, False)
        assert dp.treehash(0)[0] == op.treehash()
    assert rep.unresolved_ties == 0
    dp.close(); dg.close(); op.close()


def test_batched_mode_streaming_many_queries_and_determinism():
    """streaming planner + batched kernel; a second run gives identical fingerprints"""
    g = graphs.jittered_grid(100, seed=21)
    K, Q = 7, 24
    dests = [graphs.pick_destinations(g, K, seed=500 + q) for q in range(Q)]
    dg = capi.DeviceGraph(g)
    dp = capi.DevicePlanner(dg, dests, max_resident=5)
    dp.feed(capi.mt19937_words(0, 600000))
    first = dp.expand_batch(100, 8)
    second = dp.expand_batch(100, 8)
    assert [r.tree_hash for r in first] == [r.tree_hash for r in second]
    for q in (0, 7, 23):
        o = ol.OraclePlanner(g, dests[q])
        assert first[q].expansions == o.iterate_batched(100, 8)
        assert first[q].tree_hash == o.treehash() and first[q].status == 0
        st = dp.states(q, 1)
        assert np.array_equal(st["D"][0].view(np.uint64), o.matrices()["D"].view(np.uint64))
        o.close()
    dp.close(); dg.close()


def test_batched_mode_word_starvation_and_rejections():
    g = graphs.jittered_grid(50, seed=5)
    dest = graphs.pick_destinations(g, 4, seed=2)
    dg = capi.DeviceGraph(g)
    dp = capi.DevicePlanner(dg, [dest])
    dp.feed(capi.mt19937_words(0, 4 * 8 * 40 + 64))      # enough for 40 passes
    rep = dp.expand_batch(20, 4)[0]                       # asks for 80
    assert rep.status == 1 and rep.iterations == 40
    with pytest.raises(capi.ImomdError):
        dp.expand_batch(1, 0)
    dpp = capi.DevicePlanner(dg, [dest], pseudo=True)
    with pytest.raises(capi.ImomdError, match="pseudo"):
        dpp.expand_batch(1, 1)
    dpp.close(); dp.close(); dg.close()


def test_batched_mode_path_quality_vs_exact_mode():
    """same number of expansions per tree: the pseudo-complete destination graph the batched mode
    builds is as good as the exact mode's (connected, tour cost within a few percent)"""
    g = graphs.jittered_grid(200, seed=13)
    K = 8
    dest = graphs.pick_destinations(g, K, seed=9)
    dg = capi.DeviceGraph(g)
    exact = capi.DevicePlanner(dg, [dest])
    exact.feed(capi.mt19937_words(0, 2_000_000))
    exact.expand(4000)
    batched = capi.DevicePlanner(dg, [dest])
    batched.feed(capi.mt19937_words(0, 2_000_000))
    batched.expand_batch(500, 8)
    se, sb = exact.state(0), batched.state(0)
    assert se["connected"] and sb["connected"]
    fs = capi.FacadeSolver(0)
    ce, _, _ = fs.solve(se["D"], 0, K - 1)
    fs.close()
    fs = capi.FacadeSolver(0)
    cb, _, _ = fs.solve(sb["D"], 0, K - 1)
    fs.close()
    assert cb <= ce * 1.03, (cb, ce)
    exact.close(); batched.close(); dg.close()
