"""GPU tests of the round-2 additions to the expansion path, through the C ABI:
streaming planners (state slots + device work queue), the device-side TREEHASH, the batched
state readback, the exact-value service (near-ties decided on the host's glibc values through
mapped-memory mailboxes) and the full-depth parity of the timed configuration."""
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest

import oraclelib as ol
from imomd_b200 import capi, graphs
from parity import compare_with_oracle

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PROBE = os.path.join(ROOT, "oracle", "_ref", "ref_probe")


def test_resident_treehash_and_batched_states():
    g = graphs.road_like(6000, seed=2)
    K, Q = 5, 3
    dests = [graphs.pick_destinations(g, K, seed=40 + q) for q in range(Q)]
    dg = capi.DeviceGraph(g)
    dp = capi.DevicePlanner(dg, dests)
    assert dp.info() == dict(queries=Q, slots=Q, streaming=False, certified_requests=0)
    dp.feed(capi.mt19937_words(0, 200000))
    ops = [ol.OraclePlanner(g, d) for d in dests]
    for chunk in (1, 300, 900):
        dp.expand(chunk)
        st = dp.states()
        for q, o in enumerate(ops):
            o.iterate(chunk)
            h, sm = dp.treehash(q)
            assert h == o.treehash()
            one = dp.state(q)
            assert np.array_equal(st["D"][q].view(np.uint64), one["D"].view(np.uint64))
            assert np.array_equal(st["conn"][q], one["conn"])
            assert np.array_equal(st["P"][q].view(np.uint64), one["P"].view(np.uint64))
            assert np.array_equal(st["is_done"][q], one["is_done"])
            assert np.array_equal(st["D"][q].view(np.uint64), o.matrices()["D"].view(np.uint64))
    for o in ops:
        o.close()
    dp.close(); dg.close()


@pytest.mark.parametrize("slots", [1, 3, 5])
def test_streaming_planner_matches_oracle_per_query(slots):
    """More queries than state slots: CTAs pull queries from the device queue, rebuild their slot,
    run the query to completion; what every query leaves (report, matrices, tree hash) equals the
    oracle's run of that query, and a second imomd_expand starts every query afresh."""
    g = graphs.jittered_grid(120, seed=9)
    K, Q = 5, 11
    dests = [graphs.pick_destinations(g, K, seed=20 + q) for q in range(Q)]
    ops = [ol.OraclePlanner(g, d) for d in dests]
    dg = capi.DeviceGraph(g)
    dp = capi.DevicePlanner(dg, dests, max_resident=slots)
    assert dp.info()["slots"] == slots and dp.info()["streaming"]
    dp.feed(capi.mt19937_words(0, 400000))
    want = []
    for o in ops:
        e = o.iterate(700)
        want.append((e, o.treehash(), o.matrices()))
    for _ in range(2):
        reps = dp.expand(700)
        st = dp.states()
        for q in range(Q):
            e, h, m = want[q]
            assert reps[q].status == 0 and reps[q].iterations == 700 and reps[q].expansions == e
            assert reps[q].tree_hash == h == dp.treehash(q)[0]
            assert reps[q].unresolved_ties == 0 and reps[q].busy_cycles > 0
            assert np.array_equal(st["D"][q].view(np.uint64), m["D"].view(np.uint64))
            assert np.array_equal(st["P"][q].view(np.uint64), m["P"].view(np.uint64))
    with pytest.raises(capi.ImomdError, match="not resident"):
        dp.tree(0, 0)
    for o in ops:
        o.close()
    dp.close(); dg.close()


def test_streaming_after_set_queries():
    """new destinations for every query of a streaming planner, then a run (the bench's step)"""
    g = graphs.jittered_grid(90, seed=3)
    K, Q = 4, 6
    first = [graphs.pick_destinations(g, K, seed=q) for q in range(Q)]
    second = [graphs.pick_destinations(g, K, seed=100 + q) for q in range(Q)]
    dg = capi.DeviceGraph(g)
    dp = capi.DevicePlanner(dg, first, max_resident=2)
    dp.feed(capi.mt19937_words(0, 200000))
    dp.expand(300)
    probs = np.stack([capi.selection_probabilities(g, d) for d in second])
    dp.clear_words()
    dp.set_queries(np.ascontiguousarray(np.asarray(second, np.uint32).ravel()), np.ascontiguousarray(probs.ravel()))
    dp.feed(capi.mt19937_words(0, 200000))
    reps = dp.expand(300)
    for q, d in enumerate(second):
        o = ol.OraclePlanner(g, d)
        assert reps[q].expansions == o.iterate(300) and reps[q].tree_hash == o.treehash()
        o.close()
    dp.close(); dg.close()


def test_widened_tie_band_is_certified_by_the_host():
    """tie band 1e-5: many nearest / heuristic decisions go to the host mailboxes while the kernel
    runs (mapped pinned memory, answered by imomd_sync's wait loop with glibc); the trees still
    equal the oracle's bit for bit and nothing stays unresolved."""
    g = graphs.jittered_grid(60, seed=11)
    dest = graphs.pick_destinations(g, 6, seed=5)
    op = ol.OraclePlanner(g, dest, 0.7)
    dg = capi.DeviceGraph(g)
    dp = capi.DevicePlanner(dg, [dest], None, 0.7, tie_band=1e-5)
    dp.feed(capi.mt19937_words(0, 300000))
    rep = dp.expand(600)[0]
    assert rep.status == 0 and rep.expansions == op.iterate(600)
    compare_with_oracle(dp, op, 6, False)
    assert rep.certified_ties > 20 and rep.unresolved_ties == 0
    assert dp.info()["certified_requests"] >= rep.certified_ties
    # the lookup hook shares the service (lookups exactly between two frontier vertices would need
    # it; here it just must keep working next to it)
    lk = np.zeros(64, capi.NN_DTYPE)
    lk["tree"] = np.arange(64) % 6
    lk["lat"] = np.linspace(g.lat.min(), g.lat.max(), 64); lk["lon"] = np.linspace(g.lon.min(), g.lon.max(), 64)
    idx, _ = dp.nearest(lk)
    for i in range(64):
        ok, want, d, ties, _ = op.nearest(int(lk["tree"][i]), float(lk["lat"][i]), float(lk["lon"][i]))
        assert int(idx[i]) == want
    dp.close(); dg.close(); op.close()


def test_many_concurrent_certifications_streaming():
    """every CTA of a streaming launch talks to its own mailbox at the same time"""
    g = graphs.jittered_grid(80, seed=4)
    K, Q = 4, 40
    dests = [graphs.pick_destinations(g, K, seed=300 + q) for q in range(Q)]
    dg = capi.DeviceGraph(g)
    dp = capi.DevicePlanner(dg, dests, goal_bias=0.5, max_resident=16, tie_band=1e-5)
    dp.feed(capi.mt19937_words(0, 300000))
    reps = dp.expand(400)
    assert sum(r.certified_ties for r in reps) > 100
    for q in range(0, Q, 5):
        o = ol.OraclePlanner(g, dests[q], 0.5)
        assert reps[q].expansions == o.iterate(400) and reps[q].tree_hash == o.treehash()
        assert reps[q].unresolved_ties == 0
        o.close()
    dp.close(); dg.close()


def test_full_size_grid_1m_full_depth_matches_reference():
    """BASELINE configs[2] at the depth the bench times (1000 x 1000 grid, K = 12, 3000 passes):
    TREEHASH, expansions, tree sizes and the distance matrix of the CUDA path equal the oracle's,
    and -- where the compiled reference travelled with the snapshot -- the unmodified reference's."""
    g = graphs.jittered_grid(1000, seed=123)
    dest = graphs.pick_destinations(g, 12, seed=7)
    dg = capi.DeviceGraph(g)
    dp = capi.DevicePlanner(dg, [dest])
    dp.feed(capi.mt19937_words(0, 3000 * 12 * 8 + 4096))
    rep = dp.expand(3000)[0]
    assert rep.status == 0 and rep.unresolved_ties == 0
    h, _ = dp.treehash(0)
    op = ol.OraclePlanner(g, dest)
    assert rep.expansions == op.iterate(3000)
    assert h == op.treehash()
    compare_with_oracle(dp, op, 12, False, check_frontier=False)
    op.close()
    if os.path.exists(REF_PROBE):
        fd, gf = tempfile.mkstemp(suffix=".bin", prefix="imomd_graph_")
        os.close(fd)
        try:
            g.write_probe_file(gf)
            out = json.loads(subprocess.check_output(
                [REF_PROBE, "--graph", gf, "--dest", ",".join(map(str, dest)), "--iters", "3000",
                 "--goal-bias", "1.0"], text=True).strip().splitlines()[-1])
        finally:
            os.unlink(gf)
        assert int(out["treehash"], 16) == h
        assert out["expansions"] == rep.expansions and out["tree_vertices"] == rep.tree_vertices
    dp.close(); dg.close()


def test_config3_road_graph_10m_k22_prefix():
    """BASELINE configs[3] shape on one GPU: 10 M-vertex road-like graph (71 % degree-2 chains:
    jump-point search), 20 destinations (K = 22 trees); 400 passes against the oracle."""
    g = graphs.road_like(10_000_000, seed=11)
    K = 22
    dest = graphs.pick_destinations(g, K, seed=100)
    dg = capi.DeviceGraph(g)
    dp = capi.DevicePlanner(dg, [dest])
    dp.feed(capi.mt19937_words(0, 400 * K * 8 + 4096))
    rep = dp.expand(400)[0]
    op = ol.OraclePlanner(g, dest)
    assert rep.status == 0 and rep.expansions == op.iterate(400)
    assert dp.treehash(0)[0] == op.treehash()
    compare_with_oracle(dp, op, K, False, check_frontier=False)
    op.close(); dp.close(); dg.close()


def test_ga_eval_reports_bad_tours_from_the_device():
    ctx = capi.GaContext(0)
    K = 6
    rng = np.random.default_rng(0)
    D = rng.uniform(1.0, 9.0, (K, K))
    tours = rng.integers(0, K, (50, 12)).astype(np.int32)
    lens = np.full(50, 12, np.int32)
    ctx.eval(D, tours, lens)
    bad = tours.copy(); bad[17, 5] = K
    with pytest.raises(capi.ImomdError, match="tour 17"):
        ctx.eval(D, bad, lens)
    bad_len = lens.copy(); bad_len[3] = 13
    with pytest.raises(capi.ImomdError, match="tour 3"):
        ctx.eval(D, tours, bad_len)
    # the context keeps working afterwards
    got = ctx.eval(D, tours, lens)
    want = np.array([sum(D[a, b] for a, b in zip(t[:-1], t[1:])) for t in tours])
    assert np.allclose(got, want, rtol=1e-15)
    ctx.close()
