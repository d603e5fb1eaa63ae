"""Parity of the CUDA path (through the C ABI, libimomd_b200.so) against the oracle
restatement on the same seeded inputs.  Bars (BASELINE.json north_star): nearest-vertex
indices bit-exact, batch-size-1 parent arrays exact, costs within 1e-9 relative -- here
costs are compared BIT FOR BIT (they are sums of host-supplied edge weights); only the
min-heuristic *values*, which go through CUDA's libm, get a 1e-12 relative tolerance."""
import numpy as np
import pytest

import oraclelib as ol
from imomd_b200 import capi, graphs
from parity import compare_with_oracle

pytestmark = pytest.mark.gpu


def _pair(g, dest, goal_bias=1.0, pseudo=False, grid_dim=0):
    op = ol.OraclePlanner(g, dest, goal_bias, pseudo)
    dg = capi.DeviceGraph(g, 0, grid_dim)
    dp = capi.DevicePlanner(dg, [dest], None, goal_bias, pseudo)
    return op, dg, dp


@pytest.mark.parametrize("width,K,grid_dim,gb", [(60, 6, 0, 1.0), (60, 6, 16, 1.0), (90, 12, 8, 1.0),
                                                 (50, 4, 4, 0.5), (70, 2, 0, 1.0), (40, 5, 32, 0.0),
                                                 (150, 12, 0, 1.0)])
def test_grid_expansion_matches_oracle(width, K, grid_dim, gb):
    g = graphs.jittered_grid(width, seed=width)
    dest = graphs.pick_destinations(g, K, seed=K)
    op, dg, dp = _pair(g, dest, gb, grid_dim=grid_dim)
    compare_with_oracle(dp, op, K, False)
    dp.feed(capi.mt19937_words(0, 600000))
    for chunk in (1, 7, 92, 400, 1000):
        rep = dp.expand(chunk)[0]
        e = op.iterate(chunk)
        assert rep.status == 0 and rep.iterations == chunk and rep.expansions == e
        compare_with_oracle(dp, op, K, False)
    assert rep.unresolved_ties == 0
    dp.close(); dg.close(); op.close()


def test_roadlike_jps_and_pseudo_mode():
    g = graphs.road_like(6000, seed=2)
    dest = graphs.pick_destinations(g, 6, seed=1)
    for pseudo in (False, True):
        op, dg, dp = _pair(g, dest, 1.0, pseudo)
        dp.feed(capi.mt19937_words(0, 300000))
        for chunk in (50, 450, 1000):
            rep = dp.expand(chunk)[0]
            assert rep.status == 0 and rep.expansions == op.iterate(chunk)
            compare_with_oracle(dp, op, 6, False)
        if pseudo:
            a, b = dp.connection_log(), op.connection_log()
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].astype(np.uint64), b[1])
        dp.close(); dg.close(); op.close()


def test_nearest_indices_bit_exact():
    """10^4 random (tree state, sample) lookups: indices identical, distances to 1e-12."""
    g = graphs.jittered_grid(200, seed=11)
    dest = graphs.pick_destinations(g, 6, seed=4)
    op, dg, dp = _pair(g, dest)
    dp.feed(capi.mt19937_words(0, 400000))
    rng = np.random.default_rng(0)
    total = 0
    for chunk in (200, 800, 2000):
        dp.expand(chunk); op.iterate(chunk)
        n = 3500
        lk = np.empty(n, capi.NN_DTYPE)
        lk["query"] = 0
        lk["tree"] = rng.integers(0, 6, n)
        lk["lat"] = rng.uniform(g.lat.min() - 1e-3, g.lat.max() + 1e-3, n)
        lk["lon"] = rng.uniform(g.lon.min() - 1e-3, g.lon.max() + 1e-3, n)
        # a third of the samples sit exactly on destination roots (the goal-biased case)
        roots = rng.integers(0, 6, n // 3)
        lk["lat"][: n // 3] = g.lat[np.asarray(dest)[roots]]
        lk["lon"][: n // 3] = g.lon[np.asarray(dest)[roots]]
        idx, dist = dp.nearest(lk)
        for i in range(n):
            ok, want, d, ties, second = op.nearest(int(lk["tree"][i]), float(lk["lat"][i]), float(lk["lon"][i]))
            assert ok and ties == 1
            assert int(idx[i]) == want, (i, idx[i], want, dist[i], d, second)
            assert abs(dist[i] - d) <= 1e-12 * max(d, 1.0)
        total += n
    assert total >= 10000
    dp.close(); dg.close(); op.close()


def test_multi_query_batch():
    """config 5 shape: many independent queries with K = 2 on one road graph."""
    g = graphs.road_like(3000, seed=7)
    rng = np.random.default_rng(3)
    Q = 40
    dests = [graphs.pick_destinations(g, 2, seed=int(s)) for s in rng.integers(0, 100000, Q)]
    dg = capi.DeviceGraph(g)
    dp = capi.DevicePlanner(dg, dests)
    dp.feed(capi.mt19937_words(0, 100000))
    reps = dp.expand(400)
    for q in range(Q):
        o = ol.OraclePlanner(g, dests[q])
        assert reps[q].status == 0 and reps[q].expansions == o.iterate(400)
        compare_with_oracle(dp, o, 2, False, q=q)
        o.close()
    dp.close(); dg.close()


def test_reset_reproduces():
    g = graphs.jittered_grid(60, seed=3)
    dest = graphs.pick_destinations(g, 5, seed=9)
    op, dg, dp = _pair(g, dest)
    dp.feed(capi.mt19937_words(0, 100000))
    dp.expand(500)
    a = [dp.tree(k) for k in range(5)]
    dp.reset()
    dp.expand(200); dp.expand(300)
    for k in range(5):
        p2, c2 = dp.tree(k)
        assert np.array_equal(a[k][0], p2) and np.array_equal(a[k][1].view(np.uint64), c2.view(np.uint64))
    op.iterate(500)
    compare_with_oracle(dp, op, 5, False)
    dp.close(); dg.close(); op.close()


def test_profiling_counters_are_opt_in():
    """imomd_get_profile / imomd_get_phase_cycles: the kernels keep the per-request cycle counters
    only after imomd_set_profiling(p, 1); results do not depend on the switch."""
    g = graphs.jittered_grid(50, seed=8)
    dest = graphs.pick_destinations(g, 4, seed=4)
    op, dg, dp = _pair(g, dest)
    dp.feed(capi.mt19937_words(0, 100000))
    dp.expand(150)
    prof = dp.profile(0); prof.pop("bfs_vertices_levels")
    assert all(c == 0 and n == 0 for c, n in prof.values())
    dp.set_profiling(True)
    dp.expand(150)
    prof = dp.profile(0); prof.pop("bfs_vertices_levels")
    assert prof["nn"][1] > 0 and prof["nn"][0] > 0 and prof["serial"][0] > 0
    assert dp.phase_cycles(0)[0].sum() > 0
    op.iterate(300)
    compare_with_oracle(dp, op, 4, False)
    dp.close(); dg.close(); op.close()


def test_walk_to_root_and_word_starvation():
    g = graphs.jittered_grid(50, seed=5)
    dest = graphs.pick_destinations(g, 5, seed=2)
    op, dg, dp = _pair(g, dest)
    words = capi.mt19937_words(0, 200000)
    fed = done = 0
    starved = 0
    while done < 300:
        dp.feed(words[fed:fed + 900]); fed += 900
        rep = dp.expand(300 - done)[0]
        done += rep.iterations
        starved += rep.status == 1
        assert rep.status in (0, 1)
    assert starved > 2
    op.iterate(300)
    compare_with_oracle(dp, op, 5, False)
    par, _ = dp.tree(0)
    vs = np.nonzero(par != capi.NONE)[0]
    v = int(vs[len(vs) // 2])
    path = dp.walk_to_root(0, v)
    assert path[0] == v and path[-1] == dest[0]
    for a, b in zip(path[:-1], path[1:]):
        assert par[a] == b
    dp.close(); dg.close(); op.close()


def test_ga_eval_bit_exact():
    ctx = capi.GaContext(0)
    rng = np.random.default_rng(4)
    for K in (4, 12, 22, 32):
        D = rng.uniform(100.0, 5000.0, (K, K)); D = (D + D.T) / 2
        D[rng.random((K, K)) < 0.2] = np.inf
        np.fill_diagonal(D, 0.0)
        n, stride = 15000, 64
        lens = rng.integers(1, min(stride, 2 * K + 2), n).astype(np.int32)
        tours = rng.integers(0, K, (n, stride)).astype(np.int32)
        got = ctx.eval(D, tours, lens)
        want = np.empty(n)
        ol.oracle().orc_path_costs(K, np.ascontiguousarray(D).ravel(), n, stride, tours.ravel(), lens, want)
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), K
    ctx.close()


def test_full_size_grid_1m_matches_oracle_prefix():
    """BASELINE config 3 (1000 x 1000 grid, K = 12): the first 600 passes against the
    oracle (hash of every parent and cost bit), then structural invariants at depth."""
    g = graphs.jittered_grid(1000, seed=123)
    dest = graphs.pick_destinations(g, 12, seed=7)
    op, dg, dp = _pair(g, dest)
    dp.feed(capi.mt19937_words(0, 2_000_000))
    rep = dp.expand(600)[0]
    assert rep.status == 0 and rep.expansions == op.iterate(600)
    compare_with_oracle(dp, op, 12, False, check_frontier=False)
    rep = dp.expand(2400)[0]
    assert rep.status == 0
    deg = g.degree()
    for k in (0, 5, 11):
        par, cost = dp.tree(k)
        vis = par != capi.NONE
        v = np.nonzero(vis)[0]
        v = v[v != dest[k]]
        # every parent is a graph neighbour and is itself visited
        rows = [g.col[g.row_ptr[x]:g.row_ptr[x + 1]] for x in v[:2000]]
        assert all(par[x] in r for x, r in zip(v[:2000], rows))
        assert vis[par[v]].all()
        # frontier = unvisited neighbours of visited vertices
        fr = dp.frontier(k)
        assert not vis[fr.astype(np.int64)].any()
        nb = np.unique(np.concatenate([g.col[g.row_ptr[x]:g.row_ptr[x + 1]] for x in np.nonzero(vis)[0]]))
        assert np.array_equal(np.sort(nb[~vis[nb]]).astype(np.uint64), fr)
    dp.close(); dg.close(); op.close()


def test_nearest_batch_tile_kernel_matches_exact_path():
    """imomd_nearest_batch (bulk-async staged brute-force tiles) returns the same indices as the
    pruned exact path and the oracle, over several queries / trees and ragged tile sizes."""
    g = graphs.jittered_grid(150, seed=8)
    rng = np.random.default_rng(5)
    Q, K = 5, 6
    dests = [graphs.pick_destinations(g, K, seed=int(s)) for s in rng.integers(0, 10**6, Q)]
    dg = capi.DeviceGraph(g)
    dp = capi.DevicePlanner(dg, dests)
    dp.feed(capi.mt19937_words(0, 300000))
    dp.expand(1200)
    n = 4000
    lk = np.empty(n, capi.NN_DTYPE)
    lk["query"] = rng.integers(0, Q, n); lk["tree"] = rng.integers(0, K, n)
    lk["lat"] = rng.uniform(g.lat.min(), g.lat.max(), n); lk["lon"] = rng.uniform(g.lon.min(), g.lon.max(), n)
    # some trees get a single lookup, one gets hundreds (ragged tiles)
    lk["query"][:600] = 2; lk["tree"][:600] = 3
    idx_a, dist_a = dp.nearest(lk)
    idx_b, dist_b, info = capi.nearest_batch(dp, lk)
    assert np.array_equal(idx_a, idx_b)
    assert np.array_equal(dist_a.view(np.uint64), dist_b.view(np.uint64))
    assert info["bytes"] > 0 and info["kernel_ms"] > 0
    o = ol.OraclePlanner(g, dests[2]); o.iterate(1200)
    for i in range(0, 600, 7):
        ok, want, d, ties, _ = o.nearest(3, float(lk["lat"][i]), float(lk["lon"][i]))
        assert int(idx_b[i]) == want
    o.close(); dp.close(); dg.close()


def test_nearest_batch_tiny_frontiers():
    """frontiers of one to four records (fresh planner: the roots' neighbours; fewer than two chunk
    heads, so the screen has no bound and everything is ranked exactly), and an empty batch"""
    g = graphs.jittered_grid(60, seed=2)
    K = 4
    dests = [graphs.pick_destinations(g, K, seed=s) for s in (3, 4, 5)]
    dg = capi.DeviceGraph(g)
    dp = capi.DevicePlanner(dg, dests)
    dp.feed(capi.mt19937_words(0, 20000))
    rng = np.random.default_rng(1)
    for passes in (0, 1, 3):
        if passes:
            dp.expand(passes if passes == 1 else 2)
        n = 3 * K * 5
        lk = np.empty(n, capi.NN_DTYPE)
        lk["query"] = np.repeat(np.arange(3), K * 5); lk["tree"] = np.tile(np.repeat(np.arange(K), 5), 3)
        lk["lat"] = rng.uniform(g.lat.min(), g.lat.max(), n); lk["lon"] = rng.uniform(g.lon.min(), g.lon.max(), n)
        idx_a, dist_a = dp.nearest(lk)
        idx_b, dist_b, info = capi.nearest_batch(dp, lk)
        assert (idx_a != capi.NONE).all()
        assert np.array_equal(idx_a, idx_b)
        assert np.array_equal(dist_a.view(np.uint64), dist_b.view(np.uint64))
        # a batch that touches one tree only: the other trees are skipped by the preparation
        one = lk[lk["tree"] == 2][:3]
        idx_c, dist_c, _ = capi.nearest_batch(dp, one)
        assert np.array_equal(idx_c, idx_a[lk["tree"] == 2][:3])
    idx_e, dist_e, _ = capi.nearest_batch(dp, lk[:0])
    assert len(idx_e) == 0 and len(dist_e) == 0
    dp.close(); dg.close()


def test_nearest_batch_multi_segment_frontiers():
    """frontiers of several hundred chunks: every tile is cut into several segments that different
    CTAs rank; the merged answers equal the pruned exact path bit for bit and the oracle's indices"""
    g = graphs.jittered_grid(400, seed=3)
    K = 4
    dests = [graphs.pick_destinations(g, K, seed=s) for s in (1, 2)]
    dg = capi.DeviceGraph(g)
    dp = capi.DevicePlanner(dg, dests)
    dp.feed(capi.mt19937_words(0, 400000))
    dp.expand(6000)
    rng = np.random.default_rng(9)
    n = 2 * K * 7                       # ragged: 7 lookups per frontier = tiles of 4 + 3
    lk = np.empty(n, capi.NN_DTYPE)
    lk["query"] = np.repeat(np.arange(2), K * 7); lk["tree"] = np.tile(np.repeat(np.arange(K), 7), 2)
    lk["lat"] = rng.uniform(g.lat.min(), g.lat.max(), n); lk["lon"] = rng.uniform(g.lon.min(), g.lon.max(), n)
    # a few samples that coincide with frontier vertices (distance 0) and with visited vertices
    fr = dp.frontier(1, 0)
    lk["lat"][:3] = g.lat[fr[:3].astype(np.int64)]; lk["lon"][:3] = g.lon[fr[:3].astype(np.int64)]
    idx_a, dist_a = dp.nearest(lk)
    idx_b, dist_b, info = capi.nearest_batch(dp, lk)
    assert info["segments"] >= 2 * (2 * K * 2), info          # two or more segments per tile
    assert np.array_equal(idx_a, idx_b)
    assert np.array_equal(dist_a.view(np.uint64), dist_b.view(np.uint64))
    o = ol.OraclePlanner(g, dests[0]); o.iterate(6000)
    for i in range(0, K * 7):
        ok, want, d, ties, _ = o.nearest(int(lk["tree"][i]), float(lk["lat"][i]), float(lk["lon"][i]))
        assert int(idx_b[i]) == want
    o.close(); dp.close(); dg.close()


@pytest.mark.parametrize("width,hubs,spokes", [(40, 10, 3), (40, 14, 9)])
def test_wide_adjacency_rows(width, hubs, spokes):
    """max degree 5..8 (8-slot kernels) and 9..13 (16-slot kernels)"""
    g = graphs.hub_grid(width, seed=4, hubs=hubs, spokes=spokes)
    assert g.degree().max() > (8 if spokes > 4 else 4)
    dest = graphs.pick_destinations(g, 5, seed=2)
    op, dg, dp = _pair(g, dest)
    dp.feed(capi.mt19937_words(0, 200000))
    for chunk in (100, 700):
        rep = dp.expand(chunk)[0]
        assert rep.status == 0 and rep.expansions == op.iterate(chunk)
        compare_with_oracle(dp, op, 5, False)
    dp.close(); dg.close(); op.close()
