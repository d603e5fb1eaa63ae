"""Host logic of the CUDA path: csrc/imomd_core.cuh compiled for the host (one-thread Cta
policy, tests/hostsim) against the oracle restatement -- no GPU needed.  The same checks
run against the real kernels in tests/test_cuda_parity.py (-m gpu)."""
import numpy as np
import pytest

import hostsimlib as hs
import oraclelib as ol
from imomd_b200 import graphs
from parity import compare_with_oracle


def _words(n, seed=0):
    w = np.empty(n, np.uint32)
    ol.oracle().orc_mt19937_words(seed, n, w)
    return w


def _pair(g, dest, goal_bias=1.0, pseudo=False, grid_dim=0):
    op = ol.OraclePlanner(g, dest, goal_bias, pseudo)
    P0 = op.matrices()["P"]
    hp = hs.HostSimPlanner(g, [dest], [P0], goal_bias, pseudo, grid_dim)
    return op, hp


@pytest.mark.parametrize("width,K,grid_dim,gb", [(60, 6, 0, 1.0), (60, 6, 16, 1.0), (90, 12, 8, 1.0),
                                                 (50, 4, 4, 0.5), (70, 2, 0, 1.0), (40, 5, 32, 0.0)])
def test_grid_expansion_matches_oracle(width, K, grid_dim, gb):
    g = graphs.jittered_grid(width, seed=width)
    dest = graphs.pick_destinations(g, K, seed=K)
    op, hp = _pair(g, dest, gb, grid_dim=grid_dim)
    compare_with_oracle(hp, op, K, True)
    hp.feed(_words(400000))
    total = 0
    for chunk in (1, 7, 92, 400):
        rep = hp.expand(chunk)[0]
        e = op.iterate(chunk)
        assert rep.status == 0 and rep.iterations == chunk and rep.expansions == e
        total += chunk
        compare_with_oracle(hp, op, K, True)
        for k in range(K):
            assert hp.check_hash(k) == 0
    assert rep.unresolved_ties == 0
    assert rep.tree_vertices == sum(int((op.tree(k)[0] != 0xFFFFFFFF).sum()) for k in range(K))
    op.close(); hp.close()


def test_roadlike_jps_and_pseudo_mode():
    """degree-2 chains (JPS) + pseudo-mode connection log"""
    g = graphs.road_like(6000, seed=2)
    dest = graphs.pick_destinations(g, 6, seed=1)
    for pseudo in (False, True):
        op, hp = _pair(g, dest, 1.0, pseudo)
        hp.feed(_words(300000))
        for chunk in (50, 450, 1000):
            rep = hp.expand(chunk)[0]
            assert rep.status == 0 and rep.expansions == op.iterate(chunk)
            compare_with_oracle(hp, op, 6, True)
        if pseudo:
            a, b = hp.connection_log(), op.connection_log()
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].astype(np.uint64), b[1])
        op.close(); hp.close()


def test_resume_after_word_starvation():
    """NEED_WORDS stops a query inside a pass and the next call resumes at the same tree."""
    g = graphs.jittered_grid(50, seed=5)
    dest = graphs.pick_destinations(g, 5, seed=2)
    op, hp = _pair(g, dest)
    words = _words(200000)
    fed, done, target = 0, 0, 300
    guard = 0
    while done < target:
        hp.feed(words[fed:fed + 700]); fed += 700
        rep = hp.expand(target - done)[0]
        done += rep.iterations
        assert rep.status in (0, 1)
        guard += 1
        assert guard < 1000
    assert guard > 3            # it really was starved several times
    op.iterate(target)
    compare_with_oracle(hp, op, 5, True)
    op.close(); hp.close()


def test_nearest_hook_prunes_and_matches():
    g = graphs.jittered_grid(120, seed=11)
    dest = graphs.pick_destinations(g, 4, seed=4)
    op, hp = _pair(g, dest, grid_dim=32)
    hp.feed(_words(200000))
    hp.expand(1500); op.iterate(1500)
    rng = np.random.default_rng(0)
    examined = []
    for _ in range(400):
        k = int(rng.integers(0, 4))
        lat = float(rng.uniform(g.lat.min(), g.lat.max()))
        lon = float(rng.uniform(g.lon.min(), g.lon.max()))
        ok, idx, d, ties, second = op.nearest(k, lat, lon)
        i2, d2, st = hp.nearest(k, lat, lon)
        assert ok and idx == i2 and d == d2
        examined.append(st[2])
    n_chunks_total = sum(len(hp.frontier(k)) for k in range(4)) / 32
    assert np.mean(examined) < n_chunks_total   # the cell bound does discard work
    op.close(); hp.close()


def test_multi_query_batch_independent():
    """Queries of one planner do not interact (config 5: many queries, K = 2)."""
    g = graphs.road_like(3000, seed=7)
    rng = np.random.default_rng(3)
    dests = [graphs.pick_destinations(g, 2, seed=int(s)) for s in rng.integers(0, 1000, 5)]
    ops = [ol.OraclePlanner(g, d) for d in dests]
    probs = [o.matrices()["P"] for o in ops]
    hp = hs.HostSimPlanner(g, dests, probs)
    hp.feed(_words(100000))
    reps = hp.expand(400)
    for q, o in enumerate(ops):
        assert reps[q].expansions == o.iterate(400)
        compare_with_oracle(hp, o, 2, True, q=q)
        o.close()
    hp.close()


@pytest.mark.parametrize("width,hubs,spokes", [(40, 10, 3), (40, 14, 9)])
def test_wide_adjacency_rows(width, hubs, spokes):
    """max degree 5..8 (8-slot rows) and 9..13 (16-slot rows)"""
    g = graphs.hub_grid(width, seed=4, hubs=hubs, spokes=spokes)
    assert g.degree().max() > (8 if spokes > 4 else 4)
    dest = graphs.pick_destinations(g, 5, seed=2)
    op, hp = _pair(g, dest)
    hp.feed(_words(200000))
    for chunk in (100, 700):
        rep = hp.expand(chunk)[0]
        assert rep.status == 0 and rep.expansions == op.iterate(chunk)
        compare_with_oracle(hp, op, 5, True)
    op.close(); hp.close()
