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


@pytest.mark.ref
@pytest.mark.parametrize("kind,K,warm", [("grid", 4, 300), ("road", 5, 500), ("grid", 7, 150)])
def test_attach_list_matches_reference_merge_step(kind, K, warm):
    """The per-vertex work of the pseudo-tree merge (imomd_attach_list: connectNewNode_ +
    updateExpandables + rewireTree_ + updateConnectionTree_, src/imomd_rrt_star.cpp:848-851
    upstream) against the UNMODIFIED reference performing the same calls: after a warm-up,
    frontier vertices of every tree are attached in runs of several vertices per launch; trees,
    frontiers, matrices, flags, the connection-node sets and the visit order (= insertion order
    of the reference's parent map, replayed into a real container by the facade) all agree."""
    R = ol.ref()
    g = graphs.jittered_grid(60, seed=11) if kind == "grid" else graphs.road_like(4000, seed=9)
    dest = graphs.pick_destinations(g, K, seed=K + 3)
    gh = R.ref_graph_from_edges(g.n, g.lat, g.lon, len(g.eu), g.eu, g.ev,
                                g.ew.ctypes.data_as(ol.C.c_void_p))
    rp = ol.RefPlanner(gh, dest, pseudo=True)
    P0 = rp.matrices()["P"]
    hp = hs.HostSimPlanner(g, [dest], [P0], 1.0, True, 0)
    hp.feed(_words(400000))
    assert hp.expand(warm)[0].expansions == rp.iterate(warm)
    compare_with_oracle(hp, rp, K, False)
    rng = np.random.default_rng(K)
    for round_ in range(6):
        for k in range(K):
            todo = []
            for _ in range(int(rng.integers(1, 6))):
                fr = rp.frontier(k)
                if len(fr) == 0:
                    break
                x = int(fr[int(rng.integers(0, len(fr)))])
                rp.attach(k, x)                       # the reference does it vertex by vertex
                todo.append(x)
            if todo:
                got = hp.attach_list(k, todo)         # the device logic: one launch for the run
                assert (got != 0xFFFFFFFF).all()
        compare_with_oracle(hp, rp, K, False)
    # connection-node sets: the logged insertions, deduplicated, are the reference's sets
    si, vx = hp.connection_log()
    for idx in range(K * (K - 1) // 2):
        assert sorted(set(int(v) for v in vx[si == idx])) == sorted(rp.connection_set(idx))
    # visit order: the same vertices the reference's parent map holds; replayed into a real
    # unordered_map it iterates like the reference's (checked end to end in test_cuda_facade.py)
    for k in range(K):
        assert sorted(int(v) for v in hp.visit_order(k)) == sorted(int(v) for v in rp.parent_order(k))
    hp.close()


def test_frontier_with_more_occupied_cells_than_the_bound_table():
    """The CTA keeps the cell bounds of a search for at most 256 occupied cells (kCtaLbCap); a
    frontier that occupies more (64 x 64 hash over a small map, uniform samples: long thin trees)
    recomputes its bounds in the second pass of the search.  Same trees as the oracle either way."""
    g = graphs.jittered_grid(120, seed=13)
    dest = graphs.pick_destinations(g, 3, seed=2)
    op, hp = _pair(g, dest, 0.0, grid_dim=64)
    hp.feed(_words(400000))
    e = hp.expand(2500)[0].expansions
    assert e == op.iterate(2500)
    # the premise: some tree's frontier is spread over more than 256 cells of the 64 x 64 hash
    cell = lambda v: (np.minimum(((g.lat[v] - g.lat.min()) / (np.ptp(g.lat) / 64)).astype(int), 63) * 64 +
                      np.minimum(((g.lon[v] - g.lon.min()) / (np.ptp(g.lon) / 64)).astype(int), 63))
    assert max(len(set(cell(np.asarray(hp.frontier(k), np.int64)).tolist())) for k in range(3)) > 256
    compare_with_oracle(hp, op, 3, True)
    op.close(); hp.close()


def test_device_treehash_matches_oracle():
    """tree_hash (the routine behind imomd_get_treehash and the streaming planners' reports) ==
    the SURVEY Appendix A TREEHASH of the oracle's trees."""
    g = graphs.road_like(5000, seed=4)
    dest = graphs.pick_destinations(g, 5, seed=3)
    op, hp = _pair(g, dest)
    hp.feed(_words(200000))
    seen = set()
    for chunk in (0, 3, 200, 700):
        if chunk:
            hp.expand(chunk); op.iterate(chunk)
        h, sm = hp.treehash()
        assert h == op.treehash()
        seen.add((h, sm))
    assert len(seen) == 4          # the fingerprints move with the trees
    op.close(); hp.close()


@pytest.mark.parametrize("slots,kind", [(1, "grid"), (2, "grid"), (3, "grid"), (2, "road")])
def test_streaming_slots_reproduce_every_query(slots, kind):
    """A streaming planner (fewer state slots than queries) rebuilds a slot for every query it
    takes -- only the records the previous query marked in the trees' dirty bitmaps -- and hashes
    its trees through the same bitmaps: per-query results equal the oracle's, whatever ran in the
    slot before ("road": jump-point hops visit vertices that were never in the frontier)."""
    g = graphs.jittered_grid(48, seed=9) if kind == "grid" else graphs.road_like(4000, seed=6)
    dests = [graphs.pick_destinations(g, 4, seed=20 + q) for q in range(5)]
    ops = [ol.OraclePlanner(g, d) for d in dests]
    probs = [o.matrices()["P"] for o in ops]
    hp = hs.HostSimPlanner(g, dests, probs, slots=slots)
    hp.feed(_words(200000))
    for rounds in range(2):          # a second expand() starts every query afresh
        reps = hp.expand(250)
        for q, o in enumerate(ops):
            if rounds == 0:
                e = o.iterate(250)
                assert reps[q].expansions == e
            assert reps[q].status == 0 and reps[q].iterations == 250
            assert reps[q].tree_hash == o.treehash() == hp.treehash(q)[0]
            s = hp.state(q)
            assert np.array_equal(s["D"].view(np.uint64), o.matrices()["D"].view(np.uint64))
            assert np.array_equal(s["P"].view(np.uint64), o.matrices()["P"].view(np.uint64))
    for o in ops:
        o.close()
    hp.close()


def test_widened_tie_band_is_certified_exactly():
    """With the tie band widened to 1e-5 a visible share of the arg-min / heuristic decisions goes
    through the exact-value service (Cta::certify: the reference's haversine on glibc); the trees
    must still equal the oracle's bit for bit, with nothing left unresolved."""
    g = graphs.jittered_grid(60, seed=11)
    dest = graphs.pick_destinations(g, 6, seed=5)
    op = ol.OraclePlanner(g, dest, 0.7)
    hp = hs.HostSimPlanner(g, [dest], [op.matrices()["P"]], 0.7, tie_band=1e-5)
    hp.feed(_words(300000))
    rep = hp.expand(600)[0]
    assert rep.expansions == op.iterate(600)
    compare_with_oracle(hp, op, 6, True)
    assert rep.certified_ties > 20 and hp.certified() >= rep.certified_ties
    assert rep.unresolved_ties == 0
    # detached service: the same decisions are counted as unresolved instead
    hp2 = hs.HostSimPlanner(g, [dest], [op.matrices()["P"]], 0.7, tie_band=1e-5)
    hp2.set_certify(False)
    hp2.feed(_words(300000))
    rep2 = hp2.expand(600)[0]
    assert rep2.certified_ties == 0 and rep2.unresolved_ties > 20 and hp2.certified() == 0
    op.close(); hp.close(); hp2.close()


@pytest.mark.parametrize("kind,K,gb,B", [("grid60", 6, 1.0, 1), ("grid60", 6, 1.0, 4), ("grid90", 12, 1.0, 16),
                                         ("road", 5, 1.0, 8), ("road", 7, 0.5, 3), ("grid60", 4, 0.3, 32)])
def test_batched_mode_matches_its_restatement(kind, K, gb, B):
    """imomd_expand_batch's device logic (attach per tree, rewiring as a label-correcting
    relaxation, connection checks with the final costs, selection-probability tests) against the
    oracle's restatement of the same semantics: trees bit for bit, K x K state."""
    g = {"grid60": lambda: graphs.jittered_grid(60, seed=2), "grid90": lambda: graphs.jittered_grid(90, seed=3),
         "road": lambda: graphs.road_like(6000, seed=5)}[kind]()
    dest = graphs.pick_destinations(g, K, seed=11)
    op = ol.OraclePlanner(g, dest, gb)
    hp = hs.HostSimPlanner(g, [dest], [op.matrices()["P"]], gb)
    hp.feed(_words(400000))
    for steps in (1, 7, 60):
        rep = hp.expand_batch(steps, B)[0]
        e = op.iterate_batched(steps, B)
        assert rep.status == 0 and rep.expansions == e and rep.iterations == steps * B
        assert hp.treehash()[0] == op.treehash()
        compare_with_oracle(hp, op, K, True)
    op.close(); hp.close()
