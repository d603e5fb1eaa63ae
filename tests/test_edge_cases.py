"""Edge cases of the expansion path, each on all three implementations: the unmodified reference
(oracle/_ref), the CPU restatement (oracle/) and the device logic -- on the host simulation here,
through the C ABI on a GPU in the `gpu`-marked twin.  Scenarios: the maximum number of trees
(30 objectives, K = 32), a source without neighbours, source and target in different components
(frontiers run dry, the planner idles), the smallest graph (two vertices), a vertex of the maximum
degree as a root."""
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


def _scenarios():
    out = {}
    g = graphs.jittered_grid(40, seed=8)
    out["k32"] = (g, graphs.pick_destinations(g, 32, seed=5), [1, 40, 160])
    # vertex 0 has no edge at all; a 5 x 5 patch next to it carries the other destinations
    lat = 37.0 + np.arange(26) * 1e-4 + np.random.default_rng(1).uniform(-2e-5, 2e-5, 26)
    lon = -122.0 + (np.arange(26) % 5) * 1e-4 + np.random.default_rng(2).uniform(-2e-5, 2e-5, 26)
    eu, ev = [], []
    for r in range(5):
        for c in range(5):
            v = 1 + r * 5 + c
            if c < 4: eu.append(v); ev.append(v + 1)
            if r < 4: eu.append(v); ev.append(v + 5)
    out["isolated_source"] = (graphs.from_edges(lat, lon, eu, ev), [0, 7, 25], [1, 10, 60])
    # two components: 0..11 (a path) and 12..35 (a 4 x 6 patch)
    n = 36
    lat = 37.0 + np.arange(n) * 7e-5 + np.random.default_rng(3).uniform(-1e-5, 1e-5, n)
    lon = -122.0 + (np.arange(n) % 6) * 9e-5 + np.random.default_rng(4).uniform(-1e-5, 1e-5, n)
    eu = list(range(0, 11)); ev = list(range(1, 12))
    for r in range(4):
        for c in range(6):
            v = 12 + r * 6 + c
            if c < 5: eu.append(v); ev.append(v + 1)
            if r < 3: eu.append(v); ev.append(v + 6)
    out["two_components"] = (graphs.from_edges(lat, lon, eu, ev), [0, 5, 20, 35], [1, 15, 80])
    out["two_vertices"] = (graphs.from_edges([37.0, 37.0004], [-122.0, -122.0003], [0], [1]), [0, 1], [1, 3, 10])
    g = graphs.hub_grid(30, seed=6, hubs=6, spokes=9)
    hub = int(np.argmax(g.degree()))
    rest = [d for d in graphs.pick_destinations(g, 4, seed=9) if d != hub][:3]
    out["hub_root"] = (g, [hub] + rest, [1, 50, 300])
    return out


SCENARIOS = _scenarios()


@pytest.mark.ref
@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_reference_oracle_and_device_logic_agree(name):
    g, dest, chunks = SCENARIOS[name]
    K = len(dest)
    R = ol.ref()
    gh = R.ref_graph_from_edges(g.n, g.lat, g.lon, len(g.eu), g.eu, g.ev, g.ew.ctypes.data_as(ol.C.c_void_p))
    rp, op = ol.RefPlanner(gh, dest), ol.OraclePlanner(g, dest)
    hp = hs.HostSimPlanner(g, [dest], [op.matrices()["P"]], 1.0, False, 0)
    hp.feed(_words(200000))
    compare_with_oracle(hp, rp, K, False)
    for chunk in chunks:
        e_ref, e_orc = rp.iterate(chunk), op.iterate(chunk)
        rep = hp.expand(chunk)[0]
        assert e_ref == e_orc == rep.expansions and rep.status == 0 and rep.iterations == chunk
        compare_with_oracle(hp, rp, K, False)
        compare_with_oracle(hp, op, K, True)
    if name in ("isolated_source", "two_components"):
        assert rp.flags()["connected"] == 0
    if name == "two_components":
        # every frontier has run dry: further passes are no-ops everywhere
        rep = hp.expand(20)[0]
        assert rep.active_trees == 0 and rep.expansions == 0 and rep.iterations == 20
        assert rp.iterate(20) == 0 and op.iterate(20) == 0
        compare_with_oracle(hp, rp, K, False)
    if name == "two_vertices":
        assert rp.flags()["connected"] == 1
    hp.close(); op.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_cuda_edge_cases(name):
    from imomd_b200 import capi
    g, dest, chunks = SCENARIOS[name]
    K = len(dest)
    op = ol.OraclePlanner(g, dest)
    dg = capi.DeviceGraph(g)
    dp = capi.DevicePlanner(dg, [dest])
    dp.feed(capi.mt19937_words(0, 200000))
    compare_with_oracle(dp, op, K, False)
    for chunk in chunks:
        rep = dp.expand(chunk)[0]
        assert rep.status == 0 and rep.iterations == chunk and rep.expansions == op.iterate(chunk)
        compare_with_oracle(dp, op, K, False)
    dp.close(); dg.close(); op.close()
