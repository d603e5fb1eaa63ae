"""The CUDA path (through the C ABI and through the C++ facade) against the committed golden
vectors of the unmodified reference: OSM road graphs parsed by the reference's parser,
checkpoints of its expansion loop, its nearest-frontier answers, its whole anytime runs."""
import numpy as np
import pytest

import goldenlib as gl
from imomd_b200 import capi, graphs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", gl.expansion_cases(), ids=lambda c: f"{c['graph']}-K{len(c['dest'])}{'-pseudo' if c['pseudo'] else ''}")
def test_expansion_checkpoints(case):
    g = gl.graph(case["graph"])
    K = len(case["dest"])
    dg = capi.DeviceGraph(g)
    dp = capi.DevicePlanner(dg, [case["dest"]], None, case["goal_bias"], case["pseudo"])
    dp.feed(capi.mt19937_words(0, 400000))
    done = 0
    for cp in case["checkpoints"]:
        if cp["iters"] > done:
            rep = dp.expand(cp["iters"] - done)[0]
            assert rep.status == 0 and rep.expansions == cp["expansions"]
            done = cp["iters"]
        trees = [dp.tree(k) for k in range(K)]
        st = dp.state()
        st["is_done"] = [int(x) for x in st["is_done"]]
        st["ds_parent"] = [int(x) for x in st["ds_parent"]]
        gl.check_checkpoint(st, trees, cp, K, False)
        assert [len(dp.frontier(k)) for k in range(K)] == cp["frontier_sizes"]
    if case["pseudo"]:
        si, vx = dp.connection_log()
        for idx, want in enumerate(case["connection_sets"]):
            assert sorted(set(int(v) for v in vx[si == idx])) == want
    dp.close(); dg.close()


@pytest.mark.parametrize("name", ["san_pablo", "grid90"])
def test_nearest_golden(name):
    z = np.load(gl.GOLDEN + "/nearest.npz")
    g = gl.graph(name)
    dest = [int(x) for x in z[f"{name}_dest"]]
    dg = capi.DeviceGraph(g); dp = capi.DevicePlanner(dg, [dest])
    dp.feed(capi.mt19937_words(0, 200000))
    dp.expand(int(z[f"{name}_iters"]))
    n = len(z[f"{name}_tree"])
    lk = np.empty(n, capi.NN_DTYPE)
    lk["query"] = 0; lk["tree"] = z[f"{name}_tree"]; lk["lat"] = z[f"{name}_lat"]; lk["lon"] = z[f"{name}_lon"]
    idx, dist = dp.nearest(lk)
    assert np.array_equal(idx.astype(np.uint64), z[f"{name}_idx"])          # bit-exact indices
    want = z[f"{name}_dist"]
    assert np.all(np.abs(dist - want) <= 1e-12 * np.maximum(want, 1.0))
    dp.close(); dg.close()


@pytest.mark.parametrize("run", gl.runs(), ids=lambda r: r["graph"])
def test_facade_anytime_runs(run):
    g = gl.graph(run["graph"])
    got = capi.facade_run(g, run["dest"], max_iter=run["max_iter"], max_time=600, ga=tuple(run["ga"]))
    assert got["iterations"] == run["iterations"] and got["expansions"] == run["expansions"]
    rows = [[int(np.array([r[1]]).view(np.uint64)[0]), r[2], r[3]] for r in got["rows"]]
    assert rows == run["rows"]
    assert int(np.array([got["final_cost"]]).view(np.uint64)[0]) == run["final_cost"]
    assert [int(v) for v in got["path"]] == run["path"]


@pytest.mark.parametrize("run", gl.runs_pseudo(), ids=lambda r: r["graph"])
def test_facade_pseudo_mode_runs(run):
    """pseudo_mode runs (tree merge, src/imomd_rrt_star.cpp:744-916 upstream) recorded from the
    unmodified reference by tests/golden/make_golden_pseudo.py: every logged row (cost bits, tree
    size, iteration), the final cost and the printed path."""
    if run["kind"] == "bugtrap":
        g = graphs.bugtrap_grid(80, seed=3)[0]
    else:
        g = gl.graph(run["graph"])
    got = capi.facade_run(g, run["dest"], goal_bias=run["goal_bias"], max_iter=run["max_iter"],
                          max_time=600, ga=tuple(run["ga"]), pseudo=True)
    assert got["iterations"] == run["iterations"] and got["expansions"] == run["expansions"]
    rows = [[int(np.array([r[1]]).view(np.uint64)[0]), r[2], r[3]] for r in got["rows"]]
    assert rows == run["rows"]
    assert int(np.array([got["final_cost"]]).view(np.uint64)[0]) == run["final_cost"]
    assert [int(v) for v in got["path"]] == run["path"]


def test_ga_eval_golden():
    z = np.load(gl.GOLDEN + "/path_costs.npz")
    ctx = capi.GaContext(0)
    got = ctx.eval(z["D"], z["tours"], z["lens"])
    assert np.array_equal(got.view(np.uint64), z["cost"].view(np.uint64))
    ctx.close()
