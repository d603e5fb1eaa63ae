"""The oracle restatement against the committed golden vectors of the unmodified reference
(tests/golden/, see make_golden.py).  Needs neither a GPU nor the reference checkout: this is
the pin that travels."""
import numpy as np
import pytest

import goldenlib as gl
import oraclelib as ol


def _oracle_state(op):
    m, f = op.matrices(), op.flags()
    return dict(D=m["D"], conn=m["conn"], P=m["P"], mh_id=m["mh_id"], mh_val=m["mh_val"],
                is_done=[int(x) for x in f["is_done"]], ds_parent=[int(x) for x in f["ds_parent"]],
                connected=f["connected"], dm_updated=f["dm_updated"])


@pytest.mark.parametrize("case", gl.expansion_cases(), ids=lambda c: f"{c['graph']}-K{len(c['dest'])}{'-pseudo' if c['pseudo'] else ''}")
def test_expansion_checkpoints(case):
    g = gl.graph(case["graph"])
    K = len(case["dest"])
    op = ol.OraclePlanner(g, case["dest"], case["goal_bias"], case["pseudo"])
    done = 0
    for cp in case["checkpoints"]:
        if cp["iters"] > done:
            e = op.iterate(cp["iters"] - done)
            assert e == cp["expansions"]
            done = cp["iters"]
        trees = [op.tree(k) for k in range(K)]
        gl.check_checkpoint(_oracle_state(op), trees, cp, K, True)
        assert [len(op.frontier(k)) for k in range(K)] == cp["frontier_sizes"]
    if case["pseudo"]:
        si, vx = op.connection_log()
        for idx, want in enumerate(case["connection_sets"]):
            assert sorted(set(int(v) for v in vx[si == idx])) == want
    assert op.ties() == 0
    op.close()


def test_traces():
    z = np.load(gl.GOLDEN + "/traces_san_pablo.npz")
    g = gl.graph("san_pablo")
    dest = [int(x) for x in z["dest"]]
    op = ol.OraclePlanner(g, dest)
    rows = z["rows"]
    i = 0
    for it in range(120):
        for k in range(7):
            t = op.expand_traced(k)
            got = (t.tree, t.skipped, t.rand_id, *np.array([t.rand_lat, t.rand_lon]).view(np.uint64),
                   t.x_nearest, np.array([t.nearest_dist]).view(np.uint64)[0], t.x_new, t.parent_of_new,
                   np.array([t.cost_of_new]).view(np.uint64)[0], t.tree_size_after, t.frontier_after)
            assert tuple(int(x) for x in got) == tuple(int(x) for x in rows[i]), (it, k)
            i += 1
    op.close()


@pytest.mark.parametrize("name,K", [("san_pablo", 7), ("grid90", 12)])
def test_nearest(name, K):
    z = np.load(gl.GOLDEN + "/nearest.npz")
    g = gl.graph(name)
    dest = [int(x) for x in z[f"{name}_dest"]]
    op = ol.OraclePlanner(g, dest)
    op.iterate(int(z[f"{name}_iters"]))
    tree, lat, lon = z[f"{name}_tree"], z[f"{name}_lat"], z[f"{name}_lon"]
    assert (z[f"{name}_ties"] == 1).all()          # no exact ties in the fixtures
    for i in range(len(tree)):
        ok, idx, d, ties, _ = op.nearest(int(tree[i]), float(lat[i]), float(lon[i]))
        assert idx == int(z[f"{name}_idx"][i]) and d == z[f"{name}_dist"][i]
    op.close()


def test_pins():
    z = np.load(gl.GOLDEN + "/pins.npz")
    O = ol.oracle()
    w = np.empty(3000, np.uint32)
    for seed in (0, 7):
        O.orc_mt19937_words(seed, 3000, w)
        assert np.array_equal(w, z[f"mt_{seed}"])
    x = np.empty(2000)
    for i in range(3):
        lo, hi = z[f"real_{i}_range"]
        O.orc_uniform_real(0, lo, hi, 2000, x)
        assert np.array_equal(x.view(np.uint64), z[f"real_{i}"].view(np.uint64))
    u = np.empty(2000, np.uint64)
    for i in range(5):
        lo, hi = (int(v) for v in z[f"size_{i}_range"])
        O.orc_uniform_size_t(3, lo, hi, 2000, u)
        assert np.array_equal(u, z[f"size_{i}"])
    q = np.empty(2000, np.int32)
    for i in range(5):
        lo, hi = (int(v) for v in z[f"int_{i}_range"])
        O.orc_uniform_int(0, lo, hi, 2000, q)
        assert np.array_equal(q, z[f"int_{i}"])
    pts = z["geo_pts"]
    hav = np.array([O.orc_haversine(*r) for r in pts])
    brg = np.array([O.orc_bearing(*r) for r in pts])
    assert np.array_equal(hav.view(np.uint64), z["geo_hav"].view(np.uint64))
    assert np.array_equal(brg.view(np.uint64), z["geo_bearing"].view(np.uint64))


def test_rtsp_and_path_costs():
    O = ol.oracle()
    for s in gl.rtsp():
        K = s["K"]
        h = O.orc_solver_create(1, 1, *s["ga"])
        for sol in s["solves"]:
            D = np.ascontiguousarray(gl.f64(sol["D"]))
            c = ol.C.c_double(); seq = np.empty(256, np.int32)
            n = O.orc_solver_solve(h, K, D, 0, K - 1, ol.C.byref(c), seq, 256)
            assert [int(v) for v in seq[:n]] == sol["seq"]
            assert np.array([c.value]).view(np.uint64)[0] == sol["cost"]
        O.orc_solver_destroy(h)
    z = np.load(gl.GOLDEN + "/path_costs.npz")
    out = np.empty(len(z["lens"]))
    O.orc_path_costs(z["D"].shape[0], np.ascontiguousarray(z["D"]).ravel(), len(out), z["tours"].shape[1],
                     np.ascontiguousarray(z["tours"]).ravel(), np.ascontiguousarray(z["lens"]), out)
    assert np.array_equal(out.view(np.uint64), z["cost"].view(np.uint64))
