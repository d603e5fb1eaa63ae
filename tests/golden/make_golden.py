#!/usr/bin/env python
"""Generates the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference and oracle/_ref, i.e.
`make -C oracle ref`).  The reference ships no golden vectors for this path (SURVEY.md
section 4), so these are produced by running it: road graphs parsed by its OSM parser,
tree hashes / matrices / per-expansion traces of its expansion loop, nearest-frontier
answers on its real `expandables` containers, libstdc++ random-number pins, ECI-Gen
solves and whole anytime runs.  Toolchain recorded in every file.

    python tests/golden/make_golden.py
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "imomd-rrtstar_b200"))

import oraclelib as ol  # noqa: E402
from imomd_b200 import graphs  # noqa: E402

OSM = "/root/reference/osm_data/%s.osm"
R = ol.ref()
TOOLCHAIN = R.ref_toolchain().decode()


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def osm_graph(name):
    gh = R.ref_graph_from_osm((OSM % name).encode())
    g = ol.ref_export_graph(gh, name)
    return gh, g


def pick(gh, seed, count):
    d = np.empty(count, np.uint64)
    R.ref_pick_destinations(gh, seed, count, d)
    return [int(x) for x in d]


def snapshot(rp, K, iters, expansions):
    m, f = rp.matrices(), rp.flags()
    return dict(iters=iters, expansions=int(expansions), treehash=f"{rp.treehash():016x}",
                tree_sizes=[int(R.ref_planner_tree_size(rp.ph, k)) for k in range(K)],
                frontier_sizes=[int(len(rp.frontier(k))) for k in range(K)],
                D=[int(x) for x in bits(m["D"]).ravel()], conn=[int(x) for x in m["conn"].ravel()],
                P=[int(x) for x in bits(m["P"]).ravel()], mh_id=[int(x) for x in m["mh_id"].ravel()],
                mh_val=[int(x) for x in bits(m["mh_val"]).ravel()],
                is_done=[int(x) for x in f["is_done"]], ds_parent=[int(x) for x in f["ds_parent"]],
                connected=f["connected"], dm_updated=f["dm_updated"])


def main():
    cases = []
    # ---- road graphs from the reference's own parser -------------------------------
    handles = {}
    for name in ("FRB", "FRB2", "FRB-all", "san_pablo"):
        gh, g = osm_graph(name)
        handles[name] = (gh, g)
        g.save(os.path.join(HERE, "graphs", name + ".npz"))
    # ---- expansion checkpoints (goal_bias 1.0 only in-process, see oraclelib.RefPlanner) ----
    plan = [("san_pablo", 7, 7, [1, 99, 900, 2000], False), ("FRB", 5, 7, [1, 99, 900], False),
            ("FRB-all", 10, 7, [100, 2900], False), ("FRB2", 5, 7, [100, 900], False),
            ("san_pablo", 2, 11, [100, 2900], False), ("san_pablo", 6, 3, [100, 1400], True)]
    for name, K, seed, chunks, pseudo in plan:
        gh, g = handles[name]
        dest = pick(gh, seed, K)
        rp = ol.RefPlanner(gh, dest, 1.0, pseudo)
        cps = [snapshot(rp, K, 0, 0)]
        total = 0
        for c in chunks:
            e = rp.iterate(c)
            total += c
            cps.append(snapshot(rp, K, total, e))
        case = dict(graph=name, dest=dest, goal_bias=1.0, pseudo=pseudo, checkpoints=cps)
        if pseudo:
            nsets = K * (K - 1) // 2
            case["connection_sets"] = [sorted(rp.connection_set(i)) for i in range(nsets)]
        cases.append(case)
    # a jittered grid built here (edge list is part of the fixture: same weights everywhere)
    gg = graphs.jittered_grid(90, seed=5)
    gg.save(os.path.join(HERE, "graphs", "grid90.npz"))
    ggh = R.ref_graph_from_edges(gg.n, gg.lat, gg.lon, len(gg.eu), gg.eu, gg.ev,
                                 gg.ew.ctypes.data_as(ol.C.c_void_p))
    handles["grid90"] = (ggh, gg)
    dest = graphs.pick_destinations(gg, 12, seed=7)
    rp = ol.RefPlanner(ggh, dest)
    cps = [snapshot(rp, 12, 0, 0)]
    total = 0
    for c in (100, 900, 1500):
        e = rp.iterate(c); total += c
        cps.append(snapshot(rp, 12, total, e))
    cases.append(dict(graph="grid90", dest=dest, goal_bias=1.0, pseudo=False, checkpoints=cps))
    json.dump(dict(toolchain=TOOLCHAIN, cases=cases), open(os.path.join(HERE, "expansion.json"), "w"))

    # ---- per-expansion traces --------------------------------------------------------
    gh, g = handles["san_pablo"]
    dest = pick(gh, 7, 7)
    rp = ol.RefPlanner(gh, dest)
    rows = []
    for it in range(120):
        for k in range(7):
            t = rp.expand_traced(k)
            rows.append((t.tree, t.skipped, t.rand_id, bits([t.rand_lat])[0], bits([t.rand_lon])[0],
                         t.x_nearest, bits([t.nearest_dist])[0], t.x_new, t.parent_of_new,
                         bits([t.cost_of_new])[0], t.tree_size_after, t.frontier_after))
    np.savez_compressed(os.path.join(HERE, "traces_san_pablo.npz"), dest=np.asarray(dest, np.uint64),
                        rows=np.asarray(rows, np.uint64), toolchain=np.array(TOOLCHAIN))

    # ---- nearest frontier vertex on the reference's real containers -------------------
    rng = np.random.default_rng(0)
    nn = {}
    for name, K, iters in (("san_pablo", 7, 800), ("grid90", 12, 900)):
        gh, g = handles[name]
        dest = pick(gh, 7, K) if name != "grid90" else graphs.pick_destinations(g, K, seed=7)
        rp = ol.RefPlanner(gh, dest)
        rp.iterate(iters)
        n = 1500
        tree = rng.integers(0, K, n)
        lat = rng.uniform(g.lat.min(), g.lat.max(), n)
        lon = rng.uniform(g.lon.min(), g.lon.max(), n)
        third = n // 3
        roots = rng.integers(0, K, third)
        lat[:third] = g.lat[np.asarray(dest)[roots]]
        lon[:third] = g.lon[np.asarray(dest)[roots]]
        idx = np.empty(n, np.uint64); dist = np.empty(n); ties = np.empty(n, np.int32)
        for i in range(n):
            ok, a, d, t, _ = rp.nearest(int(tree[i]), float(lat[i]), float(lon[i]))
            idx[i], dist[i], ties[i] = a, d, t
        nn[name] = dict(dest=np.asarray(dest, np.uint64), iters=iters, tree=tree.astype(np.int32), lat=lat,
                        lon=lon, idx=idx, dist=dist, ties=ties)
    np.savez_compressed(os.path.join(HERE, "nearest.npz"), toolchain=np.array(TOOLCHAIN),
                        **{f"{k}_{f}": v for k, d in nn.items() for f, v in d.items()})

    # ---- libstdc++ random numbers and geometry ---------------------------------------
    pins = {}
    w = np.empty(3000, np.uint32)
    for seed in (0, 7):
        R.ref_mt19937_words(seed, 3000, w); pins[f"mt_{seed}"] = w.copy()
    x = np.empty(2000)
    for i, (lo, hi) in enumerate(((0.0, 1.0), (0.0, 0.5), (0.0, 3.25e-4))):
        R.ref_uniform_real(0, lo, hi, 2000, x); pins[f"real_{i}"] = x.copy(); pins[f"real_{i}_range"] = np.array([lo, hi])
    u = np.empty(2000, np.uint64)
    for i, (lo, hi) in enumerate(((0, 6808), (0, 999999), (0, 9_999_999), (0, 4), (0, 2**32 - 2))):
        R.ref_uniform_size_t(3, lo, hi, 2000, u); pins[f"size_{i}"] = u.copy(); pins[f"size_{i}_range"] = np.array([lo, hi], np.uint64)
    q = np.empty(2000, np.int32)
    for i, (lo, hi) in enumerate(((0, 1), (0, 7), (0, 4), (1, 23), (-5, 17))):
        R.ref_uniform_int(0, lo, hi, 2000, q); pins[f"int_{i}"] = q.copy(); pins[f"int_{i}_range"] = np.array([lo, hi], np.int64)
    pts = np.column_stack((rng.uniform(-80, 80, 2000), rng.uniform(-179, 179, 2000),
                           rng.uniform(-80, 80, 2000), rng.uniform(-179, 179, 2000)))
    pts[1000:, 2:] = pts[1000:, :2] + rng.uniform(-1e-3, 1e-3, (1000, 2))
    pins["geo_pts"] = pts
    pins["geo_hav"] = np.array([R.ref_haversine(*r) for r in pts])
    pins["geo_bearing"] = np.array([R.ref_bearing(*r) for r in pts])
    np.savez_compressed(os.path.join(HERE, "pins.npz"), toolchain=np.array(TOOLCHAIN), **pins)

    # ---- ECI-Gen solves (solver-owned RNG persists across solves) -----------------------
    rtsp = []
    for K, p_missing in ((4, 0.0), (7, 0.3), (12, 0.5), (22, 0.6)):
        rs = R.ref_solver_create(1, 1, 2000, 300, 5)
        r2 = np.random.default_rng(K)
        solves = []
        for _ in range(4):
            D = r2.uniform(100.0, 5000.0, (K, K)); D = (D + D.T) / 2
            miss = np.triu(r2.random((K, K)) < p_missing, 1)
            D[miss | miss.T] = np.inf
            perm = r2.permutation(K)
            for a, b in zip(perm[:-1], perm[1:]):
                if np.isinf(D[a, b]):
                    D[a, b] = D[b, a] = r2.uniform(100.0, 5000.0)
            np.fill_diagonal(D, 0.0)
            D = np.ascontiguousarray(D)
            c = ol.C.c_double(); s = np.empty(256, np.int32)
            n = R.ref_solver_solve(rs, K, D.ravel(), 0, K - 1, ol.C.byref(c), s, 256)
            solves.append(dict(D=[int(v) for v in bits(D).ravel()], cost=int(bits([c.value])[0]),
                               seq=[int(v) for v in s[:n]]))
        rtsp.append(dict(K=K, ga=[2000, 300, 5], solves=solves))
    K = 22
    D = np.random.default_rng(9).uniform(10, 9000, (K, K)); D = np.ascontiguousarray((D + D.T) / 2)
    lens = rng.integers(2, 2 * K, 400).astype(np.int32)
    tours = rng.integers(0, K, (400, 64)).astype(np.int32)
    out = np.empty(400)
    R.ref_path_costs(K, D.ravel(), 400, 64, tours.ravel(), lens, out)
    np.savez_compressed(os.path.join(HERE, "path_costs.npz"), D=D, tours=tours, lens=lens, cost=out,
                        toolchain=np.array(TOOLCHAIN))
    json.dump(dict(toolchain=TOOLCHAIN, solvers=rtsp), open(os.path.join(HERE, "rtsp.json"), "w"))

    # ---- whole anytime runs (own process each: static distributions upstream) ----------
    runs = []
    for name, K, seed, max_iter in (("FRB", 5, 13, 1500), ("san_pablo", 7, 7, 3000), ("grid90", 6, 3, 2500)):
        gh, g = handles[name]
        if name == "grid90":
            dest = graphs.pick_destinations(g, K, seed=seed)
        else:
            lab = np.empty(g.n, np.uint32)
            R.ref_graph_components(gh, lab)
            big = np.bincount(lab).argmax()
            cand = pick(gh, seed, 200)
            dest = [v for v in cand if lab[v] == big][:K]
        eu, ev, ew = g.edge_list()
        gtmp = graphs.RoadGraph(lat=g.lat, lon=g.lon, row_ptr=g.row_ptr, col=g.col, w=g.w,
                                eu=np.asarray(eu, np.uint64), ev=np.asarray(ev, np.uint64), ew=ew)
        with tempfile.TemporaryDirectory() as td:
            if name == "grid90":
                gp = os.path.join(td, "g.bin"); gtmp.write_probe_file(gp)
                args = ["--graph", gp]
            else:
                args = ["--osm", OSM % name]
            out = subprocess.check_output([ol.REF_PROBE] + args + ["--dest", ",".join(map(str, dest)),
                                          "--mode", "run", "--max-iter", str(max_iter), "--max-time", "600",
                                          "--ga", "3000,400,5"], text=True)
        res = json.loads(out.strip().splitlines()[-1])
        runs.append(dict(graph=name, dest=dest, max_iter=max_iter, ga=[3000, 400, 5],
                         iterations=res["iterations"], expansions=res["expansions"],
                         rows=[[int(bits([r[1]])[0]), r[2], r[3]] for r in res["rows"]],
                         final_cost=int(bits([res["final_cost"]])[0]), path=res["path"],
                         treehash=res["treehash"]))
    json.dump(dict(toolchain=TOOLCHAIN, runs=runs), open(os.path.join(HERE, "runs.json"), "w"))
    print("golden vectors written with", TOOLCHAIN)


if __name__ == "__main__":
    main()
