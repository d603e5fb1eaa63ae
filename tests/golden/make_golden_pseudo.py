#!/usr/bin/env python
"""Golden vectors for pseudo mode (tests/golden/runs_pseudo.json) from the UNMODIFIED reference:
whole anytime runs with `pseudo_mode` on (the trees of the pseudo destinations are merged into the
source / target trees, src/imomd_rrt_star.cpp:744-916 upstream), one reference process per run.
Authoring container only (needs /root/reference and oracle/_ref).

    python tests/golden/make_golden_pseudo.py
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
import goldenlib as gl  # noqa: E402
from imomd_b200 import graphs  # noqa: E402

OSM = "/root/reference/osm_data/%s.osm"
R = ol.ref()


def bits(x):
    return int(np.array([x], np.float64).view(np.uint64)[0])


def main():
    runs = []
    plan = [("san_pablo", "osm", 6, 5, 1.0, 3000), ("FRB", "osm", 4, 9, 1.0, 1500),
            ("grid90", "grid", 6, 3, 1.0, 2500), ("bugtrap80", "bugtrap", 7, 0, 0.3, 4000)]
    for name, kind, K, seed, goal_bias, max_iter in plan:
        with tempfile.TemporaryDirectory() as td:
            if kind == "osm":
                gh = R.ref_graph_from_osm((OSM % name).encode())
                g = ol.ref_export_graph(gh, name)
                lab = np.empty(g.n, np.uint32)
                R.ref_graph_components(gh, lab)
                big = np.bincount(lab).argmax()
                cand = np.empty(200, np.uint64)
                R.ref_pick_destinations(gh, seed, 200, cand)
                dest = [int(v) for v in cand if lab[int(v)] == big][:K]
                args = ["--osm", OSM % name]
            elif kind == "grid":
                g = gl.graph(name)
                dest = graphs.pick_destinations(g, K, seed=seed)
                gp = os.path.join(td, "g.bin"); g.write_probe_file(gp)
                args = ["--graph", gp]
            else:
                g, s, t, ps = graphs.bugtrap_grid(80, seed=3)
                dest = [s] + ps + [t]
                gp = os.path.join(td, "g.bin"); g.write_probe_file(gp)
                args = ["--graph", gp]
            out = subprocess.check_output(
                [ol.REF_PROBE] + args + ["--dest", ",".join(map(str, dest)), "--mode", "run", "--pseudo",
                                         "--goal-bias", repr(goal_bias), "--max-iter", str(max_iter),
                                         "--max-time", "600", "--ga", "3000,400,5"], text=True)
        res = json.loads(out.strip().splitlines()[-1])
        assert res["rows"], f"{name}: the reference found no path"
        runs.append(dict(graph=name, kind=kind, dest=[int(d) for d in dest], goal_bias=goal_bias,
                         max_iter=max_iter, ga=[3000, 400, 5], iterations=res["iterations"],
                         expansions=res["expansions"],
                         rows=[[bits(r[1]), r[2], r[3]] for r in res["rows"]],
                         final_cost=bits(res["final_cost"]), path=res["path"]))
        print(name, "rows", len(res["rows"]), "final", res["final_cost"], "path", len(res["path"]))
    json.dump(dict(toolchain=R.ref_toolchain().decode(), runs=runs),
              open(os.path.join(HERE, "runs_pseudo.json"), "w"))


if __name__ == "__main__":
    main()
