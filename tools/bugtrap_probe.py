"""BASELINE configs[1] stand-in: bug-trap grid, pseudo mode, 5 pseudo destinations, goal_bias 0.3.
Facade vs unmodified reference: time to first solution, rows, final cost (development aid)."""
import sys, os, json, subprocess, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "imomd-rrtstar_b200")); sys.path.insert(0, ROOT)
import bench
from imomd_b200 import capi, graphs
W = int(sys.argv[1]) if len(sys.argv) > 1 else 1131
max_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
max_time = int(sys.argv[3]) if len(sys.argv) > 3 else 120
g, s, t, ps = graphs.bugtrap_grid(W, seed=123)
dest = [s] + ps + [t]
fmt = lambda rows: [(round(r[0], 4), round(r[1], 1), r[2], r[3]) for r in rows]
t0 = time.time()
got = capi.facade_run(g, dest, goal_bias=0.3, max_iter=max_iter, max_time=max_time, pseudo=True)
print("b200 wall", round(time.time() - t0, 3), "iterations", got["iterations"], "expansions", got["expansions"],
      "n_rows", len(got["rows"]), "first", fmt(got["rows"][:2]), "last", fmt(got["rows"][-1:]), "final", got["final_cost"])
with tempfile.TemporaryDirectory() as td:
    gp = os.path.join(td, "g.bin"); g.write_probe_file(gp)
    t0 = time.time()
    out = subprocess.check_output([bench.REF_PROBE, "--graph", gp, "--dest", ",".join(map(str, dest)), "--mode", "run",
                                   "--pseudo", "--goal-bias", "0.3", "--max-iter", str(max_iter), "--max-time", str(max_time)], text=True)
    res = json.loads(out.strip().splitlines()[-1])
print("ref  wall", round(time.time() - t0, 3), "iterations", res["iterations"], "expansions", res["expansions"],
      "n_rows", len(res["rows"]), "first", fmt(res["rows"][:2]), "last", fmt(res["rows"][-1:]), "final", res["final_cost"])
same = [(r[3], r[2], r[1]) for r in got["rows"]] == [(r[3], r[2], r[1]) for r in res["rows"]]
print("rows identical:", same, "paths identical:", list(got["path"]) == res["path"])
