"""Time to first solution on config 3 (one query): facade vs unmodified reference (development aid)."""
import sys, os, json, subprocess, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "imomd-rrtstar_b200")); sys.path.insert(0, ROOT)
import numpy as np
import bench
from imomd_b200 import capi
class A: grid = int(sys.argv[1]) if len(sys.argv) > 1 else 1000; queries = 1
max_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 30000
g, dests = bench.make_workload(A, 0)
dest = dests[0]
t0 = time.time()
got = capi.facade_run(g, dest, max_iter=max_iter, max_time=120)
print("b200 wall", round(time.time() - t0, 3), "iterations", got["iterations"], "rows", [(round(r[0], 4), round(r[1], 1), r[2], r[3]) for r in got["rows"][:3]], "final", got["final_cost"])
with tempfile.TemporaryDirectory() as td:
    gp = os.path.join(td, "g.bin"); g.write_probe_file(gp)
    t0 = time.time()
    out = subprocess.check_output([bench.REF_PROBE, "--graph", gp, "--dest", ",".join(map(str, dest)), "--mode", "run",
                                   "--max-iter", str(max_iter), "--max-time", "120"], text=True)
    res = json.loads(out.strip().splitlines()[-1])
print("ref wall", round(time.time() - t0, 3), "iterations", res["iterations"], "rows", [(round(r[0], 4), round(r[1], 1), r[2], r[3]) for r in res["rows"][:3]], "final", res["final_cost"])
