"""Per-query busy cycles of one bench step (streaming planner) with the queries' destinations, for
offline analysis of what makes a query slow (development aid).  Writes gpurun_out/<name>.npz."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "imomd-rrtstar_b200")); sys.path.insert(0, ROOT)
import numpy as np
import bench
from imomd_b200 import capi
class A: grid = 1000; queries = int(sys.argv[1]) if len(sys.argv) > 1 else 2368; workload = "grid"
slots = int(sys.argv[2]) if len(sys.argv) > 2 else 296
name = sys.argv[3] if len(sys.argv) > 3 else "r02_busy"
g, dests = bench.make_workload(A, 0)
dg = capi.DeviceGraph(g); dp = capi.DevicePlanner(dg, dests, max_resident=slots)
dp.feed(capi.mt19937_words(0, 3000 * 12 * 8 + 4096))
reps = dp.expand(3000)
print("kernel_ms", dp.last_kernel_ms())
np.savez_compressed(os.path.join(ROOT, "gpurun_out", name + ".npz"), dests=np.array(dests, np.int64),
                    busy=np.array([r.busy_cycles for r in reps], np.int64),
                    expansions=np.array([r.expansions for r in reps], np.int64),
                    tree_vertices=np.array([r.tree_vertices for r in reps], np.int64),
                    lat=g.lat, lon=g.lon, kernel_ms=dp.last_kernel_ms())
