"""One k_expand launch of the bench workload (296 queries x `iters` passes): the target of profiler captures."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "imomd-rrtstar_b200")); sys.path.insert(0, ROOT)
import bench
from imomd_b200 import capi
class A: grid = 1000; queries = int(sys.argv[1]) if len(sys.argv) > 1 else 296; workload = "grid"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
g, dests = bench.make_workload(A, 0)
dg = capi.DeviceGraph(g); dp = capi.DevicePlanner(dg, dests)
dp.feed(capi.mt19937_words(0, iters * 12 * 8 + 4096))
reps = dp.expand(iters)
print("kernel_ms", dp.last_kernel_ms(), "expansions", sum(r.expansions for r in reps))
