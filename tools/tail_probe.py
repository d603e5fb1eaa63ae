"""Per-query busy cycles of the bench workload of a given rank: where the batch's tail comes from (development aid)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "imomd-rrtstar_b200")); sys.path.insert(0, ROOT)
import numpy as np
import bench
from imomd_b200 import capi
class A: grid = 1000; queries = 296; workload = "grid"
g = None
for rank in [int(a) for a in sys.argv[1:]] or [0]:
    g, dests = bench.make_workload(A, rank)
    dg = capi.DeviceGraph(g); dp = capi.DevicePlanner(dg, dests); dp.set_profiling(True)
    dp.feed(capi.mt19937_words(0, 3000 * 12 * 8 + 4096))
    reps = dp.expand(3000)
    ms = dp.last_kernel_ms()
    tot, bfs = [], []
    for q in range(A.queries):
        pr = dp.profile(q); bv = pr.pop("bfs_vertices_levels")
        tot.append(sum(c for c, _ in pr.values())); bfs.append(bv)
    tot = np.array(tot, float); order = np.argsort(-tot)
    print(f"rank {rank}: kernel {ms:.0f} ms  busy cycles mean {tot.mean():.3g} p50 {np.median(tot):.3g} p90 {np.percentile(tot,90):.3g} "
          f"p99 {np.percentile(tot,99):.3g} max {tot.max():.3g} (max/mean {tot.max()/tot.mean():.2f})")
    for q in order[:4]:
        print(f"   q{q}: cycles {tot[q]:.3g} bfs vertices/levels {bfs[q]} expansions {reps[q].expansions} tree_vertices {reps[q].tree_vertices}")
    dp.close(); dg.close()
