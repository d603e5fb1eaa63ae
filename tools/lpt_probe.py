"""Would longest-first scheduling of a multi-wave batch pay?  (development aid)
Runs the san_pablo batch in chunks of passes, records every query's busy cycles per chunk and simulates
the makespan on S slots for: one launch in natural order, one launch in the (unknowable) ideal longest-first
order, and chunked launches ordered by the previous chunk's cost."""
import sys, os, heapq
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "imomd-rrtstar_b200")); sys.path.insert(0, ROOT)
import numpy as np
import bench
from imomd_b200 import capi

def busy(dp, Q):
    out = np.empty(Q)
    for q in range(Q):
        pr = dp.profile(q); pr.pop("bfs_vertices_levels")
        out[q] = sum(c for c, _ in pr.values())
    return out

def makespan(costs, order, slots):
    h = [0.0] * slots
    heapq.heapify(h)
    for q in order:
        t = heapq.heappop(h)
        heapq.heappush(h, t + costs[q])
    return max(h)

class A: workload = sys.argv[1] if len(sys.argv) > 1 else "san_pablo"; queries = int(sys.argv[2]) if len(sys.argv) > 2 else 1024; grid = 1000
chunks = int(sys.argv[3]) if len(sys.argv) > 3 else 6
iters = 3000
g, dests = bench.make_workload(A, 0)
K = len(dests[0])
dg = capi.DeviceGraph(g); dp = capi.DevicePlanner(dg, dests); dp.set_profiling(True)
dp.feed(capi.mt19937_words(0, iters * K * 8 + 4096))
per = []; prev = np.zeros(A.queries); wall = 0.0
for c in range(chunks):
    dp.expand(iters // chunks); wall += dp.last_kernel_ms()
    cur = busy(dp, A.queries); per.append(cur - prev); prev = cur
per = np.array(per); total = per.sum(0); slots = 296; nat = np.arange(A.queries)
ghz = 1.965e6   # cycles per ms
print(f"measured: {chunks} launches, {wall:.0f} ms in total")
print(f"one launch, natural order : {makespan(total, nat, slots) / ghz:.0f} ms (simulated)")
print(f"one launch, ideal LPT     : {makespan(total, np.argsort(-total), slots) / ghz:.0f} ms")
t = makespan(per[0], nat, slots)
for c in range(1, chunks):
    t += makespan(per[c], np.argsort(-per[c - 1]), slots)
print(f"{chunks} launches, LPT by the previous chunk: {t / ghz:.0f} ms")
print(f"lower bounds: work / slots {total.sum() / slots / ghz:.0f} ms, slowest query {total.max() / ghz:.0f} ms")
