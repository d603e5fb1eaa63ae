import sys
sys.path.insert(0, "/root/repo/imomd-rrtstar_b200"); sys.path.insert(0, "/root/repo")
import bench
from imomd_b200 import capi
class A: grid = 1000; queries = 296; workload = "grid"
g, dests = bench.make_workload(A, 0)
dg = capi.DeviceGraph(g); dp = capi.DevicePlanner(dg, dests)
dp.feed(capi.mt19937_words(0, 3000 * 12 * 8 + 4096))
dp.expand(3000)
c = dp.counters()
import numpy as np
print("per query mean counters [cells bounded, records ranked (nn), rewire calls, adj slots serial, reparentings, propagated, conn-checked, records ranked (rescans)]")
print((c.mean(0) / 36000).round(2), "per expansion")
for k in range(3):
    print("frontier size tree", k, len(dp.frontier(k, 0)))
