"""Where do the `unresolved_ties` of the bench workload come from? (development aid)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "imomd-rrtstar_b200")); sys.path.insert(0, ROOT)
import numpy as np
import bench
from imomd_b200 import capi
class A: grid=1000; queries=int(sys.argv[1]) if len(sys.argv)>1 else 148
g, dests = bench.make_workload(A, 0)
dg = capi.DeviceGraph(g); dp = capi.DevicePlanner(dg, dests)
dp.feed(capi.mt19937_words(0, 3000*12*8+4096))
reps = dp.expand(3000)
for q, r in enumerate(reps):
    if r.unresolved_ties:
        ti = dp.tie_info(q)
        print(q, r.unresolved_ties, r.near_ties, ti.tolist(), "gap", ti[4]-ti[3], "dest", dests[q][int(ti[1])], dests[q][int(ti[2])] if ti[2]>=0 else None)
print("total unresolved", sum(r.unresolved_ties for r in reps), "near", sum(r.near_ties for r in reps))
