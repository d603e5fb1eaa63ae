"""How local would cost propagation be if tree records were laid out in visit order? (development aid)
For every tree edge parent->child of one bench query: same 128-byte line (4 records) / same 32-record KB."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "imomd-rrtstar_b200")); sys.path.insert(0, ROOT)
import numpy as np
import bench
from imomd_b200 import capi
class A: grid = 1000; queries = 1; workload = "grid"
g, dests = bench.make_workload(A, 0)
dg = capi.DeviceGraph(g); dp = capi.DevicePlanner(dg, dests, pseudo=True)
dp.feed(capi.mt19937_words(0, 3000 * 12 * 8 + 4096))
dp.expand(3000)
tot = same4 = same32 = chain = chain4 = id4 = 0
for k in range(12):
    par, cost = dp.tree(k)
    order = dp.visit_order(k)
    idx = np.full(g.n, -1, np.int64); idx[order] = np.arange(len(order))
    vis = np.nonzero(par != 0xFFFFFFFF)[0]
    vis = vis[par[vis] != vis]                      # not the root
    p = par[vis].astype(np.int64)
    nchild = np.bincount(p, minlength=g.n)
    tot += len(vis)
    same4 += int(((idx[vis] >> 2) == (idx[p] >> 2)).sum())
    same32 += int(((idx[vis] >> 5) == (idx[p] >> 5)).sum())
    id4 += int(((vis >> 2) == (p >> 2)).sum())
    single = nchild[p] == 1
    chain += int(single.sum()); chain4 += int((((idx[vis] >> 2) == (idx[p] >> 2)) & single).sum())
print(f"tree edges {tot}: child in the parent's 128 B line -- visit-order layout {same4/tot:.2%}, id-order layout {id4/tot:.2%}; "
      f"same 1 KB (visit order) {same32/tot:.2%}; single-child links {chain} of which same line {chain4/max(chain,1):.2%}")
