"""Where the 800 passes before the bug-trap merge go (development aid): the per-request cycle counters of the
single-query latency build on the bug-trap grid, pseudo mode, goal_bias 0.3."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "imomd-rrtstar_b200")); sys.path.insert(0, ROOT)
from imomd_b200 import capi, graphs
W = int(sys.argv[1]) if len(sys.argv) > 1 else 1131
g, s, t, ps = graphs.bugtrap_grid(W, seed=123)
dest = [s] + ps + [t]
dg = capi.DeviceGraph(g)
dp = capi.DevicePlanner(dg, [dest], goal_bias=0.3, pseudo=True)
dp.set_profiling(True)
dp.feed(capi.mt19937_words(0, 900 * len(dest) * 8 + 4096))
tot = 0.0
for i in range(8):
    rep = dp.expand(100)[0]
    tot += dp.last_kernel_ms()
    pr = dp.profile(0); bv = pr.pop("bfs_vertices_levels")
    cyc = {k: round(v[0] / 1.965e6, 2) for k, v in pr.items()}
    cnt = {k: v[1] for k, v in pr.items()}
    print(f"after {100 * (i + 1)} passes: kernel {dp.last_kernel_ms():.2f} ms (sum {tot:.1f}); cumulative ms by work {cyc}; calls {cnt}; "
          f"bfs vertices/levels {bv}; expansions {rep.expansions} tree {rep.tree_vertices}", flush=True)
pc, pn = dp.phase_cycles(0)
print("serial steps by phase (ms):", [round(float(c) / 1.965e6, 2) for c in pc[:8]], "counts", [int(x) for x in pn[:8]])
