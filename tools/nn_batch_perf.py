"""Bandwidth of the brute-force tile kernel behind imomd_nearest_batch on the bench state
(296 queries x 12 trees after `iters` passes on the 1 M-vertex grid): development aid + the
`roofline_nn` figure of bench.py."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "imomd-rrtstar_b200")); sys.path.insert(0, ROOT)
import numpy as np
import bench
from imomd_b200 import capi

def measure(dp, g, Q, K, per_tree, reps=5, seed=0):
    rng = np.random.default_rng(seed)
    n = Q * K * per_tree
    lk = np.empty(n, capi.NN_DTYPE)
    lk["query"] = np.repeat(np.arange(Q), K * per_tree)
    lk["tree"] = np.tile(np.repeat(np.arange(K), per_tree), Q)
    lk["lat"] = rng.uniform(g.lat.min(), g.lat.max(), n); lk["lon"] = rng.uniform(g.lon.min(), g.lon.max(), n)
    best = None
    call_ms = 1e30
    idx_exact, _ = dp.nearest(lk[:2000])
    for _ in range(reps):
        idx, dist, info = capi.nearest_batch(dp, lk)
        assert np.array_equal(idx[:2000], idx_exact), "tile kernel differs from the pruned exact path"
        call_ms = min(call_ms, info["call_ms"])
        if best is None or info["kernel_ms"] < best["kernel_ms"]:
            best = info
    gbs = best["bytes"] / (best["kernel_ms"] / 1e3) / 1e9
    return dict(lookups=n, per_tree=per_tree, kernel_ms=round(best["kernel_ms"], 4), bytes=best["bytes"],
                gb_per_s=round(gbs, 1), refined=best["refined"], lookups_per_s=round(n / (best["kernel_ms"] / 1e3)),
                prep_ms=round(best["prep_ms"], 4), merge_ms=round(best["merge_ms"], 4), segments=best["segments"],
                call_ms=round(call_ms, 3), lookups_per_s_call=round(n / (call_ms / 1e3)))

if __name__ == "__main__":
    class A: grid = 1000; queries = int(sys.argv[1]) if len(sys.argv) > 1 else 296
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
    g, dests = bench.make_workload(A, 0)
    dg = capi.DeviceGraph(g); dp = capi.DevicePlanner(dg, dests)
    dp.feed(capi.mt19937_words(0, iters * 12 * 8 + 4096))
    reps = dp.expand(iters)
    fr = sum(len(dp.frontier(k, 0)) for k in range(12))
    print("frontier records of query 0:", fr)
    for per_tree in (1, 4, 8, 16):
        print(json.dumps(measure(dp, g, A.queries, 12, per_tree)), flush=True)
