"""BASELINE configs[3] shape on ONE GPU: synthetic 10 M-vertex road-like graph, 20 destinations (K = 22 trees);
expansion throughput for 1 and 8 queries, the single query checked against the oracle (development aid)."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "imomd-rrtstar_b200"))
import numpy as np
from imomd_b200 import capi, graphs

n_target = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 500
t0 = time.time()
g = graphs.road_like(n_target, seed=11)
print(f"graph: {g.n} vertices, {len(g.col)} arcs, max degree {g.degree().max()}, degree-2 share {(g.degree() == 2).mean():.2f}, built in {time.time() - t0:.1f} s", flush=True)
K = 22
for Q in (1, 8):
    dests = [graphs.pick_destinations(g, K, seed=100 + q) for q in range(Q)]
    t0 = time.time()
    dg = capi.DeviceGraph(g); dp = capi.DevicePlanner(dg, dests)
    t_create = time.time() - t0
    dp.feed(capi.mt19937_words(0, iters * K * 8 + 4096))
    reps = dp.expand(iters)
    ms = dp.last_kernel_ms()
    ex = sum(r.expansions for r in reps)
    print(json.dumps(dict(Q=Q, K=K, iters=iters, create_s=round(t_create, 2), kernel_ms=round(ms, 1), expansions=ex,
                          exp_per_s=round(ex / (ms / 1e3)), tree_vertices=sum(r.tree_vertices for r in reps),
                          status=[r.status for r in reps], hash_cells=dg.info()["grid_dim"])), flush=True)
    if Q == 1:
        # the same prefix on the CPU restatement (oracle/, pinned against the reference): trees,
        # frontiers, matrices and flags of all 22 trees must agree
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oraclelib as ol
        from parity import compare_with_oracle
        t0 = time.time()
        op = ol.OraclePlanner(g, dests[0])
        assert op.iterate(iters) == ex
        compare_with_oracle(dp, op, K, False, check_frontier=False)
        print(f"parity with the oracle after {iters} passes on {g.n} vertices: OK ({time.time() - t0:.1f} s of CPU)", flush=True)
        op.close()
    dp.close(); dg.close()
