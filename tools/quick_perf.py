"""Ad-hoc device timing of the expansion kernel (development aid, not the bench)."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "imomd-rrtstar_b200")); sys.path.insert(0, ROOT)
import numpy as np
from imomd_b200 import capi, graphs

def run(name, g, dests, iters_list, **kw):
    dg = capi.DeviceGraph(g)
    t0 = time.time()
    dp = capi.DevicePlanner(dg, dests, **kw)
    dp.set_profiling(True)
    t_create = time.time() - t0
    dp.feed(capi.mt19937_words(0, 3_000_000))
    out = []
    for it in iters_list:
        t0 = time.time()
        reps = dp.expand(it)
        wall = time.time() - t0
        ms = dp.last_kernel_ms()
        ex = sum(r.expansions for r in reps)
        out.append(dict(iters=it, expansions=ex, kernel_ms=round(ms, 3), wall_ms=round(wall * 1e3, 3),
                        exp_per_s=round(ex / (ms / 1e3), 1), status=[r.status for r in reps][:4],
                        tree_vertices=sum(r.tree_vertices for r in reps),
                        near=sum(r.near_ties for r in reps), unres=sum(r.unresolved_ties for r in reps)))
    if len(dests) > 1:
        tot = []
        for q in range(len(dests)):
            pr = dp.profile(q); pr.pop("bfs_vertices_levels")
            tot.append(sum(c for c, _ in pr.values()))
        tot = np.array(tot, float)
        qs = int(np.argmax(tot))
        prs = dp.profile(qs); bvs = prs.pop("bfs_vertices_levels"); ts = sum(c for c, _ in prs.values()) or 1
        print(f"  slowest query {qs}: expansions {reps[qs].expansions} tree_vertices {reps[qs].tree_vertices} bfs vertices/levels {bvs}",
              {k: f"{100*c/ts:.1f}% n={n} {c/max(n,1):.0f}cyc" for k, (c, n) in prs.items()})
        print("  per-query busy cycles: min %.3g  p50 %.3g  p90 %.3g  max %.3g  mean %.3g  (max/mean %.2f)" % (
            tot.min(), np.median(tot), np.percentile(tot, 90), tot.max(), tot.mean(), tot.max() / tot.mean()))
    prof = dp.profile(0)
    print("  bfs vertices, levels:", prof.pop("bfs_vertices_levels"))
    tot = sum(c for c, _ in prof.values()) or 1
    print("  profile q0:", {k: f"{100*c/tot:.1f}% n={n} {c/max(n,1):.0f}cyc" for k, (c, n) in prof.items()})
    print(json.dumps(dict(name=name, n=g.n, Q=len(dests), K=len(dests[0]), G=dg.info()["grid_dim"],
                          create_s=round(t_create, 3), runs=out)), flush=True)
    dp.close(); dg.close()

if __name__ == "__main__":
    which = sys.argv[1:] or ["grid300", "grid1000", "roadlike"]
    if "grid300" in which:
        g = graphs.jittered_grid(300, seed=123)
        run("grid300_K12", g, [graphs.pick_destinations(g, 12, seed=7)], [100, 900, 2000])
    if "grid1000" in which:
        g = graphs.jittered_grid(1000, seed=123)
        run("grid1000_K12", g, [graphs.pick_destinations(g, 12, seed=7)], [100, 900, 2000, 2000])
    if "roadlike" in which:
        g = graphs.road_like(7000, seed=11)
        rng = np.random.default_rng(1)
        dests = [graphs.pick_destinations(g, 2, seed=int(s)) for s in rng.integers(0, 1 << 30, 1024)]
        run("roadlike7k_Q1024_K2", g, dests, [100, 900, 2000])
    if "bugtrap" in which:
        g, s_, t_, ps = graphs.bugtrap_grid(1131, seed=123)
        run("bugtrap1131_K7_pseudo", g, [[s_] + ps + [t_]], [100] * 8, goal_bias=0.3, pseudo=True)
    if "bench" in which:
        g = graphs.jittered_grid(1000, seed=123)
        rng = np.random.default_rng(2)
        Q = int(os.environ.get("QP_Q", "296"))
        dests = [graphs.pick_destinations(g, 12, seed=int(s)) for s in rng.integers(0, 1 << 30, Q)]
        run(f"grid1000_K12_Q{Q}", g, dests, [3000])
    if "occ" in which:
        # residency experiment: same per-query work, 296 vs 592 queries on a graph small enough for both
        g = graphs.jittered_grid(700, seed=123)
        rng = np.random.default_rng(2)
        Q = int(os.environ.get("QP_Q", "296"))
        dests = [graphs.pick_destinations(g, 12, seed=int(s)) for s in rng.integers(0, 1 << 30, Q)]
        run(f"grid700_K12_Q{Q}", g, dests, [3000])
    if "sanpablo" in which:
        import bench
        class A: workload = "san_pablo"; queries = int(os.environ.get("QP_Q", "1024")); grid = 0
        g, dests = bench.make_workload(A, 0)
        run(f"san_pablo_Q{A.queries}_K2", g, dests, [100, 900, 2000])
    if "multi" in which:
        g = graphs.jittered_grid(1000, seed=123)
        rng = np.random.default_rng(2)
        for Q in (8, 64):
            dests = [graphs.pick_destinations(g, 12, seed=int(s)) for s in rng.integers(0, 1 << 30, Q)]
            run(f"grid1000_K12_Q{Q}", g, dests, [100, 900, 2000])
