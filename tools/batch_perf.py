"""Batched mode at bench scale (development aid): 1 M-vertex grid, K = 12.
   python tools/batch_perf.py [Q] [slots] [passes] [B]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "imomd-rrtstar_b200"))
import numpy as np
from imomd_b200 import capi, graphs

Q = int(sys.argv[1]) if len(sys.argv) > 1 else 296
slots = int(sys.argv[2]) if len(sys.argv) > 2 else 296
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 3000
Bs = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [8]
K = 12
g = graphs.jittered_grid(1000, seed=123)
dests = []
for q in range(Q):
    rng = np.random.default_rng(1000 + q)
    dests.append([int(v) for v in rng.choice(g.n, size=K, replace=False)])
dg = capi.DeviceGraph(g)
probs = np.stack([capi.selection_probabilities(g, d) for d in dests])
dp = capi.DevicePlanner(dg, dests, probs, max_resident=slots if slots < Q else 0)
dp.set_profiling(True)
print(dp.info(), flush=True)
dp.feed(capi.mt19937_words(0, passes * K * 8 + 4096))
for B in Bs:
    if not dp.info()["streaming"]:
        dp.reset()
    t0 = time.time()
    reps = dp.expand_batch(passes // B, B)
    wall = time.time() - t0
    ms = dp.last_kernel_ms()
    ex = sum(r.expansions for r in reps)
    busy = np.array([r.busy_cycles for r in reps], np.float64)
    st = dp.states()
    conn = sum(1 for q in range(Q) if np.isfinite(st["D"][q]).all())
    # where the steps went (query-averaged): attach batches and relaxations of the trees summed
    # (group cycles) / slowest tree (attach + relaxation) / that part as the block saw it /
    # connection checks + tests, and the groups' attach cycles by kind of work
    pc = np.zeros(8); pn = np.zeros(8); gc = np.zeros(8)
    nq = min(Q, 64)
    for q in range(nq):
        prof = dp.profile(q)
        names = ["nn", "rescan", "mhins", "conn", "bfs", "check", "serial"]
        for i, nme in enumerate(names):
            pc[i + 1] += prof[nme][0]; pn[i + 1] += prof[nme][1]
        c, _ = dp.phase_cycles(q)
        gc += c.astype(np.float64)
    mc = lambda x: round(float(x) / nq / 1.965e6, 1)
    cnt = dp.counters()[:nq].astype(np.float64).mean(axis=0)
    print(json.dumps(dict(B=B, per_query_ms=dict(sum_tree_attach=mc(pc[1]), sum_tree_relaxation=mc(pc[4]),
                                                 slowest_tree_per_step=mc(pc[2]), attach_and_relax_part=mc(pc[3]),
                                                 connect_and_tests=mc(pc[5])),
                          steps=int(pn[2] / nq), relax_levels_per_query=int(pn[3] / nq), relax_items_per_query=int(pn[4] / nq),
                          relax_changes_per_query=int(cnt[4]), checked_vertices_per_query=int(cnt[6]),
                          group_ms_by_work=dict(nn=mc(gc[1]), rescan=mc(gc[2]), mhins=mc(gc[3]), serial=mc(gc[7])))), flush=True)
    print(json.dumps(dict(mode="batched", Q=Q, slots=slots, passes=passes, B=B, kernel_ms=round(ms, 1), wall_s=round(wall, 3),
                          exp_per_s=round(ex / (ms / 1e3)), expansions=ex,
                          busy_ms_min_mean_max=[round(float(x) / 1.965e6, 1) for x in (busy.min(), busy.mean(), busy.max())],
                          fully_connected_queries=conn, tree_vertices=sum(r.tree_vertices for r in reps),
                          unresolved=sum(r.unresolved_ties for r in reps), certified=sum(r.certified_ties for r in reps),
                          status=sorted(set(r.status for r in reps)))), flush=True)
if not dp.info()["streaming"]:
    dp.reset()
reps = dp.expand(passes)
ms = dp.last_kernel_ms()
ex = sum(r.expansions for r in reps)
st = dp.states()
conn = sum(1 for q in range(Q) if np.isfinite(st["D"][q]).all())
print(json.dumps(dict(mode="exact", Q=Q, passes=passes, kernel_ms=round(ms, 1), exp_per_s=round(ex / (ms / 1e3)), expansions=ex,
                      fully_connected_queries=conn, tree_vertices=sum(r.tree_vertices for r in reps))), flush=True)
dp.close(); dg.close()
