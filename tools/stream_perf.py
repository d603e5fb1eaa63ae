"""Streaming planner at bench scale (development aid): 1 M-vertex grid, K = 12, Q queries through
`slots` state slots; kernel time, expansions/s, per-query busy-cycle statistics.
   python tools/stream_perf.py [Q] [slots] [iters] [first_seed]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "imomd-rrtstar_b200"))
import numpy as np
from imomd_b200 import capi, graphs

Q = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
slots = int(sys.argv[2]) if len(sys.argv) > 2 else 296
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3000
seed0 = int(sys.argv[4]) if len(sys.argv) > 4 else 1000
K = 12
g = graphs.jittered_grid(1000, seed=123)
dests = []
for q in range(Q):
    rng = np.random.default_rng(seed0 + q)
    dests.append([int(v) for v in rng.choice(g.n, size=K, replace=False)])
dg = capi.DeviceGraph(g)
t0 = time.time()
probs = np.stack([capi.selection_probabilities(g, d) for d in dests])
dp = capi.DevicePlanner(dg, dests, probs, max_resident=slots)
print("create", round(time.time() - t0, 2), "s", dp.info(), flush=True)
dp.feed(capi.mt19937_words(0, iters * K * 8 + 4096))
for rep_i in range(2):
    t0 = time.time()
    reps = dp.expand(iters)
    wall = time.time() - t0
    ms = dp.last_kernel_ms()
    ex = sum(r.expansions for r in reps)
    busy = np.array([r.busy_cycles for r in reps], np.float64)
    print(json.dumps(dict(Q=Q, slots=slots, iters=iters, kernel_ms=round(ms, 1), wall_s=round(wall, 3),
                          exp_per_s=round(ex / (ms / 1e3)), expansions=ex,
                          busy_ms_min_mean_max=[round(float(x) / 1.965e6, 1) for x in (busy.min(), busy.mean(), busy.max())],
                          sum_busy_over_slots_ms=round(float(busy.sum()) / 1.965e6 / slots, 1),
                          unresolved=sum(r.unresolved_ties for r in reps), certified=sum(r.certified_ties for r in reps),
                          near=sum(r.near_ties for r in reps), status=sorted(set(r.status for r in reps)))), flush=True)
    if rep_i == 0 and not dp.info()["streaming"]:
        dp.reset()
dp.close(); dg.close()
