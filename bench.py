#!/usr/bin/env python
"""bench.py -- RRT* expansions/s of the IMOMD-RRT* multi-tree expansion hot path.

  python bench.py --gpus N --steps K --warmup W            the CUDA path (this repo)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU path

Workload (BASELINE.json configs[2], the configuration the north-star target is quoted on):
jittered 1000 x 1000 grid road graph (1 M vertices), K = 12 trees per query (source, target,
10 destinations), default goal_bias 1.0, seed-0 mt19937 sample stream.  One *step* is one pass
of the hot path over one batch of synthetic input: `--queries` independent queries, each
expanded from scratch for `--iters` passes of expandTreeLayers_ (default 3000, the window
BASELINE.md quotes the CPU numbers on).  Metric: expansions/s = non-early-returning
expandTree_ calls per second, whole job (all queries, all GPUs).

  value   device-resident: query state already in HBM, timed region = the persistent
          expansion kernel (CUDA events on the planner's stream)
  e2e     through the C ABI with HOST buffers every step: destinations + selection
          probabilities + the mt19937 word stream go host->device (pinned memory), the
          state is rebuilt, the kernel runs, per-query reports and the K x K distance /
          connection-node matrices come back device->host
  roofline       algorithmic bytes of the expansion kernel (work counters x the per-unit
                 figures of SURVEY.md 8d, DESIGN.md "Roofline") / its device time, against
                 the measured HBM copy bandwidth of MEASURED_PEAKS.json
  cpu_baseline   the unmodified reference (oracle/_ref/ref_probe) on the box's host cores,
                 one process per core, a bounded sample of the same workload

A step processes `--queries` (2368) queries per GPU through `--slots` (296 = two CTAs per SM)
sets of per-query device state: the planner STREAMS -- CTAs take the next query from a device
work queue as their slot falls free, so a step is eight "waves" deep and ends with the slowest
tail of the last ones instead of the slowest query of a single wave (per-query busy time spreads
from 0.55 x to 2.3 x its mean: `batch` in the line).

  parity   the unmodified reference runs queries 0 .. cores-1 of the same batch at the same depth
           (it is the cpu_baseline leg); its TREEHASH, expansion count and tree
           sizes are compared with what the timed launch reported for those queries.  A mismatch
           fails the run.
  san_pablo_1024, ga   side legs (BASELINE configs[4], GA fitness path), see DESIGN.md section 7

With N > 1 (torchrun) every rank drives its own GPU with its own `--queries` queries
(independent queries shard with no data-path collective): weak scaling, value = all ranks'
expansions / max-over-ranks device time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "imomd-rrtstar_b200"))

import numpy as np  # noqa: E402

METRIC = "rrt_star_expansions_per_s"
UNIT = "expansions/s"
K_TREES = 12
REF_PROBE = os.path.join(ROOT, "oracle", "_ref", "ref_probe")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--queries", type=int, default=2368, help="independent queries per GPU and step")
    ap.add_argument("--slots", type=int, default=296,
                    help="sets of per-query device state (queries stream through them); 0 = one per query")
    ap.add_argument("--no-side-legs", action="store_true", help="skip the san_pablo_1024 / ga side measurements")
    ap.add_argument("--iters", type=int, default=3000, help="expandTreeLayers_ passes per query")
    ap.add_argument("--grid", type=int, default=1000, help="grid width (1000 = 1 M vertices)")
    ap.add_argument("--workload", default="grid", choices=["grid", "san_pablo"],
                    help="grid = BASELINE configs[2] (default); san_pablo = configs[4]: 1024 "
                         "source/target queries (K = 2) on the bundled san_pablo road graph")
    ap.add_argument("--no-ttfs", action="store_true", help="skip the time-to-first-solution side measurement")
    ap.add_argument("--ttfs-only", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--batch", type=int, default=16,
                    help="B of the batched-mode leg (B expansions per tree per step); 0 skips the leg")
    ap.add_argument("--cpu-sample-iters", type=int, default=0,
                    help="iterations of the CPU sample (0 = same as --iters)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def make_workload(args, rank):
    from imomd_b200 import graphs

    if getattr(args, "workload", "grid") == "san_pablo":
        # configs[4]: the reference's san_pablo.osm as parsed by its own parser (fixture exported
        # by tests/golden/make_golden.py), rebuilt from its edge list so that both arms see the
        # same adjacency order; source/target pairs from the largest component
        src = graphs.RoadGraph.load(os.path.join(ROOT, "tests", "golden", "graphs", "san_pablo.npz"))
        g = graphs.from_edges(src.lat, src.lon, *src.edge_list(), name="san_pablo")
        lab = g.components()
        big = np.nonzero(lab == np.bincount(lab).argmax())[0]
        rng = np.random.default_rng(11 + rank)
        dests = [[int(v) for v in rng.choice(big, size=2, replace=False)] for _ in range(args.queries)]
        return g, dests
    g = graphs.jittered_grid(args.grid, seed=123)
    # query q of rank r: destinations drawn with seed 1000 + r*queries + q (query 0 of rank 0
    # uses seed 7)
    dests = []
    for q in range(args.queries):
        seed = 7 if (rank == 0 and q == 0) else 1000 + rank * args.queries + q
        rng = np.random.default_rng(seed)
        dests.append([int(v) for v in rng.choice(g.n, size=K_TREES, replace=False)])
    return g, dests


def workload_name(args, iters, K):
    if getattr(args, "workload", "grid") == "san_pablo":
        return f"san_pablo_osm_K{K}_iters{iters}"
    return f"grid{args.grid}x{args.grid}_K{K}_iters{iters}"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (profiling recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4)
                          if len(r) > 2 + i and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(self.rows)}


# ----------------------------------------------------------------- CPU reference ---

def run_reference_sample(g, dests, iters, cores, graph_file=None):
    """`cores` concurrent ref_probe processes, one query each; returns (expansions/s, info)."""
    if not os.path.exists(REF_PROBE):
        return None, "oracle/_ref/ref_probe not built"
    own = graph_file is None
    if own:
        fd, graph_file = tempfile.mkstemp(suffix=".bin", prefix="imomd_graph_")
        os.close(fd)
        g.write_probe_file(graph_file)
    try:
        procs = []
        t0 = time.time()
        for d in dests[:cores]:
            procs.append(subprocess.Popen(
                [REF_PROBE, "--graph", graph_file, "--dest", ",".join(map(str, d)), "--iters", str(iters),
                 "--goal-bias", "1.0"], stdout=subprocess.PIPE, text=True))
        outs = [json.loads(p.communicate()[0].strip().splitlines()[-1]) for p in procs]
        wall = time.time() - t0
    finally:
        if own:
            os.unlink(graph_file)
    exp = sum(o["expansions"] for o in outs)
    secs = max(o["seconds"] for o in outs)          # loop time only (graph load excluded)
    return exp / secs, {"expansions": exp, "loop_seconds": secs, "wall_seconds": wall,
                        "per_process_exp_per_s": [round(o["expansions_per_s"], 1) for o in outs],
                        "toolchain": outs[0]["toolchain"], "runs": outs}


class _stdout_to_stderr:
    """The facade prints the reference's progress rows on stdout (C++ iostream); keep the bench's
    stdout to the one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def measure_ttfs(g, dest, device, max_iter=2500):
    from imomd_b200 import capi
    with _stdout_to_stderr():
        got = capi.facade_run(g, dest, max_iter=max_iter, max_time=300, device=device)
    out = {"b200_s": got["rows"][0][0] if got["rows"] else None,
           "b200_first_cost": got["rows"][0][1] if got["rows"] else None,
           "b200_first_iteration": got["rows"][0][3] if got["rows"] else None, "max_iter": max_iter}
    if os.path.exists(REF_PROBE):
        fd, gf = tempfile.mkstemp(suffix=".bin", prefix="imomd_graph_")
        os.close(fd)
        try:
            g.write_probe_file(gf)
            res = json.loads(subprocess.check_output(
                [REF_PROBE, "--graph", gf, "--dest", ",".join(map(str, dest)), "--mode", "run",
                 "--max-iter", str(max_iter), "--max-time", "300"], text=True).strip().splitlines()[-1])
        finally:
            os.unlink(gf)
        rows = res["rows"]
        out.update({"reference_s": rows[0][0] if rows else None,
                    "reference_first_cost": rows[0][1] if rows else None,
                    "reference_first_iteration": rows[0][3] if rows else None,
                    "same_first_solution": bool(rows and got["rows"] and rows[0][1] == got["rows"][0][1]
                                                and rows[0][3] == got["rows"][0][3])})
    return out


def measure_ttfs_bugtrap(device, width=1131, max_iter=10000):
    """BASELINE configs[1] (stand-in for the missing sanf.osm): bug-trap grid, pseudo
    mode, 5 pseudo destinations, goal_bias 0.3; the drop-in facade next to the unmodified
    reference: time to first solution, and the time and cost after `max_iter` passes."""
    from imomd_b200 import capi, graphs
    g, s, t, ps = graphs.bugtrap_grid(width, seed=123)
    dest = [s] + ps + [t]
    with _stdout_to_stderr():
        got = capi.facade_run(g, dest, goal_bias=0.3, max_iter=max_iter, max_time=600, pseudo=True,
                              device=device)
    rows = got["rows"]
    out = {"workload": f"bugtrap grid {width}x{width} ({g.n} vertices), pseudo mode, 5 pseudo destinations, "
                       f"goal_bias 0.3, {max_iter} passes",
           "b200_first_s": rows[0][0] if rows else None, "b200_first_cost": rows[0][1] if rows else None,
           "b200_last_s": rows[-1][0] if rows else None, "b200_final_cost": got["final_cost"],
           "expansions": got["expansions"]}
    if os.path.exists(REF_PROBE):
        fd, gf = tempfile.mkstemp(suffix=".bin", prefix="imomd_graph_")
        os.close(fd)
        try:
            g.write_probe_file(gf)
            res = json.loads(subprocess.check_output(
                [REF_PROBE, "--graph", gf, "--dest", ",".join(map(str, dest)), "--mode", "run", "--pseudo",
                 "--goal-bias", "0.3", "--max-iter", str(max_iter), "--max-time", "600"],
                text=True).strip().splitlines()[-1])
        finally:
            os.unlink(gf)
        rr = res["rows"]
        out.update({"reference_first_s": rr[0][0] if rr else None,
                    "reference_first_cost": rr[0][1] if rr else None,
                    "reference_last_s": rr[-1][0] if rr else None,
                    "reference_final_cost": res["final_cost"],
                    "identical_rows": [(r[3], r[2], r[1]) for r in rows] == [(r[3], r[2], r[1]) for r in rr],
                    "identical_path": list(got["path"]) == res["path"]})
    return out


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def reference_arm(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    g, dests = make_workload(args, 0)
    while len(dests) < cores:      # more cores than queries: more queries of the same family
        rng = np.random.default_rng(5000 + len(dests))
        dests.append([int(v) for v in rng.choice(g.n, size=len(dests[0]), replace=False)])
    iters = args.cpu_sample_iters or args.iters
    fd, gf = tempfile.mkstemp(suffix=".bin", prefix="imomd_graph_")
    os.close(fd)
    g.write_probe_file(gf)
    try:
        for _ in range(args.warmup):
            run_reference_sample(g, dests, min(iters, 200), cores, gf)
        rates, t_total, exp_total = [], 0.0, 0
        for s in range(args.steps):
            rate, info = run_reference_sample(g, dests, iters, cores, gf)
            if rate is None:
                emit({"impl": "reference", "unavailable": info})
                return 0
            rates.append(rate)
            t_total += info["loop_seconds"]
            exp_total += info["expansions"]
    finally:
        os.unlink(gf)
    value = exp_total / t_total
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "impl": "reference",
        "config": {"workload": workload_name(args, iters, len(dests[0])),
                   "graph": g.name, "trees_per_query": len(dests[0]),
                   "iters_per_query": iters, "goal_bias": 1.0,
                   "queries_per_step": cores, "note": "one reference process per host core"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": f"{cores} queries x {iters} passes per step, unmodified reference "
                                   f"(oracle/_ref/ref_probe), {cpu_model()}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------- CUDA arm ----

def algorithmic_bytes(counters, avg_degree, K):
    """SURVEY.md 8(d) per-unit figures x work counters (see DESIGN.md 'Roofline')."""
    c = counters.sum(axis=0).astype(np.float64)
    cells, recs_nn, calls, slots, reparent, desc, conn, recs_rescan = c
    nn = 8.0 * cells + 32.0 * (recs_nn + recs_rescan)          # cell headers + 32 B records
    rewire = calls * (8.0 + 24.0 * avg_degree) + 24.0 * reparent + 16.0 * desc
    connect = 12.0 * (K - 1) * conn
    return nn + rewire + connect, {"nn_bytes": nn, "rewire_bytes": rewire, "connect_bytes": connect}


def ga_leg(device):
    """GA path: EciGenSolver::solveRTSP through the facade's solver (host
    insertion / crossover bookkeeping, every fitness batch on the GPU: k_ga_eval, one warp per
    tour) on pseudo-complete K x K matrices, next to the unmodified reference's solver on the same
    matrices (reference timings: cpu_baseline.ga_solve_ms).  Same mt19937 streams, so cost and
    visiting sequence must be identical."""
    import struct
    from imomd_b200 import capi
    out = {}
    for K in (7, 12, 22):
        rng = np.random.default_rng(100 + K)
        D = rng.uniform(100.0, 5000.0, (K, K)); D = (D + D.T) / 2
        np.fill_diagonal(D, 0.0)
        D = np.ascontiguousarray(D)
        fs = capi.FacadeSolver(device)
        ms, costs, evaluated, seq = [], [], 0, None
        for _ in range(5):
            t0 = time.perf_counter()
            cost, seq, ev = fs.solve(D, 0, K - 1)
            ms.append((time.perf_counter() - t0) * 1e3)
            costs.append(cost); evaluated = ev
        fs.close()
        # fitness kernel alone: the 15 000 tours a default solve evaluates, in one launch
        ctx = capi.GaContext(device)
        tours = np.stack([np.concatenate(([0], rng.permutation(np.arange(1, K - 1)), [K - 1])) for _ in range(15000)]).astype(np.int32)
        lens = np.full(len(tours), K, np.int32)
        ctx.eval(D, tours, lens)
        t0 = time.perf_counter()
        for _ in range(10):
            ctx.eval(D, tours, lens)
        call_ms = (time.perf_counter() - t0) * 1e3 / 10
        ctx.close()
        entry = {"solve_ms_b200": [round(x, 3) for x in ms], "cost": costs[-1], "tours_evaluated_per_solve": int(evaluated),
                 "ga_eval_call_ms_15000_tours": round(call_ms, 3), "ga_eval_tours_per_s": round(15000 / (call_ms / 1e3))}
        if os.path.exists(REF_PROBE):
            fd, mf = tempfile.mkstemp(suffix=".bin", prefix="imomd_D_")
            with os.fdopen(fd, "wb") as f:
                f.write(struct.pack("<Q", K) + D.tobytes())
            try:
                ref = json.loads(subprocess.check_output([REF_PROBE, "--mode", "rtsp", "--matrix", mf, "--solves", "5"],
                                                         text=True).strip().splitlines()[-1])
            finally:
                os.unlink(mf)
            entry.update({"solve_ms_reference": [round(x, 3) for x in ref["ms"]],
                          "identical_costs": [float(c) for c in ref["costs"]] == [float(c) for c in costs],
                          "identical_sequence": ref["last_sequence"] == [int(v) for v in seq]})
        out[f"K{K}"] = entry
    return out


def san_pablo_leg(device, rank, world, iters=3000, n_queries=1024):
    """BASELINE configs[4]: 1024 source/target queries (K = 2) on san_pablo.osm, STRONG scaling:
    the same 1024 queries over however many GPUs there are (sharding.run_sharded, query q on rank
    q mod N, no data-path collective), every rank's planner streaming its share; per-query
    TREEHASHes gathered and folded into one fingerprint that must not depend on the number of GPUs
    (the driver's per-N lines carry it).  The job is bounded by its slowest query (one source /
    target pair keeps rewiring most of the 6 809-vertex map: 0.67 s of the 0.8 s a single GPU needs
    for all 1024), so more GPUs cannot shorten it much; dynamic chunks of 32 queries (measured on
    2 GPUs: 5.2 s) leave most SMs without a CTA."""
    import hashlib
    import torch
    from imomd_b200 import graphs, sharding
    src = graphs.RoadGraph.load(os.path.join(ROOT, "tests", "golden", "graphs", "san_pablo.npz"))
    g = graphs.from_edges(src.lat, src.lon, *src.edge_list(), name="san_pablo")
    lab = g.components()
    big = np.nonzero(lab == np.bincount(lab).argmax())[0]
    rng = np.random.default_rng(11)
    dests = [[int(v) for v in rng.choice(big, size=2, replace=False)] for _ in range(n_queries)]
    ex = sharding.CudaExecutor(device, max_resident=-1)
    ex(g, dests[:8], 50)                      # warm-up: graph upload, kernel load
    ex.kernel_ms = 0.0
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    # static shares (query q on rank q mod world): a K = 2 query is a few hundred microseconds of
    # work per pass, chunks small enough to balance would leave most SMs of a GPU without a CTA
    res = sharding.run_sharded(g, dests, iters, ex, world=world, rank=rank, dynamic_chunk=0, tag="san_pablo_1024")
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    kernel_s = ex.kernel_ms / 1e3
    ex.close()
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([wall, kernel_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall, kernel_s = float(t[0]), float(t[1])
    exps = sum(r["expansions"] for r in res)
    fp = hashlib.sha256(np.asarray([r["treehash"] for r in res], np.uint64).tobytes()).hexdigest()[:16]
    return {"workload": f"san_pablo_osm_K2_iters{iters}_{n_queries}_queries", "scaling": "strong", "n_gpus": world,
            "expansions": int(exps), "seconds": wall, "expansions_per_s": exps / wall,
            "max_rank_kernel_seconds": kernel_s, "tree_vertices": int(sum(r["tree_vertices"] for r in res)),
            "unresolved_ties": int(sum(r["unresolved_ties"] for r in res)),
            "slowest_query_seconds": max(r.get("busy_cycles", 0) for r in res) / 1.965e9,
            "mean_query_seconds": float(np.mean([r.get("busy_cycles", 0) for r in res])) / 1.965e9,
            "treehash_fingerprint": fp, "note": "fingerprint = sha256 over the 1024 per-query TREEHASHes in query "
            "order: identical for every GPU count"}


def ga_islands_leg(device, rank, world, generations=5):
    """BASELINE configs[3]: GA islands, one per GPU, own population (mt19937{rank}), best-tour
    migration after every generation = one NCCL all-gather of 528 bytes per rank.  K = 22
    pseudo-complete matrix (half of the pairs unconnected, as after a partial exploration)."""
    import torch
    from imomd_b200 import sharding
    K = 22
    rng = np.random.default_rng(4)
    D = rng.uniform(100.0, 5000.0, (K, K)); D = (D + D.T) / 2
    miss = np.triu(rng.random((K, K)) < 0.5, 1); D[miss | miss.T] = np.inf
    perm = rng.permutation(K)
    for a, b in zip(perm[:-1], perm[1:]):
        if np.isinf(D[a, b]):
            D[a, b] = D[b, a] = rng.uniform(100.0, 5000.0)
    np.fill_diagonal(D, 0.0)
    D = np.ascontiguousarray(D)
    # migration latency: the all-gather alone, device-timed
    mig_us = None
    if world > 1:
        import torch.distributed as dist
        rec = sharding.pack_best(1.0, [0, 1, 2]).to(torch.device("cuda", device))
        bucket = [torch.empty_like(rec) for _ in range(world)]
        for _ in range(5):
            dist.all_gather(bucket, rec)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(50):
            dist.all_gather(bucket, rec)
        e1.record(); torch.cuda.synchronize()
        mig_us = e0.elapsed_time(e1) * 1e3 / 50
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cost, tour, hist, evaluated = sharding.run_islands(D, 0, K - 1, generations, device=device)
    wall = time.perf_counter() - t0
    total_eval, best = float(evaluated), cost
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([total_eval], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        w = torch.tensor([wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(w, op=dist.ReduceOp.MAX)
        b = torch.tensor([cost], dtype=torch.float64, device="cuda")
        dist.all_reduce(b, op=dist.ReduceOp.MIN)
        total_eval, wall, best = float(t[0]), float(w[0]), float(b[0])
    return {"islands": world, "K": K, "generations": generations, "best_cost": best,
            "best_by_generation": [h["best"] for h in hist], "island0_local_by_generation": [h["local"] for h in hist],
            "tours_evaluated": int(total_eval), "seconds": wall, "tours_per_s": total_eval / wall,
            "migration_allgather_us": mig_us}


def batched_leg(dp, args, g, K, exact_states, peak, device):
    """The BATCHED mode (imomd_expand_batch) on the very queries of the timed batch: B expansions per
    tree per step, trees of a query attached concurrently, rewiring as a relaxation, connection
    checks with the step's final costs.  Not the reference's sequence (the exact mode above is the
    parity mode): reported beside it with its own roofline figure, a determinism check (two runs,
    identical TREEHASHes) and the quality of what it builds -- the RTSP tour over the distance
    matrix of the first queries against the exact mode's at the same number of passes."""
    from imomd_b200 import capi
    B = args.batch
    steps = args.iters // B
    streaming = dp.info()["streaming"]
    if not streaming:
        dp.reset()
    dp.expand_batch(steps, B)                                   # warm-up
    first = None
    ms, exps = [], []
    for _ in range(2):
        if not streaming:
            dp.reset()
        reps = dp.expand_batch(steps, B)
        bad = sorted({r.status for r in reps if r.status != 0})
        if bad:
            return {"error": f"device run status {bad}"}
        ms.append(dp.last_kernel_ms()); exps.append(sum(r.expansions for r in reps))
        hashes = [int(r.tree_hash) for r in reps] if streaming else [dp.treehash(q)[0] for q in range(min(len(reps), 16))]
        if first is None:
            first = hashes
    counters = dp.counters()
    bytes_step, parts = algorithmic_bytes(counters, g.arcs / g.n, K)
    st = dp.states()
    # quality: (1) the pairwise connection distances both modes found, entry by entry; (2) the RTSP
    # tour over the whole matrix for (up to 32) queries that both modes fully connected
    De, Db = exact_states["D"], st["D"]
    off = ~np.eye(K, dtype=bool)[None, :, :]
    fe, fb = np.isfinite(De) & off, np.isfinite(Db) & off
    both_f = fe & fb
    ratio_d = (Db[both_f] / De[both_f]) if both_f.any() else np.zeros(0)
    full_e = np.isfinite(De).all(axis=(1, 2)); full_b = np.isfinite(Db).all(axis=(1, 2))
    ratios = []
    for q in np.nonzero(full_e & full_b)[0][:32]:
        fs = capi.FacadeSolver(device); ce, _, _ = fs.solve(np.ascontiguousarray(De[q]), 0, K - 1); fs.close()
        fs = capi.FacadeSolver(device); cb, _, _ = fs.solve(np.ascontiguousarray(Db[q]), 0, K - 1); fs.close()
        ratios.append(cb / ce)
    both = len(ratios)
    conn_b, conn_e = int(full_b.sum()), int(full_e.sum())
    launch_s = ms[-1] / 1e3
    return {"B": B, "steps_per_query": steps, "passes_per_query": steps * B,
            "value": sum(exps) / (sum(ms) / 1e3), "unit": UNIT, "kernel_ms": [round(x, 1) for x in ms],
            "expansions_per_launch": int(exps[-1]),
            "deterministic": first == hashes,
            "roofline": {"bound": "hbm", "achieved": bytes_step / launch_s / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": bytes_step / launch_s / 1e9 / peak, "kernel": "k_expand_batch",
                         "algorithmic_bytes_per_launch": bytes_step, **{k: float(v) for k, v in parts.items()}},
            "fully_connected_queries": {"batched": conn_b, "exact": conn_e, "of": len(reps)},
            "connected_pairs": {"batched": int(fb.sum() // 2), "exact": int(fe.sum() // 2), "both": int(both_f.sum() // 2),
                                "distance_ratio_batched_over_exact": {
                                    "mean": float(ratio_d.mean()) if ratio_d.size else None,
                                    "median": float(np.median(ratio_d)) if ratio_d.size else None,
                                    "p95": float(np.percentile(ratio_d, 95)) if ratio_d.size else None},
                                "note": "tree pairs with a connection (finite distance-matrix entry) after the same number "
                                        "of passes, all queries; ratio over the pairs both modes connected"},
            "tour_cost_vs_exact": {"queries_compared": both, "mean_ratio": float(np.mean(ratios)) if ratios else None,
                                   "max_ratio": float(np.max(ratios)) if ratios else None,
                                   "min_ratio": float(np.min(ratios)) if ratios else None,
                                   "note": "RTSP tour (EciGenSolver) over the distance matrix each mode built in the same "
                                           "number of passes, up to 32 queries that both modes fully connected"},
            "unresolved_ties": int(sum(r.unresolved_ties for r in reps)),
            "note": "semantics: DESIGN.md 'Batched mode'; parity of this mode is against its CPU restatement "
                    "(oracle/imomd_oracle.cpp batched_step), bit for bit, in tests/test_cuda_batched.py"}


def b200_arm(args):
    import torch
    from imomd_b200 import capi

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback)")
    # Time to first solution of ONE query through the drop-in C++ facade next to the unmodified
    # reference, in a process of its own and BEFORE this one touches the device: what a user of
    # the drop-in starts is a fresh process on an idle GPU.  (Measured: the same bug-trap run takes
    # 0.27-0.28 s that way, 0.32-0.35 s inside this process after it has cycled 143 GB through the
    # device allocator, 0.73 s in a child process while this one still holds its context.)
    ttfs_side = {}
    if not args.no_ttfs and not args.no_cpu_baseline and world == 1:
        try:
            out = subprocess.check_output([sys.executable, os.path.abspath(__file__), "--ttfs-only", "--grid",
                                           str(args.grid), "--gpus", "1"], text=True, timeout=900)
            ttfs_side = json.loads(out.strip().splitlines()[-1])
        except Exception as ex:
            ttfs_side = {"ttfs": {"error": repr(ex)}}
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    device = local
    torch.cuda.set_device(device)
    g, dests = make_workload(args, rank)
    Q, K = args.queries, len(dests[0])
    dg = capi.DeviceGraph(g, device=device)
    probs = np.stack([capi.selection_probabilities(g, d) for d in dests])
    slots = min(args.slots, Q) if args.slots > 0 else 0
    dp = capi.DevicePlanner(dg, dests, probs, max_resident=slots)
    layout = dp.info()
    n_words = args.iters * K * 8 + 4096
    words_np = capi.mt19937_words(0, n_words)

    # pinned host staging buffers for the end-to-end path
    pin_words = torch.empty(n_words, dtype=torch.int32).pin_memory()
    pin_dest = torch.empty(Q * K, dtype=torch.int32).pin_memory()
    pin_prob = torch.empty(Q * K * K, dtype=torch.float64).pin_memory()
    w_view = pin_words.numpy().view(np.uint32); w_view[:] = words_np
    d_view = pin_dest.numpy().view(np.uint32); d_view[:] = np.asarray(dests, np.uint32).ravel()
    p_view = pin_prob.numpy(); p_view[:] = probs.ravel()
    h2d = w_view.nbytes + d_view.nbytes + p_view.nbytes
    import ctypes
    d2h = Q * (ctypes.sizeof(capi.Report) + capi.sizeof_query_state())

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        """destinations + probabilities + sample stream in, kernel, reports + K x K state out.
        Returns (device ms of the kernel, expansions, e2e seconds, reports, states)."""
        t0 = time.perf_counter()
        dp.clear_words()
        dp.set_queries(d_view, p_view)           # H2D destinations + probabilities, rebuild state
        dp.feed(w_view)                          # H2D sample stream
        reps = dp.expand(args.iters)             # kernel + D2H reports
        st = dp.states()                         # D2H K x K matrices of every query, one copy
        e2e_s = time.perf_counter() - t0
        bad = [r.status for r in reps if r.status != 0]
        if bad:
            raise SystemExit(f"bench.py: device run status {bad[:4]}")
        return dp.last_kernel_ms(), sum(r.expansions for r in reps), e2e_s, reps, st

    for _ in range(max(args.warmup, 0)):
        one_step()
    barrier()
    sampler = ClockSampler(device)
    sampler.start()
    ker_ms, exps, e2e_secs = [], [], []
    reps = None
    for _ in range(args.steps):
        ms, ex, es, reps, states = one_step()
        ker_ms.append(ms); exps.append(ex); e2e_secs.append(es)
    barrier()
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    counters = dp.counters()
    n_check = min(os.cpu_count() or 1, Q)
    tree_hashes = [int(r.tree_hash) for r in reps[:n_check]] if layout["streaming"] else \
        [dp.treehash(q)[0] for q in range(n_check)]
    t_dev = sum(ker_ms) / 1e3
    t_e2e = sum(e2e_secs)
    exp_total = sum(exps)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e = torch.tensor([float(exp_total)], dtype=torch.float64, device="cuda")
        dist.all_reduce(e, op=dist.ReduceOp.SUM)
        t_dev, t_e2e = float(t[0]), float(t[1])
        exp_total = float(e[0])
    side = {}
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    if args.batch > 0 and rank == 0 and getattr(args, "workload", "grid") == "grid":
        try:
            side["batched"] = batched_leg(dp, args, g, K, states, peak, device)
        except Exception as ex:
            side["batched"] = {"error": repr(ex)}
    if not args.no_side_legs and getattr(args, "workload", "grid") == "grid":
        # side legs run on every rank (they shard); rank 0 reports them
        dp.close(); dp = None
        try:
            side["san_pablo_1024"] = san_pablo_leg(device, rank, world)
        except Exception as ex:
            side["san_pablo_1024"] = {"error": repr(ex)}
        try:
            side["ga_islands"] = ga_islands_leg(device, rank, world)
        except Exception as ex:
            side["ga_islands"] = {"error": repr(ex)}
    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return 0

    avg_deg = g.arcs / g.n
    # counters are cumulative over the last set_queries (= one step); per launch:
    bytes_step, parts = algorithmic_bytes(counters, avg_deg, K)
    launch_s = ker_ms[-1] / 1e3
    achieved = bytes_step / launch_s / 1e9
    busy = np.asarray([r.busy_cycles for r in reps], np.float64)
    sm_hz = (sampler.summary()["sm_mhz"] or 1965.0) * 1e6
    n_slots = layout["slots"]
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r02_k_expand_traffic.json")) as f:
            cap = json.load(f)
        traffic = cap["dram_bytes_read"] + cap["dram_bytes_write"]
        traffic_note = cap.get("note")
    except (OSError, KeyError, ValueError):
        traffic_note = None
    line = {
        "metric": METRIC, "value": exp_total / t_dev, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(args, args.iters, K),
                   "graph": g.name, "vertices": g.n,
                   "trees_per_query": K, "iters_per_query": args.iters, "goal_bias": 1.0,
                   "queries_per_gpu": Q, "state_slots_per_gpu": n_slots, "streaming": layout["streaming"],
                   "l2": "inputs_larger_than_l2 (state rebuilt every step)",
                   "hash_cells_per_axis": dg.info()["grid_dim"]},
        "e2e": {"value": exp_total / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h)},
        # k_fill_headers + k_expand per step (a streaming planner rebuilds its slots inside k_expand;
        # a resident one adds k_init_trees + k_lower_bounds)
        "gpu_launches": int(args.steps * (2 if layout["streaming"] else 4)),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "kernel": "k_expand",
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_step,
                     "launch_ms": ker_ms[-1], "traffic_note": traffic_note,
                     "note": "latency-bound: one CTA per query walks dependent parent/child chains; "
                             "see DESIGN.md 'Roofline'", **{k: float(v) for k, v in parts.items()}},
        "clocks": sampler.summary(),
        "batch": {"queries": Q, "slots": n_slots,
                  "per_query_busy_ms": {"min": float(busy.min() / sm_hz * 1e3), "mean": float(busy.mean() / sm_hz * 1e3),
                                        "max": float(busy.max() / sm_hz * 1e3)},
                  "slot_utilisation": float(busy.sum() / sm_hz / (n_slots * launch_s)),
                  "kernel_ms_per_step": {"min": min(ker_ms), "mean": float(np.mean(ker_ms)), "max": max(ker_ms)},
                  "note": "busy = SM cycles a query held its CTA (incl. slot rebuild and tree hash); utilisation = "
                          "sum of busy / (slots x launch time): what the tail of the launch leaves idle"},
        "tree_vertices_per_step": int(sum(r.tree_vertices for r in reps)),
        "near_ties": int(sum(r.near_ties for r in reps)),
        "certified_ties": int(sum(r.certified_ties for r in reps)),
        "unresolved_ties": int(sum(r.unresolved_ties for r in reps)),
    }
    line.update(side)
    if dp is not None:
        dp.close(); dp = None
    # time to first solution of ONE query through the drop-in C++ facade (planner construction ->
    # first accepted RTSP tour, same 100-pass cadence) next to the unmodified reference
    line.update(ttfs_side)
    if not args.no_side_legs and world == 1:
        try:
            line["ga"] = ga_leg(device)
        except Exception as ex:
            line["ga"] = {"error": repr(ex)}
    parity_failed = None
    if not args.no_cpu_baseline:
        cores = min(os.cpu_count() or 1, Q)
        iters = args.cpu_sample_iters or args.iters
        rate, info = run_reference_sample(g, dests, iters, cores)
        if rate is None:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": cores, "kind": "reference",
                                    "sample": info}
        else:
            line["cpu_baseline"] = {
                "value": rate, "unit": UNIT, "cores": cores, "kind": "reference",
                "sample": f"queries 0..{cores - 1} of the timed batch x {iters} passes, one unmodified-reference "
                          f"process per core ({cpu_model()}); single-core mean "
                          f"{np.mean(info['per_process_exp_per_s']):.0f} expansions/s"}
            if "ga" in line and isinstance(line["ga"], dict):
                line["cpu_baseline"]["ga_solve_ms"] = {k: v.get("solve_ms_reference") for k, v in line["ga"].items()
                                                       if isinstance(v, dict)}
            # parity of the TIMED batch: the reference's fingerprints of those very queries
            if iters == args.iters:
                same = 0
                diffs = []
                for q, o in enumerate(info["runs"]):
                    ok = (int(o["treehash"], 16) == tree_hashes[q] and
                          int(o["expansions"]) == int(reps[q].expansions) and
                          int(o["tree_vertices"]) == int(reps[q].tree_vertices))
                    same += bool(ok)
                    if not ok:
                        diffs.append({"query": q, "reference": [o["treehash"], o["expansions"], o["tree_vertices"]],
                                      "b200": [f"{tree_hashes[q]:016x}", int(reps[q].expansions),
                                               int(reps[q].tree_vertices)]})
                line["parity"] = {"queries_checked": len(info["runs"]), "identical": same, "depth": iters,
                                  "what": "TREEHASH (every vertex, parent and cost bit of the K trees), expansions, "
                                          "tree vertices of the timed launch vs the unmodified reference",
                                  "unresolved_ties_in_batch": line["unresolved_ties"]}
                if diffs:
                    line["parity"]["mismatches"] = diffs[:8]
                    parity_failed = f"{len(diffs)} of {len(info['runs'])} checked queries differ from the reference"
    emit(line)
    dg.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    if parity_failed:
        print("bench.py: PARITY FAILURE: " + parity_failed, file=sys.stderr)
        return 3
    return 0


_JSON_OUT = None


def emit(line):
    """the ONE JSON line, on the process's real stdout"""
    (_JSON_OUT or sys.stdout).write(json.dumps(line) + "\n")
    (_JSON_OUT or sys.stdout).flush()


def main():
    global _JSON_OUT
    args = parse()
    # stdout carries exactly one JSON line: whatever libraries print (NCCL's version banner, the
    # facade's progress rows) goes to stderr
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return reference_arm(args)
    if args.ttfs_only:
        # child of the main arm: time to first solution through the drop-in facade (same device)
        device = dist_env()[2]
        g, dests = make_workload(argparse.Namespace(grid=args.grid, queries=1, workload="grid"), 0)
        res = {}
        try:
            res["ttfs"] = measure_ttfs(g, dests[0], device)
        except Exception as ex:
            res["ttfs"] = {"error": str(ex)}
        try:
            res["ttfs_bugtrap"] = measure_ttfs_bugtrap(device)
        except Exception as ex:
            res["ttfs_bugtrap"] = {"error": str(ex)}
        res["ttfs"]["process"] = res["ttfs_bugtrap"]["process"] = "a fresh process per measurement leg (bench.py --ttfs-only)"
        emit(res)
        return 0
    return b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
