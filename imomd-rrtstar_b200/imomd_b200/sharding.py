"""Multi-GPU plumbing: independent queries are partitioned over ranks (one process per GPU),
each rank holds a full copy of the road graph, there is no collective on the data path
(the expansion of ONE query does not shard).  The only collectives are the
gathers of per-query results / counters at the end, and -- for the GA islands of config 4 --
the per-generation best-tour migration (`migrate_best`).

`torch.distributed` is used as plumbing: backend "nccl" on GPUs, "gloo" in the CPU tests.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

TOUR_CAP = 64            # tour slots of a migration record: {f64 cost, i32 len, i32 tour[64]}


def shard(n_queries: int, world: int, rank: int):
    """Query q runs on GPU q mod world."""
    return list(range(rank, n_queries, world))


def _gather(n_queries, mine, local, world):
    if world == 1:
        out = [None] * n_queries
        for q, res in zip(mine, local):
            out[q] = res
        return out
    gathered = [None] * world
    dist.all_gather_object(gathered, list(zip(mine, local)))
    out = [None] * n_queries
    for part in gathered:
        for q, res in part:
            out[q] = res
    return out


def run_sharded(graph, dests, iters, executor, world=None, rank=None, dynamic_chunk=0, tag="imomd"):
    """Run `executor(graph, my_dests, iters)` on this rank's queries and gather every rank's
    per-query results in query order.  The executor returns one picklable result per query
    (the CUDA executor: capi.DevicePlanner; the tests: the oracle).

    dynamic_chunk = 0: query q runs on rank q mod world (static).  dynamic_chunk = c > 0: ranks
    take chunks of c consecutive queries from a shared counter in the process group's store
    (TCPStore `add`, no collective): a rank whose queries turn out cheap simply takes more, so a
    job ends with the slowest CHUNK instead of the slowest static share.  Queries are independent:
    the per-query results do not depend on who ran them."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
        rank = dist.get_rank() if dist.is_initialized() else 0
    n = len(dests)
    if dynamic_chunk <= 0 or world == 1:
        mine = shard(n, world, rank)
        local = executor(graph, [dests[i] for i in mine], iters) if mine else []
        assert len(local) == len(mine)
        return _gather(n, mine, local, world)
    store = dist.distributed_c10d._get_default_store()
    key = f"{tag}_next_query"
    mine, local = [], []
    while True:
        start = store.add(key, dynamic_chunk) - dynamic_chunk
        if start >= n:
            break
        ids = list(range(start, min(start + dynamic_chunk, n)))
        res = executor(graph, [dests[i] for i in ids], iters)
        assert len(res) == len(ids)
        mine += ids
        local += res
    return _gather(n, mine, local, world)


class CudaExecutor:
    """Executor over the C ABI on one GPU.  The road graph is uploaded once and kept across
    calls; every call builds a planner for its queries (streaming through `max_resident` state
    slots when there are more queries than that), runs `iters` passes and returns, per query,
    the report figures, the device-computed TREEHASH and the distance matrix."""

    def __init__(self, device: int, max_resident: int = -1):
        self.device, self.max_resident = device, max_resident
        self._graph, self._dg = None, None
        self.kernel_ms = 0.0

    def _device_graph(self, graph):
        from . import capi
        if self._graph is not graph:
            self.close()
            self._dg, self._graph = capi.DeviceGraph(graph, device=self.device), graph
        return self._dg

    def __call__(self, graph, my_dests, iters):
        from . import capi
        dp = capi.DevicePlanner(self._device_graph(graph), my_dests, max_resident=self.max_resident)
        # multi-query planners are fed up front: a query that starved (status 1) would need its
        # own remainder on a re-call, so starvation is an error here, not a resume point
        dp.feed(capi.mt19937_words(0, iters * len(my_dests[0]) * 8 + 4096))
        reps = dp.expand(iters)
        self.kernel_ms += dp.last_kernel_ms()
        bad = [(q, r.status) for q, r in enumerate(reps) if r.status != 0]
        if bad:
            dp.close()
            raise capi.ImomdError(f"device run status {bad[:4]} (query, status)")
        st = dp.states()
        hashes = [int(r.tree_hash) for r in reps] if dp.info()["streaming"] else \
            [dp.treehash(q)[0] for q in range(len(reps))]
        out = [dict(expansions=int(r.expansions), tree_vertices=int(r.tree_vertices), status=int(r.status),
                    treehash=hashes[q], unresolved_ties=int(r.unresolved_ties), busy_cycles=int(r.busy_cycles),
                    D=st["D"][q].copy())
               for q, r in enumerate(reps)]
        dp.close()
        return out

    def close(self):
        if self._dg is not None:
            self._dg.close()
            self._dg, self._graph = None, None


def cuda_executor(device: int, max_resident: int = -1):
    return CudaExecutor(device, max_resident)


def pack_best(cost: float, tour) -> torch.Tensor:
    """Migration record as one int64 tensor: [cost bits, len, tour[TOUR_CAP]]."""
    rec = np.zeros(2 + TOUR_CAP, np.int64)
    rec[0] = np.array([cost], np.float64).view(np.int64)[0]
    n = min(len(tour), TOUR_CAP)
    rec[1] = n
    rec[2:2 + n] = np.asarray(tour[:n], np.int64)
    return torch.from_numpy(rec)


def unpack_best(rec: torch.Tensor):
    a = rec.cpu().numpy()
    cost = float(a[:1].view(np.float64)[0])
    n = int(a[1])
    return cost, [int(v) for v in a[2:2 + n]]


def migrate_best(cost: float, tour, device=None):
    """All-gather every island's best tour and return the global best (lowest cost, lowest
    rank on ties) plus every rank's cost.  272 bytes per rank over NCCL (NVLink / NVSwitch on a
    B200 box): latency-bound, one collective per GA generation."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return cost, list(tour), 0, [cost]
    world = dist.get_world_size()
    mine = pack_best(cost, tour)
    if device is not None:
        mine = mine.to(device)
    bucket = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(bucket, mine)
    recs = [unpack_best(b) for b in bucket]
    costs = [c for c, _ in recs]
    winner = int(np.argmin(np.asarray(costs)))      # argmin takes the first = lowest rank on ties
    return recs[winner][0], recs[winner][1], winner, costs


def run_islands(D, source, target, generations, device=0, seed_base=0, ga=(10000, 1000, 5)):
    """GA islands (BASELINE config 4): island = rank, own population from mt19937{seed_base+rank},
    best-tour migration after every generation.  Returns (global best cost, tour, history)."""
    from . import capi

    rank = dist.get_rank() if dist.is_initialized() else 0
    fs = capi.FacadeSolver(device, ga=(ga[0], ga[1], generations))
    fs.island_begin(D, source, target, seed_base + rank)
    history = []
    dev = torch.device("cuda", device) if (dist.is_initialized() and dist.get_backend() == "nccl") else None
    for gen in range(generations + 1):
        cost, tour = fs.island_best()
        gcost, gtour, winner, costs = migrate_best(cost, tour, dev)
        history.append(dict(generation=gen, local=cost, best=gcost, winner=winner))
        if winner != rank and gcost < cost:
            fs.island_accept(gcost, gtour)
        if gen == generations or not fs.island_step():
            break
    cost, tour = fs.island_best()
    evaluated = fs.island_evaluated()
    fs.close()
    return cost, tour, history, evaluated
