"""N > 1 host-side logic on the CPU: world_size-2 `gloo` processes.  Queries are sharded with no
data-path collective and the gathered per-query results do not depend on the world size; the GA
migration record survives the all-gather and the global best is chosen deterministically."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _oracle_executor(graph, my_dests, iters):
    import oraclelib as ol
    out = []
    for d in my_dests:
        op = ol.OraclePlanner(graph, d)
        e = op.iterate(iters)
        out.append(dict(expansions=int(e), treehash=int(op.treehash()), D=op.matrices()["D"]))
        op.close()
    return out


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "imomd-rrtstar_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from imomd_b200 import graphs, sharding
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = graphs.road_like(2500, seed=9)
    rng = np.random.default_rng(1)
    dests = [graphs.pick_destinations(g, 2, seed=int(s)) for s in rng.integers(0, 10**6, 7)]
    res = sharding.run_sharded(g, dests, 150, _oracle_executor)
    # the same job with ranks pulling chunks of two queries from the shared counter
    dyn = sharding.run_sharded(g, dests, 150, _oracle_executor, dynamic_chunk=2, tag="gloo_test")
    # island migration: rank r offers cost 100 - r (so the last rank wins), ties -> lowest rank
    cost, tour, winner, costs = sharding.migrate_best(100.0 - rank, [0, rank + 1, 9])
    tie = sharding.migrate_best(5.0, [0, 7 + rank, 9])
    if rank == 0:
        q.put(dict(hashes=[r["treehash"] for r in res], exps=[r["expansions"] for r in res],
                   dyn_hashes=[r["treehash"] for r in dyn],
                   mine=sharding.shard(len(dests), world, rank), best=(cost, tour, winner, costs), tie=tie))
    dist.barrier()
    dist.destroy_process_group()


def test_sharding_and_migration_world2():
    sys.path.insert(0, os.path.join(ROOT, "imomd-rrtstar_b200"))
    from imomd_b200 import graphs, sharding
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process run of the same queries
    g = graphs.road_like(2500, seed=9)
    rng = np.random.default_rng(1)
    dests = [graphs.pick_destinations(g, 2, seed=int(s)) for s in rng.integers(0, 10**6, 7)]
    solo = sharding.run_sharded(g, dests, 150, _oracle_executor, world=1, rank=0)
    assert got["hashes"] == [r["treehash"] for r in solo]
    assert got["exps"] == [r["expansions"] for r in solo]
    assert got["dyn_hashes"] == got["hashes"]          # who ran a query does not matter
    assert got["mine"] == [0, 2, 4, 6]
    cost, tour, winner, costs = got["best"]
    assert (cost, tour, winner, costs) == (99.0, [0, 2, 9], 1, [100.0, 99.0])
    assert got["tie"][2] == 0 and got["tie"][1] == [0, 7, 9]


def test_shard_partition_properties():
    sys.path.insert(0, os.path.join(ROOT, "imomd-rrtstar_b200"))
    from imomd_b200 import sharding
    for n in (0, 1, 7, 1024):
        for world in (1, 2, 4, 8):
            parts = [sharding.shard(n, world, r) for r in range(world)]
            flat = sorted(x for p in parts for x in p)
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    rec = sharding.pack_best(1234.5678, [0, 3, 1, 2])
    assert sharding.unpack_best(rec) == (1234.5678, [0, 3, 1, 2])
