"""Multi-GPU correctness probe (torchrun, NCCL): the same queries sharded over however many ranks
there are -- static and dynamic (chunks pulled from a shared counter) -- must give the per-query
TREEHASHes of a single GPU; GA islands with best-tour migration must not be worse than island 0
alone.  Rank 0 prints one JSON line.
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/mgpu_check.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "imomd-rrtstar_b200"))
import numpy as np, torch, torch.distributed as dist
from imomd_b200 import graphs, sharding

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
g = graphs.road_like(8000, seed=9)
rng = np.random.default_rng(1)
dests = [graphs.pick_destinations(g, 3, seed=int(s)) for s in rng.integers(0, 10**6, 37)]
ex = sharding.CudaExecutor(local, max_resident=4)          # streaming planners on every rank
static = sharding.run_sharded(g, dests, 400, ex)
dynamic = sharding.run_sharded(g, dests, 400, ex, dynamic_chunk=3, tag="mgpu_check")
ex.close()
K = 14
r2 = np.random.default_rng(4)
D = r2.uniform(100.0, 5000.0, (K, K)); D = (D + D.T) / 2; np.fill_diagonal(D, 0.0)
cost, tour, hist, evaluated = sharding.run_islands(np.ascontiguousarray(D), 0, K - 1, 3, device=local, ga=(2000, 300, 3))
if world > 1:
    t = torch.tensor([cost], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN)
    best = float(t[0])
else:
    best = cost
if rank == 0:
    print(json.dumps(dict(world=world, static=[r["treehash"] for r in static], dynamic=[r["treehash"] for r in dynamic],
                          expansions=[r["expansions"] for r in static], islands_best=best, island0_first=hist[0]["local"],
                          island0_final=cost)))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
