"""GA islands over N GPUs (torchrun): config 4 of BASELINE.json in miniature.
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/run_islands.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "imomd-rrtstar_b200"))
import numpy as np, torch, torch.distributed as dist
from imomd_b200 import sharding
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
K = 22
rng = np.random.default_rng(4)
D = rng.uniform(100.0, 5000.0, (K, K)); D = (D + D.T) / 2
miss = np.triu(rng.random((K, K)) < 0.5, 1); D[miss | miss.T] = np.inf
perm = rng.permutation(K)
for a, b in zip(perm[:-1], perm[1:]):
    if np.isinf(D[a, b]): D[a, b] = D[b, a] = rng.uniform(100.0, 5000.0)
np.fill_diagonal(D, 0.0)
cost, tour, hist, evaluated = sharding.run_islands(np.ascontiguousarray(D), 0, K - 1, 5, device=local)
if world > 1:
    t = torch.tensor([cost], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN)
    best = float(t[0])
else:
    best = cost
if rank == 0:
    print(json.dumps(dict(islands=world, best=best, island0=cost, history=hist, evaluated_island0=evaluated)))
if world > 1:
    dist.destroy_process_group()
