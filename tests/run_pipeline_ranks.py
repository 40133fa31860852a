"""Helper of test_gpu_pipeline.py (torchrun, 2 ranks, NCCL): sharded pipeline == single-rank pipeline, bit for bit."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import density_planner_b200 as dp  # noqa: E402
from density_planner_b200 import dist as D  # noqa: E402
from density_planner_b200.pipeline import OccupancyPipeline  # noqa: E402
from test_gpu_pipeline import _setup, T, STRIDE  # noqa: E402

rank, world, local_rank = D.init_from_env()
dev = torch.device("cuda", local_rank)
args = dp.default_args()
car = dp.Car(args, device=dev)
N = 200_003
_, _, denv, xe0, rho0, xref, uref = _setup(dp, args, car, N, seed=11)   # same seed on every rank: the same global set
pipe = OccupancyPipeline(car, denv, args, T, STRIDE, N, rho0_bound=2.0)
lo, hi = D.shard_range(N, rank, world)
res = pipe.run(xe0[lo:hi].pin_memory(), rho0[lo:hi].pin_memory(), xref, uref, args.dt_sim)       # 2 collectives inside
shard_buf = res.ibuf.clone()
cost = res.cost_slices.clone()
full = pipe.run(xe0.to(dev), rho0.to(dev), xref, uref, args.dt_sim, local=True)
ok = bool((shard_buf == full.ibuf).all()) and bool((cost == full.cost_slices).all()) and int(full.count.sum()) == N * pipe.Ts
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("PIPELINE_RANKS_OK" if int(flag.item()) == 1 else "PIPELINE_RANKS_MISMATCH")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
