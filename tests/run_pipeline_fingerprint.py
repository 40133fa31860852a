"""Helper of test_gpu_pipeline.py: one pipeline pass in a fresh process (so that DP_BIN_* knobs apply); prints a fingerprint."""
import hashlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import density_planner_b200 as dp  # noqa: E402
from density_planner_b200.pipeline import OccupancyPipeline  # noqa: E402
from test_gpu_pipeline import _setup, T, STRIDE  # noqa: E402

args = dp.default_args()
car = dp.Car(args)
N = 90_007
_, _, denv, xe0, rho0, xref, uref = _setup(dp, args, car, N, seed=5)
# a tight cloud too (warp aggregation has something to merge): shrink the initial errors of the second half
xe0[N // 2:] *= 0.02
pipe = OccupancyPipeline(car, denv, args, T, STRIDE, N, rho0_bound=2.0)
res = pipe.run(xe0.cuda(), rho0.cuda(), xref, uref, args.dt_sim, to_host=True)
print(hashlib.sha256(res.ibuf.numpy().tobytes()).hexdigest())
