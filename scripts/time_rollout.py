#!/usr/bin/env python
"""CUDA-event timing of one rollout configuration: python scripts/time_rollout.py impl [N] [T] [stride] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import density_planner_b200 as dp
if os.environ.get("DP_LIB_VARIANT"):  # development: a variant build made by scripts/build_variant.sh
    import density_planner_b200._lib as _lib
    _lib.LIB_PATH = os.path.join(ROOT, "build", "variants", os.environ["DP_LIB_VARIANT"] + ".so")
impl = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
T = int(sys.argv[3]) if len(sys.argv) > 3 else 100
stride = int(sys.argv[4]) if len(sys.argv) > 4 else 10
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
args = dp.default_args(); car = dp.Car(args)
torch.manual_seed(3)
up, uref, xref = car.get_valid_ref(args)
uref, xref = uref[:, :, :T].cuda(), xref[:, :, :T + 1].cuda()
xe0 = car.sample_xe0(N).cuda()
for _ in range(2):
    car.compute_density(xe0, xref, uref, args.dt_sim, log_density=True, impl=impl, stride=stride)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    car.compute_density(xe0, xref, uref, args.dt_sim, log_density=True, impl=impl, stride=stride)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print("impl=%s N=%d T=%d stride=%d: %.3f ms  %.4e sample-steps/s" % (impl, N, T, stride, ms, N * T / ms * 1e3))
