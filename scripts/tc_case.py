#!/usr/bin/env python
"""One tcgen05 rollout case vs the oracle: python scripts/tc_case.py N T [stride]  (env DP_IMPL selects the kernel, DP_TC_VARIANT its variant)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import density_planner_b200 as dp
import dp_oracle as O
N, T = int(sys.argv[1]), int(sys.argv[2])
stride = int(sys.argv[3]) if len(sys.argv) > 3 else 1
args = dp.default_args(); car = dp.Car(args)
torch.manual_seed(3)
up, uref, xref = car.get_valid_ref(args)
uref, xref = uref[:, :, :T], xref[:, :, :T + 1]
xe0 = car.sample_xe0(N)
xe, rho, margin = car.compute_density(xe0.cuda(), xref, uref, args.dt_sim, log_density=True, impl=os.environ.get("DP_IMPL", "tc"), return_margin=True, stride=stride)
torch.cuda.synchronize()
o_xe, o_rho, o_m, _ = O.rollout(O.load_weights(), xe0[:, :, 0].numpy(), xref[0].numpy(), uref[0].numpy(), args.dt_sim, want_margin=True, stride=stride)
xe, rho, margin = xe.cpu().numpy(), rho[:, 0, :].cpu().numpy(), margin.cpu().numpy()
safe = np.minimum(margin, o_m) > 2e-5
print("mode=%s N=%d T=%d stride=%d: states %.2e | logdens safe %.2e rel %.2e | unsafe %d" % (os.environ.get("DP_TC_VARIANT", "-"), N, T, stride,
      np.abs(xe - o_xe).max(), np.abs(rho - o_rho)[safe].max(), (np.abs(rho - o_rho)[safe] / (np.abs(o_rho[safe]) + .1)).max(), int((~safe).sum())))
