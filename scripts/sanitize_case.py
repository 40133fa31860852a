#!/usr/bin/env python
"""Small end-to-end case that touches every kernel family once: rollout (device and chunked host path), normaliser, collision
cost forward+backward, occupancy binning with the fused dot product, integrator forward+adjoint.  Meant for a memory checker
where one is available (compute-sanitizer is closed on the pool this was developed on) and as a quick sanity run."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import density_planner_b200 as dp
from density_planner_b200 import grid as G
from bench import synthetic_env

args = dp.default_args(); car = dp.Car(args)
dev = torch.device("cuda", 0)
torch.manual_seed(5)
T, N = 20, 1000
up, uref, xref = car.get_valid_ref(args)
uref, xref = uref[:, :, :T].cuda(), xref[:, :, :T + 1].cuda()
xe0 = car.sample_xe0(N).cuda()
xe, rho = car.compute_density(xe0, xref, uref, args.dt_sim, log_density=True, stride=5)
xh, rh = car.compute_density(xe0.cpu(), xref.cpu(), uref.cpu(), args.dt_sim, log_density=True, stride=5)
assert torch.equal(xh, xe.cpu())
grid, gx, gy = synthetic_env(args, 0, dev)
denv = dp.DeviceEnv(grid, gx, gy, args, device=dev)
x = (xe + xref[:, :, ::5]).detach().requires_grad_(True)
r = dp.normalize_density(rho, torch.ones(N, 1, 1, device=dev)).detach().requires_grad_(True)
c = dp.get_cost_coll(x, r, denv)
c.sum().backward()
out = G.bin_occupancy(denv, x[:, 0, :].detach().t().contiguous(), x[:, 1, :].detach().t().contiguous(),
                      r[:, 0, :].detach().t().contiguous(), args)
upc = up.cuda().requires_grad_(True)
_, xr = car.up2ref_traj(xref[:, :, :1].contiguous(), upc, args)
xr.sum().backward()
torch.cuda.synchronize()
print("sanitize case ok: cost %.4e" % float(c.detach().sum()))
