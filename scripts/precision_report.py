#!/usr/bin/env python
"""Error budget report of the rollout kernels against the reference goldens and the CPU oracle.

    python scripts/precision_report.py [impl ...]         (default: tc ffma)

For every golden rollout it prints the worst ratio error/tolerance for states (rtol 1e-4 + 1e-5) and log-density
(rtol 1e-3 + 1e-4) with the clip-flip samples (margin < 2e-5) reported apart, and the same for a T=1000 planner-sized
case (500 samples) against the oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch

import density_planner_b200 as dp
import dp_oracle as O


def report(tag, xe, rho, margin, xe_ref, rho_ref, margin_ref, xref):
    x, xr = xe + xref[None], xe_ref + xref[None]
    s_ratio = (np.abs(x - xr) / (1e-4 * np.abs(xr) + 1e-5)).max()
    err = np.abs(rho - rho_ref)
    tol = 1e-3 * np.abs(rho_ref) + 1e-4
    risky = np.minimum(margin, margin_ref) < 2e-5
    ratio = (err / tol).max(axis=1)
    safe_ratio = ratio[~risky].max() if (~risky).any() else 0.0
    rel = (err / (np.abs(rho_ref) + 1e-2))[~risky].max() if (~risky).any() else 0.0
    print("%-34s states err/tol %.3f | log-density err/tol %.3f (safe; rel %.2e, abs %.2e) | risky %d (worst %.1f)" % (
        tag, s_ratio, safe_ratio, rel, err[~risky].max(), int(risky.sum()), ratio[risky].max() if risky.any() else 0.0))


def main(impls):
    args = dp.default_args()
    car = dp.Car(args)
    blob = O.load_weights()
    gdir = os.path.join(ROOT, "tests", "golden")
    for impl in impls:
        for name in ("rollout_cfg1", "rollout_n2000_t300", "rollout_n4000_t100"):
            g = np.load(os.path.join(gdir, name + ".npz"))
            stride = int(g["stride"]) if "stride" in g.files else 1
            xe0 = torch.from_numpy(g["xe0"]).unsqueeze(-1).cuda()
            xref = torch.from_numpy(g["xref"]).unsqueeze(0).cuda()
            uref = torch.from_numpy(g["uref"]).unsqueeze(0).cuda()
            xe, rho, m = car.compute_density(xe0, xref, uref, float(g["dt"]), stride=stride, impl=impl, return_margin=True,
                                             cutting=True, log_density=True)
            report("%s %s" % (impl, name), xe.cpu().numpy(), rho[:, 0, :].cpu().numpy(), m.cpu().numpy(), g["xe_traj"],
                   g["rholog_traj"], g["clip_margin"], g["xref"][:, ::stride])
        for seed, N, T in ((21, 500, 1000), (22, 4096, 1000), (23, 20000, 100)):
            torch.manual_seed(seed)
            up, uref, xref = car.get_valid_ref(args)
            uref, xref = uref[:, :, :T], xref[:, :, :T + 1]
            xe0 = car.sample_xe0(N)
            xe, rho, m = car.compute_density(xe0.cuda(), xref, uref, args.dt_sim, log_density=True, impl=impl, return_margin=True)
            o_xe, o_rho, o_m, _ = O.rollout(blob, xe0[:, :, 0].numpy(), xref[0].numpy(), uref[0].numpy(), args.dt_sim,
                                            want_margin=True)
            report("%s oracle N=%d T=%d" % (impl, N, T), xe.cpu().numpy(), rho[:, 0, :].cpu().numpy(), m.cpu().numpy(), o_xe, o_rho,
                   o_m, xref[0].numpy())


if __name__ == "__main__":
    main(sys.argv[1:] or ["tc", "ffma"])
