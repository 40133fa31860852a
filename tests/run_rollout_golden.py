"""Run one golden rollout case in a fresh process: python tests/run_rollout_golden.py <golden name> <impl>.

Kernel variants are picked once per process from environment knobs (DP_TC_BOUNDED, DP_TC_VARIANT), so the tests that
cover the non-default variants go through this script.  Exit code 0 = within the tolerances of conftest.py."""
import sys

import numpy as np
import torch

import conftest  # noqa: F401  (puts the repo root and tests/golden on sys.path)
from conftest import golden, assert_states_close, assert_rho_close


def main(name, impl):
    import density_planner_b200 as dp
    car = dp.Car(dp.default_args())
    g = golden(name + ".npz")
    stride = int(g["stride"]) if "stride" in g.files else 1
    dev = torch.device("cuda")
    xe0 = torch.from_numpy(g["xe0"]).unsqueeze(-1).to(dev)
    xref = torch.from_numpy(g["xref"]).unsqueeze(0).to(dev)
    uref = torch.from_numpy(g["uref"]).unsqueeze(0).to(dev)
    xe, rho, margin = car.compute_density(xe0, xref, uref, float(g["dt"]), stride=stride, impl=impl, return_margin=True,
                                          cutting=True, log_density=True)
    xe, rho, margin = xe.cpu().numpy(), rho[:, 0, :].cpu().numpy(), margin.cpu().numpy()
    assert_states_close(xe, g["xe_traj"], g["xref"][:, ::stride], name)
    assert_rho_close(rho, g["rholog_traj"], np.minimum(margin, g["clip_margin"]), name)
    print("ok", name, impl)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
