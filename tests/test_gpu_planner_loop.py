"""Config 4 (gradient optimisation loop with the kernel-backed density cost) on the 10 artificial environments: the
differentiable chain up -> reference integrator -> density prediction -> cost terms runs on the kernels end to end.  The
planner's own loop stays with the reference (it runs unchanged through `dp.attach`); here the same chain is driven by a
plain gradient step to check that it is usable for optimisation: finite gradients, a directional-derivative check of the
smooth terms, and a decreasing cost."""
import types

import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu


def _setup(dp, oracle, args, seed):
    e = golden("env_seed%d.npz" % seed)
    pt = golden("planner_terms.npz")
    nn_g = golden("density_nn.npz")
    gx, gy = oracle.env_gradient(e["grid"])
    cc = dp.CollisionCost(torch.from_numpy(e["grid"]), torch.from_numpy(gx), torch.from_numpy(gy), args)
    car = dp.Car(args)
    pred = dp.EgoPredictor(car, args, torch.from_numpy(nn_g["xref0"]).reshape(1, 5, 1), torch.from_numpy(nn_g["xe0"]).unsqueeze(-1),
                           torch.from_numpy(nn_g["rho0"]).reshape(-1, 1, 1), density_nn=dp.DensityNN(device="cuda"))
    xrefN = torch.from_numpy(pt["xrefN"]).reshape(1, 5, 1)
    return car, cc, pred, xrefN


def _cost(dp, car, cc, pred, xrefN, args, up, with_coll=True):
    uref, xref = car.up2ref_traj(pred.xref0.cuda(), up, args, short=True)
    x_traj, rho_traj = pred.predict_density(up, xref, use_nn=True)
    goal, _ = dp.get_cost_goal(x_traj, rho_traj, xrefN, args)
    bounds, _ = dp.get_cost_bounds(x_traj, rho_traj, car)
    cost = args.weight_goal * goal + args.weight_bounds * bounds.sum() + args.weight_uref * args.weight_uref_effort * (uref ** 2).sum()
    if with_coll:
        cost = cost + args.weight_coll * cc.get_cost_coll(x_traj, rho_traj).sum()
    return cost


@pytest.mark.parametrize("seed", list(range(10)))
def test_gradient_steps_on_the_kernel_backed_cost(dp, oracle, args, seed):
    car, cc, pred, xrefN = _setup(dp, oracle, args, seed)
    g = torch.Generator().manual_seed(100 + seed)
    up = (0.5 * torch.randn(1, 2, 10, generator=g)).cuda().requires_grad_(True)
    costs = []
    for it in range(6):
        cost = _cost(dp, car, cc, pred, xrefN, args, up)
        (grad,) = torch.autograd.grad(cost, up)
        assert bool(torch.isfinite(cost)) and bool(torch.isfinite(grad).all()) and float(grad.abs().max()) > 0
        costs.append(float(cost))
        with torch.no_grad():
            up -= 0.02 * grad / grad.abs().max()  # normalised step, like the reference's clamp by max_gradient
    assert min(costs[1:]) < costs[0]


def test_directional_derivative_of_the_smooth_terms(dp, oracle, args):
    """goal + bounds + control effort are smooth in up: autograd through integrator adjoint, batched NN and the cost
    kernels' analytic backward agrees with a central finite difference."""
    car, cc, pred, xrefN = _setup(dp, oracle, args, 0)
    g = torch.Generator().manual_seed(7)
    up = (0.5 * torch.randn(1, 2, 10, generator=g)).cuda().requires_grad_(True)
    cost = _cost(dp, car, cc, pred, xrefN, args, up, with_coll=False)
    (grad,) = torch.autograd.grad(cost, up)
    d = torch.randn(1, 2, 10, generator=g).cuda()
    d = d / d.norm()
    eps = 2e-2
    with torch.no_grad():
        cp = _cost(dp, car, cc, pred, xrefN, args, up + eps * d, with_coll=False)
        cm = _cost(dp, car, cc, pred, xrefN, args, up - eps * d, with_coll=False)
    fd = float(cp - cm) / (2 * eps)
    an = float((grad * d).sum())
    assert abs(fd - an) <= 0.05 * max(abs(fd), abs(an)) + 1e-5, (fd, an)
