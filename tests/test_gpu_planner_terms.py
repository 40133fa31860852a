"""GPU parity tests of the planner-side terms: the reference-trajectory integrator with
its analytic backward, and the goal / bounds cost terms, against goldens minted from the unmodified reference and against
the CPU oracle."""
import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu


def test_reference_integrator_forward_and_gradient(car, oracle, args):
    g = golden("planner_terms.npz")
    up = torch.from_numpy(g["int_up"]).cuda().requires_grad_(True)
    xref0 = torch.from_numpy(g["int_xref0"]).reshape(1, 5, 1).repeat(6, 1, 1).cuda()
    uref, xref = car.up2ref_traj(xref0, up, args, short=False)
    assert xref.is_cuda and xref.shape == (6, 5, 1001) and uref.shape == (6, 2, 999 + 0) or uref.shape == (6, 2, 1000)
    assert np.abs(xref.detach().cpu().numpy() - g["int_xref"]).max() < 2e-4
    w = torch.from_numpy(g["int_w"]).cuda()
    loss = (w * xref).sum() + 0.1 * (uref ** 2).sum()
    loss.backward()
    assert abs(float(loss) - float(g["int_loss"])) < 1e-4 * abs(float(g["int_loss"]))
    assert np.abs(up.grad.cpu().numpy() - g["int_gup"]).max() < 2e-4 * np.abs(g["int_gup"]).max()
    us, xs = car.up2ref_traj(xref0, up.detach(), args, short=True)
    assert np.abs(xs.cpu().numpy() - g["int_xref_short"]).max() < 2e-4 and (us.cpu().numpy() == g["int_uref_short"]).all()
    # shared start state (batch dimension 1) and its gradient
    x0 = torch.from_numpy(g["int_xref0"]).reshape(1, 5, 1).cuda().requires_grad_(True)
    u = torch.from_numpy(g["int_uref"]).cuda()
    xr = car.compute_xref_traj(x0, u, args)
    (w[:, :, :xr.shape[2]] * xr).sum().backward()
    _, gx0 = oracle.xref_traj_grad(g["int_xref"][:, :, :xr.shape[2]], g["int_w"][:, :, :xr.shape[2]], 0.01)
    assert np.abs(x0.grad[0, :, 0].cpu().numpy() - gx0.sum(axis=0)).max() < 2e-4 * np.abs(gx0.sum(axis=0)).max()


def test_reference_integrator_vs_oracle_batch(car, oracle, args):
    """100 trajectories (the planner's initialisation batch), bit-level agreement of the op order with the oracle."""
    rs = np.random.RandomState(5)
    B, T = 100, 1000
    uref = rs.uniform(-3, 3, (B, 2, T)).astype(np.float32)
    x0 = np.array([0.0, -25.0, 1.5, 3.0, 0.0], np.float32)
    got = car.compute_xref_traj(torch.from_numpy(x0).reshape(1, 5, 1).cuda(), torch.from_numpy(uref).cuda(), args)
    ref = oracle.xref_traj(x0, uref[:, :, :got.shape[2] - 1], 0.01)
    assert np.abs(got.cpu().numpy() - ref).max() < 5e-4


@pytest.mark.parametrize("ev", [0, 1])
@pytest.mark.parametrize("mx", [0, 1])
def test_goal_and_bounds_costs_vs_reference_golden(dp, car, args, ev, mx):
    g = golden("planner_terms.npz")
    tag = "e%d_m%d" % (ev, mx)
    x = torch.from_numpy(g["cost_x"]).cuda().requires_grad_(not mx)
    rho = torch.from_numpy(g["cost_rho"]).unsqueeze(1).cuda().requires_grad_(not mx)
    xrefN = torch.from_numpy(g["xrefN"]).reshape(1, 5, 1)
    cg, close = dp.get_cost_goal(x, rho, xrefN, args, evaluate=bool(ev), get_max=bool(mx))
    cb, inb = dp.get_cost_bounds(x, rho, car, evaluate=bool(ev), get_max=bool(mx))
    assert bool(close) == bool(g["close_" + tag]) and bool(inb) == bool(g["inb_" + tag])
    assert abs(float(cg) - float(g["goal_" + tag])) <= 1e-4 * abs(float(g["goal_" + tag]))
    assert cb.shape == (1,) and abs(float(cb) - float(g["bounds_" + tag][0])) <= 1e-4 * abs(float(g["bounds_" + tag][0]))
    if not mx:
        (cg.sum() + cb.sum()).backward()
        gx, gr = x.grad.cpu().numpy(), rho.grad[:, 0, :].cpu().numpy()
        assert np.abs(gx - g["gx_" + tag]).max() <= 1e-5 * np.abs(g["gx_" + tag]).max()
        assert np.abs(gr - g["grho_" + tag]).max() <= 1e-5 * np.abs(g["grho_" + tag]).max()


def test_bounds_cost_inside_and_native_layout(dp, car, args):
    g = golden("planner_terms.npz")
    x = torch.from_numpy(g["cost_x"]).cuda()
    rho = torch.from_numpy(g["cost_rho"]).unsqueeze(1).cuda()
    lo = car.X_MIN_MP[0, :, 0].numpy()
    hi = np.minimum(car.X_MAX_MP[0, :, 0].numpy(), 1e6)
    xin = torch.minimum(torch.maximum(x, torch.from_numpy(lo + 0.5).cuda().reshape(1, 5, 1)),
                        torch.from_numpy(hi - 0.5).cuda().reshape(1, 5, 1))
    cb, inb = dp.get_cost_bounds(xin, rho, car)
    assert float(cb) == 0.0 and bool(inb)
    # the kernel-native [t][d][n] storage (a permuted view) gives the same numbers as the reference layout
    xt = x.permute(2, 1, 0).contiguous().permute(2, 1, 0)
    rt = rho.permute(2, 1, 0).contiguous().permute(2, 1, 0)
    a, _ = dp.get_cost_bounds(x, rho, car)
    b, _ = dp.get_cost_bounds(xt, rt, car)
    assert abs(float(a) - float(b)) <= 1e-12 * abs(float(a))


def test_attach_binds_the_kernel_backed_calls(dp, oracle, args):
    """`dp.attach(planner)` on a stand-in for the reference's planner / ego objects (same attribute names): every bound
    call returns what the reference returned for the same inputs (goldens of env seed 0 and the planner terms)."""
    import types
    e = golden("env_seed0.npz")
    pt = golden("planner_terms.npz")
    gx, gy = oracle.env_gradient(e["grid"])
    env = types.SimpleNamespace(grid=torch.from_numpy(e["grid"]), grid_gradientX=torch.from_numpy(gx),
                                grid_gradientY=torch.from_numpy(gy))
    ref_system = types.SimpleNamespace()  # the reference's Car; only up2ref_traj is (re)bound on it
    ego = types.SimpleNamespace(env=env, args=args, system=ref_system,
                                xref0=torch.from_numpy(e["le_xref0"]).reshape(1, 5, 1),
                                xe0=torch.from_numpy(e["le_xe0"]).unsqueeze(-1),
                                rho0=torch.from_numpy(e["le_rho0"]).reshape(-1, 1, 1),
                                xrefN=torch.from_numpy(pt["xrefN"]).reshape(1, 5, 1),
                                predict_density=lambda *a, **k: (_ for _ in ()).throw(AssertionError("NN branch called")))
    planner = types.SimpleNamespace(ego=ego, get_cost_coll=None, get_cost_goal=lambda *a, **k: None, get_cost_bounds=None,
                                    get_cost_coll_initialize=lambda x: (_ for _ in ()).throw(AssertionError("reference init")))
    dp.attach(planner)
    up = torch.from_numpy(e["le_up"])
    uref_traj, xref_traj = ego.system.up2ref_traj(ego.xref0, up, args, short=True)
    assert uref_traj.device.type == "cpu" and xref_traj.shape == (1, 5, 101)
    x_traj, rho_traj = ego.predict_density(up, xref_traj, use_nn=False)
    assert np.abs(x_traj[:, :2, :].numpy() - e["le_x"]).max() < 2e-4
    c = planner.get_cost_coll(x_traj, rho_traj)
    assert abs(float(c) - e["le_cost"][0]) <= 2e-3 * abs(e["le_cost"][0])
    ci = planner.get_cost_coll_initialize(x_traj[:100])
    assert np.abs(ci.numpy() - e["le_cost_init"]).max() <= 1e-4 * np.abs(e["le_cost_init"]).max()
    xs = torch.from_numpy(pt["cost_x"])
    rs = torch.from_numpy(pt["cost_rho"]).unsqueeze(1)
    cg, close = planner.get_cost_goal(xs, rs, evaluate=False, get_max=False)
    cb, inb = planner.get_cost_bounds(xs, rs, evaluate=False, get_max=False)
    assert cg.device.type == "cpu" and abs(float(cg) - float(pt["goal_e0_m0"])) <= 1e-4 * float(pt["goal_e0_m0"])
    assert abs(float(cb) - float(pt["bounds_e0_m0"][0])) <= 1e-4 * float(pt["bounds_e0_m0"][0]) and not bool(inb)


def test_density_nn_on_gpu_matches_reference_golden(dp, oracle, args):
    """The density-NN branch on the fused tcgen05 MLP kernel (dp_dnn_forward / dp_dnn_backward, all 101 time points x 500
    samples in one launch) against the reference's 101-call loop: states, normalised density, loss and the gradient w.r.t.
    the input parameters -- the same golden and tolerances as the CPU cross-check in tests/test_abi_and_host.py."""
    import types
    g = golden("density_nn.npz")
    nn = dp.DensityNN(device="cuda")
    assert nn.backend == "kernel" and nn.handle
    pred = dp.EgoPredictor(types.SimpleNamespace(), args, torch.from_numpy(g["xref0"]).reshape(1, 5, 1),
                           torch.from_numpy(g["xe0"]).unsqueeze(-1), torch.from_numpy(g["rho0"]).reshape(-1, 1, 1), density_nn=nn)
    up = torch.from_numpy(g["up"]).requires_grad_(True)
    x_traj, rho_traj = pred.predict_density(up, torch.from_numpy(g["xref_short"]), use_nn=True)
    assert x_traj.device.type == "cpu" and np.abs(x_traj.detach().numpy() - g["x_traj"]).max() < 1e-4
    # normalised density: the contract allows rtol 1e-3 on the log-density; the split-fp16 kernel carries ~22 mantissa bits per
    # layer (measured: 1.2e-4 of the maximum, the plain-torch fp32 pass 0.4e-4), bound at 3e-4
    assert np.abs(rho_traj.detach()[:, 0, :].numpy() - g["rho_traj"]).max() < 3e-4 * g["rho_traj"].max()
    loss = (torch.from_numpy(g["w"]) * x_traj).sum() + 1e3 * (torch.from_numpy(g["wr"]).unsqueeze(1) * rho_traj).sum()
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) < 2e-4 * abs(float(g["loss"]))
    # the golden differentiates through the reference integrator as well (x_traj = xe_nn + xref(up)): that part comes from the
    # oracle's adjoint with the upstream gradient w summed over the samples (as in the CPU cross-check)
    uref_long = np.repeat(g["up"], 100, axis=2)[:, :, :1000]
    xref_long = oracle.xref_traj(g["xref0"], uref_long, 0.01)
    gx = np.zeros((1, 5, 1001))
    gx[0, :, ::10] = g["w"].sum(axis=0)
    gu, _ = oracle.xref_traj_grad(xref_long, gx, 0.01)
    total = up.grad.numpy() + gu.reshape(1, 2, 10, 100).sum(axis=3)
    assert np.abs(total - g["gup"]).max() < 2e-3 * np.abs(g["gup"]).max()


@pytest.mark.parametrize("activation", ["relu", "tanh"])
@pytest.mark.parametrize("dims", [(31, 6), (31, 150, 6), (20, 160, 33, 7), (31, 150, 150, 150, 150, 150, 150, 150, 6)])
def test_density_nn_kernel_vs_fp64(dp, dims, activation):
    """dp_dnn_forward / dp_dnn_backward on random ReLU and tanh MLPs of 1, 2, 3 and 8 layers (odd widths, ragged row counts that are not
    a multiple of the 128-row tile, more tiles than SMs) against the same network in float64: values to fp32-level accuracy
    (three split-fp16 products per layer: ~22 mantissa bits), gradients w.r.t. the rows likewise."""
    gen = torch.Generator().manual_seed(len(dims))
    L = len(dims) - 1
    z = {}
    for l in range(L):
        z["linear_list.%d.weight" % l] = (torch.randn(dims[l + 1], dims[l], generator=gen) / dims[l] ** 0.5).numpy()
        z["linear_list.%d.bias" % l] = (0.1 * torch.randn(dims[l + 1], generator=gen)).numpy()
    nn = dp.DensityNN(weights=z, device="cuda", activation=activation)
    assert nn.backend == "kernel"
    act = torch.relu if activation == "relu" else torch.tanh
    for rows in (1, 127, 128, 1000, 148 * 128 * 2 + 77):
        x = torch.randn(rows, dims[0], generator=gen)
        xc = x.cuda().requires_grad_(True)
        y = nn.mlp(xc)
        w = torch.randn(rows, dims[-1], generator=gen)
        (y * w.cuda()).sum().backward()
        xd = x.double().requires_grad_(True)
        h = xd
        for l in range(L):
            h = torch.addmm(torch.from_numpy(z["linear_list.%d.bias" % l]).double(), h, torch.from_numpy(z["linear_list.%d.weight" % l]).double().t())
            if l < L - 1:
                h = act(h)
        (h * w.double()).sum().backward()
        err = float((y.detach().cpu().double() - h.detach()).abs().max()) / float(h.detach().abs().max())
        # a pre-activation within rounding of zero may take the other ReLU branch in fp32: such rows (a handful in 40 M
        # pre-activations) differ by a whole path; bound their number, and check everything else at fp32 level
        grow = (xc.grad.cpu().double() - xd.grad).abs().max(dim=1).values / float(xd.grad.abs().max())
        if activation == "tanh":
            # the backward gate 1 - h^2 is formed from the fp16 image of h the forward pass keeps (11 bits on the gate; the
            # planner's gradient contract is 2e-3)
            assert err < 1e-5 and float(grow.max()) < 2e-3, (dims, rows, err, float(grow.max()))
            continue
        flipped = int((grow > 1e-5).sum())
        # measured on the 8-layer stack: 3.7e-6 (values), 4.4e-6 (gradients) of the maximum -- about 1.3e-6 per layer
        assert err < 1e-5 and flipped <= max(2, rows // 2000), (dims, rows, err, flipped, float(grow.max()))
    nn.close()


@pytest.mark.parametrize("seed", [0, 1])
def test_cost_coll_initialize_gradient_and_goal_with_long_density(dp, oracle, args, seed):
    """(1) get_cost_coll_initialize forward + analytic backward against the reference's autograd (golden minted from
    MotionPlannerGrad.get_cost_coll_initialize, :309-336, on the first 100 LE trajectories): the planner's initialisation
    phase differentiates this cost.  (2) the goal term when rho_traj has more time points than x_traj (the LE branch: 101
    states, 1001 densities): the reference weights the last state with rho_traj[:, 0, -1] (MotionPlanner.py:139-141)."""
    g = golden("cost_init_grad.npz")
    e = golden("env_seed%d.npz" % seed)
    gx, gy = oracle.env_gradient(e["grid"])
    cc = dp.CollisionCost(torch.from_numpy(e["grid"]), torch.from_numpy(gx), torch.from_numpy(gy), args)
    x5 = torch.zeros(100, 5, 101)
    x5[:, :2, :] = torch.from_numpy(e["le_x"][:100])
    x5.requires_grad_(True)
    c = cc.get_cost_coll_initialize(x5)
    assert c.shape == (100,) and c.requires_grad
    ref_c = g["s%d_cost" % seed]
    assert np.abs(c.detach().numpy() - ref_c).max() <= 1e-4 * np.abs(ref_c).max()
    w = torch.linspace(0.5, 1.5, 100)
    (w * c).sum().backward()
    ref_g = g["s%d_gx" % seed]
    assert float(x5.grad[:, 2:, :].abs().max()) == 0.0
    assert np.abs(x5.grad[:, :2, :].numpy() - ref_g).max() <= 2e-6 * np.abs(ref_g).max()
    assert ((x5.grad[:, :2, :].numpy() != 0) == (ref_g != 0)).all()
    # CUDA inputs in the kernel's own layout take the same path
    xc = x5.detach().cuda().requires_grad_(True)
    cc.get_cost_coll_initialize(xc).sum().backward()
    assert xc.grad.is_cuda and bool(torch.isfinite(xc.grad).all())
    # no gradient requested: forward only
    assert not cc.get_cost_coll_initialize(x5.detach()).requires_grad
    # (2) goal term, density longer than the states
    x_all = torch.zeros(500, 5, 101)
    x_all[:, :2, :] = torch.from_numpy(e["le_x"])
    x_all.requires_grad_(True)
    rho = torch.from_numpy(e["le_rho"]).unsqueeze(1).clone().requires_grad_(True)
    xrefN = torch.from_numpy(g["s%d_xrefN" % seed]).reshape(1, 5, 1)
    cg, close = dp.get_cost_goal(x_all, rho, xrefN, args, evaluate=True)
    ref_goal = float(g["s%d_goal" % seed])
    assert abs(float(cg) - ref_goal) <= 1e-4 * abs(ref_goal)
    cg.backward()
    assert np.abs(x_all.grad[:, :2, -1].numpy() - g["s%d_goal_gx_last" % seed]).max() <= 1e-5 * np.abs(g["s%d_goal_gx_last" % seed]).max()
    assert np.abs(rho.grad[:, 0, -1].numpy() - g["s%d_goal_grho_last" % seed]).max() <= 1e-5 * np.abs(g["s%d_goal_grho_last" % seed]).max()
    assert float(rho.grad[:, 0, 100].abs().max()) == 0.0 and float(x_all.grad[:, :, :-1].abs().max()) == 0.0
