#!/usr/bin/env python
"""Mint the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py [--only rollout,step,grid,env,data]

The reference has no tests and no golden vectors of its own, so parity is pinned
on outputs of the reference itself, generated here through oracle/ref_harness.py (stubs for the
plotting / pygame / casadi imports; nothing in the reference is modified).  Every file written here
is an input/output pair of a reference function that tests/ replays against the oracle restatement
(CPU, `-m "not gpu"`) and against the CUDA path (`-m gpu`).

Also exports the trained C3M controller weights (data/trained_controller/controller_CAR_3layers.pth.tar,
loaded by systems/sytem_CAR.py:113-125) to density_planner_b200/data/controller_CAR_3layers.npz so the
GPU box (which has no /root/reference) can run the real controller.
"""
import argparse
import io
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness as rh  # noqa: E402

torch.set_num_threads(1)  # single-thread BLAS: reproducible summation order for the goldens


def save(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v))
                                 for k, v in arrs.items()})
    print("wrote %s (%.1f KB)" % (name, os.path.getsize(path) / 1e3))


def export_weights(car):
    sd = car.controller.controller.state_dict()
    out = os.path.join(ROOT, "density_planner_b200", "data")
    os.makedirs(out, exist_ok=True)
    np.savez(os.path.join(out, "controller_CAR_3layers.npz"), **{k: v.numpy() for k, v in sd.items()})
    print("exported controller weights:", {k: tuple(v.shape) for k, v in sd.items()})


def u_raw_traj(car, x_traj, xref_traj, uref_traj):
    """Unclipped controller output along a stored trajectory (for the clip-flip audit)."""
    ctrl = car.controller
    out = torch.zeros(x_traj.shape[0], 2, uref_traj.shape[2])
    with torch.no_grad():
        for i in range(uref_traj.shape[2]):
            x = x_traj[:, :, [i]]
            xe = x[:, :4, :] - xref_traj[:, :4, [i]]
            xe[:, 2, :] += x[:, 4, :]
            out[:, :, [i]] = ctrl.controller(x[:, :4, :], xe, uref_traj[:, :, [i]])
    return out


def gen_rollout(car, args):
    # (1) config-1 shape: seed 0, get_valid_ref + sample_xe0(20), full horizon
    torch.manual_seed(0)
    with rh.ref_cwd():
        up, uref, xref = car.get_valid_ref(args)
        xe0 = car.sample_xe0(20)
        xe, rl = car.compute_density(xe0.clone(), xref, uref, args.dt_sim, cutting=True, log_density=True)
        ur = u_raw_traj(car, xe + xref, xref, uref)
    margin = (ur.abs() - 3).abs().amin(dim=(1, 2))
    save("rollout_cfg1.npz", xe0=xe0[:, :, 0], xref=xref[0], uref=uref[0], up=up, dt=np.float32(args.dt_sim),
         xe_traj=xe, rholog_traj=rl[:, 0, :], clip_margin=margin)

    # (2) N=2000,T=300 and N=4000,T=100 on other references; stored every 10th step
    for name, seed, N, T in (("rollout_n2000_t300", 1, 2000, 300), ("rollout_n4000_t100", 2, 4000, 100)):
        torch.manual_seed(seed)
        with rh.ref_cwd():
            up, uref, xref = car.get_valid_ref(args)
            uref, xref = uref[:, :, :T], xref[:, :, :T + 1]
            xe0 = car.sample_xe0(N)
            xe, rl = car.compute_density(xe0.clone(), xref, uref, args.dt_sim, cutting=True, log_density=True)
            ur = u_raw_traj(car, xe + xref, xref, uref)
        margin = (ur.abs() - 3).abs().amin(dim=(1, 2))
        clipped = (ur.abs() > 3).any(dim=1).float().mean().item()
        print(name, "fraction of (sample,step) with a clipped input: %.3f" % clipped)
        save(name + ".npz", xe0=xe0[:, :, 0], xref=xref[0], uref=uref[0], dt=np.float32(args.dt_sim),
             xe_traj=xe[:, :, ::10], rholog_traj=rl[:, 0, ::10], clip_margin=margin, stride=10)

    # (3) non-log density + rho0 given + no cutting (planner LE call shape, short)
    torch.manual_seed(3)
    with rh.ref_cwd():
        up, uref, xref = car.get_valid_ref(args)
        uref, xref = uref[:, :, :150], xref[:, :, :151]
        xe0 = car.sample_xe0(256)
        rho0 = torch.rand(256, 1, 1) + 0.5
        xe_a, rho_a = car.compute_density(xe0.clone(), xref, uref, args.dt_sim, rho0=rho0.clone(), cutting=True,
                                          log_density=False)
        xe_b, rho_b = car.compute_density(xe0.clone(), xref, uref, args.dt_sim, cutting=False, log_density=True,
                                          compute_density=True)
        xe_c, rho_c = car.compute_density(xe0.clone(), xref, uref, args.dt_sim, compute_density=False)
    assert rho_c is None
    save("rollout_modes.npz", xe0=xe0[:, :, 0], xref=xref[0], uref=uref[0], rho0=rho0[:, 0, 0],
         dt=np.float32(args.dt_sim), xe_lin=xe_a[:, :, ::5], rho_lin=rho_a[:, 0, ::5], xe_log=xe_b[:, :, ::5],
         rholog=rho_b[:, 0, ::5], xe_nodens=xe_c[:, :, ::5], stride=5)


def gen_step(car, args):
    """Single-step API on 4096 samples, incl. samples that straddle the +-3 clip boundary."""
    torch.manual_seed(4)
    N = 4096
    xref = torch.tensor([1.0, -2.0, 0.7, 4.0, 0.0]).reshape(1, 5, 1)
    with rh.ref_cwd():
        xe = torch.cat((car.sample_xe0(N // 2), car.sample_xe(N // 2)), 0)  # half/full error range: clipped and unclipped
    x = xe + xref
    uref = torch.tensor([0.8, -1.1]).reshape(1, 2, 1)
    with rh.ref_cwd():
        u = car.u_func(x, xref, uref)
        f = car.f_func(x, xref, uref)
        dudx = car.dudx_func(x.clone(), xref, uref).detach()
        dfdx = car.dfdx_func(x.clone(), xref, uref).detach()
        div = car.divf_func(x.clone(), xref, uref).detach()
        xn = car.get_next_x(x, xref, uref, args.dt_sim)
        rho = torch.rand(N) + 0.5
        rn = car.get_next_rho(x.clone(), xref, uref, rho, args.dt_sim).detach()
        rln = car.get_next_rholog(x.clone(), xref, uref, rho, args.dt_sim).detach()
        xe_c = x[:, :4, :] - xref[:, :4, :]
        xe_c[:, 2, :] += x[:, 4, :]
        uraw = car.controller.controller(x[:, :4, :], xe_c, uref).detach()
    print("step: clipped fraction u0 %.3f u1 %.3f" % ((uraw[:, 0].abs() > 3).float().mean(),
                                                       (uraw[:, 1].abs() > 3).float().mean()))
    save("step_api.npz", x=x[:, :, 0], xref=xref[0, :, 0], uref=uref[0, :, 0], dt=np.float32(args.dt_sim),
         u=u[:, :, 0], u_raw=uraw[:, :, 0], f=f[:, :, 0], dudx=dudx, dfdx=dfdx, div=div, x_next=xn[:, :, 0],
         rho=rho, rho_next=rn, rholog_next=rln)


def grid_inputs(n, seed):
    """Deterministic world positions; tests regenerate them with the same RandomState call sequence."""
    rs = np.random.RandomState(seed)
    px = rs.uniform(-13.0, 13.0, n).astype(np.float32)
    py = rs.uniform(-31.0, 11.0, n).astype(np.float32)
    return px, py


def gen_grid(car, args):
    rh.install()
    with rh.ref_cwd():
        from motion_planning.utils import pos2gridpos, gridpos2pos, pred2grid
        from systems.utils import get_density_map
    n = 400000
    px, py = grid_inputs(n, 11)
    # exact half-cell points and cell centres
    hx = (np.arange(-2, 484, dtype=np.float32) * np.float32(0.05) + np.float32(-12.0)).astype(np.float32)
    hy = (np.arange(-2, 804, dtype=np.float32) * np.float32(0.05) + np.float32(-30.0)).astype(np.float32)
    gx, gy = pos2gridpos(args, pos_x=torch.from_numpy(px), pos_y=torch.from_numpy(py))
    ghx, _ = pos2gridpos(args, pos_x=torch.from_numpy(hx))
    _, ghy = pos2gridpos(args, pos_y=torch.from_numpy(hy))
    # inverse map on float grid positions
    rs = np.random.RandomState(12)
    fgx = rs.uniform(-50, 300, 4096).astype(np.float32)
    fgy = rs.uniform(-50, 450, 4096).astype(np.float32)
    bx, by = gridpos2pos(args, pos_x=torch.from_numpy(fgx), pos_y=torch.from_numpy(fgy))
    save("pos2gridpos.npz", n=n, seed=11, gx=gx.numpy().astype(np.int16), gy=gy.numpy().astype(np.int16),
         hx=hx, hy=hy, ghx=ghx.numpy().astype(np.int16), ghy=ghy.numpy().astype(np.int16),
         fgx=fgx, fgy=fgy, bx=bx, by=by)

    # pred2grid: 10k samples around a point, rho positive
    rs = np.random.RandomState(13)
    N = 10000
    x = np.zeros((N, 5, 1), np.float32)
    x[:, 0, 0] = rs.normal(1.3, 0.8, N)
    x[:, 1, 0] = rs.normal(-12.0, 1.5, N)
    # some beyond the LOW grid edge (clamped to cell 0, utils.py:261-262).  Beyond the HIGH edge the reference
    # clamps to G (not G-1) and then fails on the slice assignment at :277, so that case has no golden.
    x[:50, 0, 0] = rs.uniform(-13.0, -11.5, 50)
    x[50:100, 1, 0] = rs.uniform(-31.0, -29.5, 50)
    rho = rs.uniform(0.0, 1.0, (N, 1, 1)).astype(np.float32)
    grid, gridpos = pred2grid(torch.from_numpy(x), torch.from_numpy(rho), args, return_gridpos=True)
    nz = grid[:, :, 0].nonzero()
    save("pred2grid.npz", x=x[:, :2, 0], rho=rho[:, 0, 0], nz=nz.numpy().astype(np.int16),
         val=grid[nz[:, 0], nz[:, 1], 0], gridpos=gridpos.numpy().astype(np.int16), shape=np.array(grid.shape))

    # get_density_map: error-state histogram, log density
    rs = np.random.RandomState(14)
    N = 20000
    xe = rs.normal(0, 0.6, (N, 2)).astype(np.float32)
    rl = rs.normal(0, 3.0, N).astype(np.float32)
    dm, rb = get_density_map(torch.from_numpy(xe.copy()), torch.from_numpy(rl.copy()), args, log_density=True,
                             type="LE", system=car)
    dm_lin, _ = get_density_map(torch.from_numpy(xe.copy()), torch.from_numpy(np.exp(rl)), args, log_density=False,
                                type="LE", system=car)
    dm_mc, _ = get_density_map(torch.from_numpy(xe.copy()), torch.ones(N), args, log_density=False, type="MC",
                               system=car)
    save("density_map.npz", xe=xe, rholog=rl, dm=dm, range_bins=np.array([[float(a), float(b)] for a, b in rb]),
         dm_lin=dm_lin, dm_mc=dm_mc)


def synth_samples(seed, N, Ts, xy_ref):
    """Synthetic weighted sample set around a reference path (numpy RandomState: reproducible in tests)."""
    rs = np.random.RandomState(1000 + seed)
    idx = rs.randint(0, xy_ref.shape[0], N)
    xy = xy_ref[idx] + (0.3 * rs.standard_normal((N, 2, Ts))).astype(np.float32)
    rho = rs.uniform(0, 1, (N, Ts)).astype(np.float32)
    rho /= rho.sum(axis=0, keepdims=True)
    return xy.astype(np.float32), rho.astype(np.float32)


def gen_env(car, args, seeds):
    rh.install()
    with rh.ref_cwd():
        from motion_planning.example_objects import create_mp_task
        from motion_planning.MotionPlannerGrad import MotionPlannerGrad
    os.makedirs("/tmp/dp_log", exist_ok=True)
    for seed in seeds:
        torch.manual_seed(seed)
        np.random.seed(seed)
        with rh.ref_cwd():
            ego = create_mp_task(args, seed)
            planner = MotionPlannerGrad(ego, path_log="/tmp/dp_log/")
        env = ego.env
        out = dict(grid=env.grid.numpy(), seed=seed)
        # sparse check values of the gradient grids (the oracle rebuilds them from `grid`)
        rs = np.random.RandomState(77 + seed)
        ci = np.stack([rs.randint(0, 241, 20000), rs.randint(0, 401, 20000), rs.randint(0, 101, 20000)], 1)
        out["chk_idx"] = ci.astype(np.int16)
        out["chk_gx"] = env.grid_gradientX.numpy()[ci[:, 0], ci[:, 1], ci[:, 2]]
        out["chk_gy"] = env.grid_gradientY.numpy()[ci[:, 0], ci[:, 1], ci[:, 2]]
        out["gx_sum"] = np.float64(env.grid_gradientX.double().sum())
        out["gy_sum"] = np.float64(env.grid_gradientY.double().sum())

        # LE sample set (the planner's own call): seed 0 and 1 only (8 s each)
        if seed < 2:
            up = 0.5 * torch.randn(1, 2, 10)
            with rh.ref_cwd():
                uref_traj, xref_traj, x_traj, rho_traj = planner.get_traj(up, compute_density=True, use_nn=False,
                                                                         plot=False)
            x_le = x_traj.clone().requires_grad_(True)
            r_le = rho_traj.clone().requires_grad_(True)
            c = planner.get_cost_coll(x_le, r_le)
            c.sum().backward()
            cmax = planner.get_cost_coll(x_traj, rho_traj, get_max=True)
            out.update(le_up=up, le_xe0=ego.xe0[:, :, 0], le_rho0=ego.rho0[:, 0, 0], le_xref0=ego.xref0[0, :, 0],
                       le_x=x_traj[:, :2, :], le_rho=rho_traj[:, 0, :], le_cost=c.detach(), le_cost_max=cmax,
                       le_gx=x_le.grad[:, :2, :], le_grho=r_le.grad[:, 0, :101])
            ci_ = planner.get_cost_coll_initialize(x_traj[:100])
            out["le_cost_init"] = ci_

        # synthetic weighted samples around a straight reference through the street
        Ts = 101
        tt = np.linspace(0, 1, Ts, dtype=np.float32)
        path = np.stack([0.5 * np.sin(3 * tt), -28 + 36 * tt], 0)[None].astype(np.float32)  # (1,2,Ts)
        xy, rho = synth_samples(seed, 20000, Ts, path)
        x5 = torch.zeros(xy.shape[0], 5, Ts)
        x5[:, :2, :] = torch.from_numpy(xy)
        x5.requires_grad_(True)
        r = torch.from_numpy(rho).unsqueeze(1).clone().requires_grad_(True)
        c = planner.get_cost_coll(x5, r)
        c.sum().backward()
        cmax = planner.get_cost_coll(x5.detach(), r.detach(), get_max=True)
        # per-slice costs (float64 accumulate of the reference's per-sample terms is not available; keep total)
        out.update(syn_cost=c.detach(), syn_cost_max=cmax, syn_gx_sum=x5.grad[:, :2, :].double().sum(dim=(0,)).numpy(),
                   syn_gx_abs=np.float64(x5.grad.abs().double().sum()), syn_grho_sum=r.grad.double().sum(dim=0).numpy()[0],
                   syn_gx_head=x5.grad[:64, :2, :], syn_grho_head=r.grad[:64, 0, :])
        save("env_seed%d.npz" % seed, **out)


def gen_planner_terms(car, args):
    """the reference-trajectory integrator up2ref_traj (ControlAffineSystem.py:341-379)
    with its gradient w.r.t. the input parameters, and the goal / bounds cost terms of MotionPlanner.get_cost
    (MotionPlanner.py:121-190) with their gradients."""
    rh.install()
    with rh.ref_cwd():
        from motion_planning.example_objects import create_mp_task
        from motion_planning.MotionPlannerGrad import MotionPlannerGrad
    os.makedirs("/tmp/dp_log", exist_ok=True)
    torch.manual_seed(0)
    np.random.seed(0)
    with rh.ref_cwd():
        ego = create_mp_task(args, 0)
        planner = MotionPlannerGrad(ego, path_log="/tmp/dp_log/")
    out = {}
    # (1) integrator: 6 parameter sets through the planner's own call (xref0 repeated per trajectory)
    g = torch.Generator().manual_seed(31)
    up = (1.5 * torch.randn(6, 2, 10, generator=g)).requires_grad_(True)
    xref0 = ego.xref0.repeat(6, 1, 1)
    uref_l, xref_l = car.up2ref_traj(xref0, up, args, short=False)
    w = torch.randn(xref_l.shape, generator=g)
    loss = (w * xref_l).sum() + 0.1 * (uref_l ** 2).sum()
    loss.backward()
    uref_s, xref_s = car.up2ref_traj(xref0, up.detach(), args, short=True)
    out.update(int_up=up.detach(), int_xref0=ego.xref0[0, :, 0], int_uref=uref_l.detach(), int_xref=xref_l.detach(),
               int_w=w, int_gup=up.grad.clone(), int_xref_short=xref_s, int_uref_short=uref_s, int_loss=loss.detach())
    # (2) goal / bounds costs on a synthetic weighted sample set with excursions beyond the planner's state bounds
    rs = np.random.RandomState(41)
    N, Ts = 3000, 101
    x = np.zeros((N, 5, Ts), np.float32)
    tt = np.linspace(0, 1, Ts, dtype=np.float32)
    x[:, 0, :] = 0.5 * np.sin(3 * tt)[None] + rs.normal(0, 4.0, (N, 1)) * tt[None] + rs.normal(0, 0.2, (N, Ts))
    x[:, 1, :] = (-28 + 40 * tt)[None] + rs.normal(0, 0.5, (N, Ts))
    x[:, 2, :] = 1.5 + rs.normal(0, 1.2, (N, Ts))
    x[:, 3, :] = 5.0 + rs.normal(0, 2.5, (N, Ts))
    x[:, 4, :] = rs.normal(0, 0.1, (N, 1))
    rho = rs.uniform(0, 1, (N, 1, Ts)).astype(np.float32)
    rho /= rho.sum(axis=0, keepdims=True)
    out.update(cost_x=x, cost_rho=rho[:, 0, :], xrefN=ego.xrefN[0, :, 0], x_min=car.X_MIN[0, :, 0], x_max=car.X_MAX[0, :, 0],
               x_min_mp=car.X_MIN_MP[0, :, 0], x_max_mp=car.X_MAX_MP[0, :, 0],
               close2goal_thr=np.float32(args.close2goal_thr), weight_goal_far=np.float32(args.weight_goal_far))
    for evaluate in (False, True):
        for get_max in (False, True):
            xt = torch.from_numpy(x).clone().requires_grad_(True)
            rt = torch.from_numpy(rho).clone().requires_grad_(True)
            tag = "e%d_m%d" % (int(evaluate), int(get_max))
            if get_max:  # evaluation-only variant (the reference's in-place re-weighting of a max output has no backward)
                with torch.no_grad():
                    cg, close = planner.get_cost_goal(xt, rt, evaluate=evaluate, get_max=True)
                    cb, inb = planner.get_cost_bounds(xt, rt, evaluate=evaluate, get_max=True)
            else:
                cg, close = planner.get_cost_goal(xt, rt, evaluate=evaluate, get_max=False)
                cb, inb = planner.get_cost_bounds(xt, rt, evaluate=evaluate, get_max=False)
                (cg.sum() + cb.sum()).backward()
                out.update({"gx_" + tag: xt.grad.clone(), "grho_" + tag: rt.grad[:, 0, :].clone()})
            out.update({"goal_" + tag: cg.detach(), "close_" + tag: np.array(bool(close)), "bounds_" + tag: cb.detach(),
                        "inb_" + tag: np.array(bool(inb))})
    # all samples inside the bounds: zero cost, in_bounds True
    xin = np.clip(x, car.X_MIN_MP[0, :, 0].numpy()[None, :, None] + 0.5, np.minimum(car.X_MAX_MP[0, :, 0].numpy(), 1e6)[None, :, None] - 0.5)
    cb, inb = planner.get_cost_bounds(torch.from_numpy(xin), torch.from_numpy(rho))
    out.update(bounds_inside=cb, inb_inside=np.array(bool(inb)))
    save("planner_terms.npz", **out)


def gen_density_nn(car, args):
    """the density-NN branch of EgoVehicle.predict_density (simulation_objects.py:280-299,
    density_training/utils.py:161-180): exports the default checkpoint's weights and mints outputs + gradient w.r.t. up."""
    rh.install()
    with rh.ref_cwd():
        from motion_planning.example_objects import create_mp_task
        sd, _, _ = torch.load(args.name_pretrained_nn, map_location="cpu", weights_only=False)
    out_dir = os.path.join(ROOT, "density_planner_b200", "data")
    np.savez(os.path.join(out_dir, "density_nn_default.npz"), **{k: v.numpy() for k, v in sd.items()})
    torch.manual_seed(0)
    np.random.seed(0)
    with rh.ref_cwd():
        ego = create_mp_task(args, 0)
    g = torch.Generator().manual_seed(51)
    up = (0.8 * torch.randn(1, 2, 10, generator=g)).requires_grad_(True)
    uref_traj, xref_traj = car.up2ref_traj(ego.xref0, up, args, short=True)
    with rh.ref_cwd():
        x_traj, rho_traj = ego.predict_density(up, xref_traj, use_nn=True)
    w = torch.randn(x_traj.shape, generator=g)
    wr = torch.randn(rho_traj.shape, generator=g)
    loss = (w * x_traj).sum() + 1e3 * (wr * rho_traj).sum()
    loss.backward()
    save("density_nn.npz", up=up.detach(), xref0=ego.xref0[0, :, 0], xe0=ego.xe0[:, :, 0], rho0=ego.rho0[:, 0, 0],
         xref_short=xref_traj.detach(), x_traj=x_traj.detach(), rho_traj=rho_traj.detach()[:, 0, :], w=w, wr=wr[:, 0, :],
         gup=up.grad.clone(), loss=loss.detach())


def gen_data(car, args):
    """compute_data (data_generation/compute_rawdata.py:9) with a fixed seed: RNG-order golden (config 1)."""
    rh.install()
    with rh.ref_cwd():
        from data_generation.compute_rawdata import compute_data
        torch.manual_seed(5)
        res = compute_data([1, 2], 20, car, args, samples_t=10, save=False, plot=False)
    out = {}
    for i, r in enumerate(res):
        for k, v in r.items():
            out["r%d_%s" % (i, k)] = v
    save("compute_data.npz", **out)


def gen_cost_init_grad(car, args):
    """MotionPlannerGrad.get_cost_coll_initialize (MotionPlannerGrad.py:309-336) differentiated w.r.t. the trajectories, as
    the planner's initialisation phase does (:138): per-trajectory costs and the gradient of a weighted sum, environments 0, 1,
    inputs = the first 100 LE trajectories of env_seed{0,1}.npz; and the goal term with a density LONGER than the states
    (MotionPlanner.py:139-141 uses rho_traj[:, 0, -1]: the LE branch returns 101 states and 1001 densities)."""
    rh.install()
    with rh.ref_cwd():
        from motion_planning.example_objects import create_mp_task
        from motion_planning.MotionPlannerGrad import MotionPlannerGrad
    os.makedirs("/tmp/dp_log", exist_ok=True)
    out = {}
    for seed in (0, 1):
        torch.manual_seed(seed)
        np.random.seed(seed)
        with rh.ref_cwd():
            ego = create_mp_task(args, seed)
            planner = MotionPlannerGrad(ego, path_log="/tmp/dp_log/")
        e = np.load(os.path.join(HERE, "env_seed%d.npz" % seed))
        x5 = torch.zeros(100, 5, 101)
        x5[:, :2, :] = torch.from_numpy(e["le_x"][:100])
        x5.requires_grad_(True)
        c = planner.get_cost_coll_initialize(x5)
        w = torch.linspace(0.5, 1.5, 100)
        (w * c).sum().backward()
        out["s%d_cost" % seed] = c.detach().numpy()
        out["s%d_gx" % seed] = x5.grad[:, :2, :].numpy()
        # goal term, density with 1001 time points against states with 101
        x_all = torch.zeros(500, 5, 101)
        x_all[:, :2, :] = torch.from_numpy(e["le_x"])
        x_all.requires_grad_(True)
        rho = torch.from_numpy(e["le_rho"]).unsqueeze(1).clone().requires_grad_(True)
        cg, close = planner.get_cost_goal(x_all, rho, evaluate=True)
        cg.backward()
        out["s%d_goal" % seed] = cg.detach().numpy()
        out["s%d_goal_gx_last" % seed] = x_all.grad[:, :2, -1].numpy()
        out["s%d_goal_grho_last" % seed] = rho.grad[:, 0, -1].numpy()
        out["s%d_goal_grho_at100" % seed] = rho.grad[:, 0, 100].numpy()
        out["s%d_xrefN" % seed] = ego.xrefN[0, :, 0].numpy()
    save("cost_init_grad.npz", **out)


def gen_plan_motion(car, args, seeds=(0, 1)):
    """BASELINE config 4: the reference's gradient planner, UNMODIFIED and un-attached, on the artificial environments
    `seeds` (plan_motion.py:82-90 seeds torch / numpy the same way): MotionPlannerGrad.find_initial_traj (100 random
    trajectories x 100 epochs on get_cost_initialize, :56-96), optimize_traj (100 epochs with the density NN, :98-160) and
    validate_ref (the LE rollout of 500 samples x 1000 steps, :366-397).  Stored: per-epoch cost traces, the best initial
    parameters, the final parameters and the cost dictionaries -- what tests/test_gpu_planner_real.py compares the
    attached planner with.  About four minutes per seed on one core."""
    rh.install()
    with rh.ref_cwd():
        from motion_planning.example_objects import create_mp_task
        from motion_planning.MotionPlannerGrad import MotionPlannerGrad
    os.makedirs("/tmp/dp_log", exist_ok=True)
    for seed in seeds:
        t0 = time.time()
        torch.manual_seed(seed)
        np.random.seed(seed)
        with rh.ref_cwd():
            ego = create_mp_task(args, seed)
            planner = MotionPlannerGrad(ego, path_log="/tmp/dp_log/")
            up_best, cost_min = planner.find_initial_traj()
            init_costs = planner.initial_traj[1]
            up, cost = planner.optimize_traj(up_best)
            opt_costs = planner.improved_traj[-1][1]
            val = planner.validate_ref(up)
        out = {"up_init_all": planner.initial_traj[0], "up_best": up_best, "cost_min": cost_min, "up_final": up,
               "xref0": ego.xref0, "xrefN": ego.xrefN}
        for k, v in init_costs.items():
            out["init_" + k] = torch.stack([c.detach() for c in v])          # (epochs, num_traj)
        for k, v in opt_costs.items():
            out["opt_" + k] = torch.stack([c.detach().reshape(()) for c in v])  # (epochs,)
        for k, v in cost.items():
            out["final_" + k] = v.detach()
        for k, v in val.items():
            out["val_" + k] = v.detach()
        save("plan_motion_seed%d.npz" % seed, **out)
        print("seed %d: %.0f s, final %s, validated %s" % (seed, time.time() - t0, {k: float(v) for k, v in cost.items()},
                                                          {k: float(v) for k, v in val.items()}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="weights,rollout,step,grid,env,data,planner,nn,initgrad,plan")
    ap.add_argument("--env-seeds", default="0,1,2,3,4,5,6,7,8,9")
    a = ap.parse_args()
    which = a.only.split(",")
    car, args = rh.make_car()
    t0 = time.time()
    if "weights" in which:
        export_weights(car)
    if "rollout" in which:
        gen_rollout(car, args)
    if "step" in which:
        gen_step(car, args)
    if "grid" in which:
        gen_grid(car, args)
    if "data" in which:
        gen_data(car, args)
    if "env" in which:
        gen_env(car, args, [int(s) for s in a.env_seeds.split(",")])
    if "planner" in which:
        gen_planner_terms(car, args)
    if "nn" in which:
        gen_density_nn(car, args)
    if "initgrad" in which:
        gen_cost_init_grad(car, args)
    if "plan" in which:
        gen_plan_motion(car, args)
    print("done in %.1f s" % (time.time() - t0))


if __name__ == "__main__":
    main()
