"""GPU parity tests of the grid transforms, the collision cost (forward + backward), binning and the normaliser."""
import numpy as np
import pytest
import torch

from conftest import golden, COST_RTOL, RHO_RTOL, RHO_ATOL

pytestmark = pytest.mark.gpu


def test_pos2gridpos_bit_exact(dp, args):
    from make_golden import grid_inputs
    g = golden("pos2gridpos.npz")
    px, py = grid_inputs(int(g["n"]), int(g["seed"]))
    gx, gy = dp.pos2gridpos(args, pos_x=torch.from_numpy(px).cuda(), pos_y=torch.from_numpy(py).cuda())
    assert gx.dtype == torch.int64 and gx.is_cuda
    assert (gx.cpu().numpy() == g["gx"]).all() and (gy.cpu().numpy() == g["gy"]).all()
    ghx, _ = dp.pos2gridpos(args, pos_x=torch.from_numpy(g["hx"]))
    _, ghy = dp.pos2gridpos(args, pos_y=torch.from_numpy(g["hy"]))
    assert (ghx.numpy() == g["ghx"]).all() and (ghy.numpy() == g["ghy"]).all()
    bx, by = dp.gridpos2pos(args, pos_x=torch.from_numpy(g["fgx"]), pos_y=torch.from_numpy(g["fgy"]))
    assert (bx.numpy() == g["bx"]).all() and (by.numpy() == g["by"]).all()
    lx, _ = dp.pos2gridpos(args, pos_x=[-12.0, 0.0, 12.0])
    assert lx.tolist() == [0, 120, 240]


def test_pos2gridpos_5m_vs_oracle(dp, oracle, args):
    rs = np.random.RandomState(5)
    px = rs.uniform(-13, 13, 5_000_000).astype(np.float32)
    py = rs.uniform(-31, 11, 5_000_000).astype(np.float32)
    gx, gy = dp.pos2gridpos(args, pos_x=torch.from_numpy(px).cuda(), pos_y=torch.from_numpy(py).cuda())
    ox, oy = oracle.pos2gridpos(px, py)
    assert (gx.cpu().numpy() == ox).all() and (gy.cpu().numpy() == oy).all()


def test_pred2grid_and_density_map(dp, car, args):
    g = golden("pred2grid.npz")
    x = torch.zeros(len(g["rho"]), 5, 1)
    x[:, :2, 0] = torch.from_numpy(g["x"])
    rho = torch.from_numpy(g["rho"]).reshape(-1, 1, 1)
    grid, gridpos = dp.pred2grid(x, rho, args, return_gridpos=True)
    ref = np.zeros((241, 401), np.float32)
    ref[g["nz"][:, 0], g["nz"][:, 1]] = g["val"]
    assert grid.shape == (241, 401, 1)
    assert (gridpos.numpy() == g["gridpos"]).all()
    got = grid[:, :, 0].numpy()
    assert ((got > 0) == (ref > 0)).all()
    assert np.allclose(got, ref, rtol=1e-5, atol=1e-10)
    # beyond the upper edge the reference raises; so do we
    xb = x.clone()
    xb[0, 0, 0] = 12.6
    with pytest.raises(RuntimeError):
        dp.pred2grid(xb, rho, args)
    d = golden("density_map.npz")
    dm, rb = dp.get_density_map(torch.from_numpy(d["xe"]), torch.from_numpy(d["rholog"].copy()), args,
                                log_density=True, type="LE", system=car)
    assert np.allclose(dm, d["dm"], rtol=1e-5, atol=1e-9)
    assert np.allclose(np.array(rb), d["range_bins"])
    dm, _ = dp.get_density_map(torch.from_numpy(d["xe"]), torch.ones(len(d["xe"])), args, type="MC", system=car)
    assert np.allclose(dm, d["dm_mc"], rtol=1e-12)


def test_binning_counts_exact_and_order_free(dp, oracle, args):
    """Per-cell sample counts are bit-exact vs the oracle; int64 mass is identical under any permutation/sharding."""
    from density_planner_b200 import _lib, grid as G
    rs = np.random.RandomState(3)
    N = 300_000
    x = rs.normal(0.5, 1.0, N).astype(np.float32)
    y = rs.normal(-10, 3.0, N).astype(np.float32)
    w = rs.uniform(0, 1, N).astype(np.float32)
    spec = _lib.BinSpec()
    spec.mode = 0
    spec.geom = G._geom(args)
    spec.mass_scale_log2 = G.mass_scale_log2(N, 1.0)
    dev = torch.device("cuda", torch.cuda.current_device())
    xt, yt, wt = (torch.from_numpy(a).cuda() for a in (x, y, w))
    mass, cnt = G.bin2d_raw(spec, xt, yt, wt, N, 1, (1, 0), (1, 0), dev)
    _, _, ocnt = oracle.pred2grid(x, y, w)
    assert (cnt[0, :241, :401].cpu().numpy() == ocnt).all()
    perm = torch.randperm(N).cuda()
    mass2, cnt2 = G.bin2d_raw(spec, xt[perm].contiguous(), yt[perm].contiguous(), wt[perm].contiguous(), N, 1, (1, 0),
                              (1, 0), dev)
    assert bool((mass == mass2).all()) and bool((cnt == cnt2).all())
    h = N // 3
    ma, ca = G.bin2d_raw(spec, xt[:h].contiguous(), yt[:h].contiguous(), wt[:h].contiguous(), h, 1, (1, 0), (1, 0), dev)
    mb, cb = G.bin2d_raw(spec, xt[h:].contiguous(), yt[h:].contiguous(), wt[h:].contiguous(), N - h, 1, (1, 0), (1, 0), dev)
    assert bool((ma + mb == mass).all()) and bool((ca + cb == cnt).all())
    # mass against a float64 accumulation
    ref = np.zeros((242, 402))
    gx, gy = oracle.pos2gridpos(x, y)
    np.add.at(ref, (np.clip(gx, 0, 241), np.clip(gy, 0, 401)), w.astype(np.float64))
    got = mass[0].cpu().numpy() / 2.0 ** spec.mass_scale_log2
    assert np.abs(got - ref).max() < 1e-6


@pytest.mark.parametrize("N", [1, 3, 4, 5, 131, 4099])
def test_binning_small_and_stray_samples(dp, oracle, args, N):
    """The lean binning kernel on tiny, ragged sample counts (fewer samples than one vector load, tails of one to three) with
    samples far outside the grid on either side (clamped to the border cells 0 / G like pred2grid) and exactly on cell
    boundaries: counts bit-exact against the oracle, mass against a float64 accumulation."""
    from density_planner_b200 import _lib, grid as G
    rs = np.random.RandomState(N)
    x = rs.normal(0.0, 2.0, N).astype(np.float32)
    y = rs.normal(-5.0, 6.0, N).astype(np.float32)
    x[::7] = np.float32(1e6)
    y[::5] = np.float32(-1e6)
    x[1::11] = np.float32(-12.0) + np.float32(0.05)   # half-cell points of the 0.1 m grid
    w = rs.uniform(0, 1, N).astype(np.float32)
    spec = _lib.BinSpec()
    spec.mode = 0
    spec.geom = G._geom(args)
    spec.mass_scale_log2 = G.mass_scale_log2(N, 1.0)
    dev = torch.device("cuda", torch.cuda.current_device())
    xt, yt, wt = (torch.from_numpy(a).cuda() for a in (x, y, w))
    mass, cnt = G.bin2d_raw(spec, xt, yt, wt, N, 1, (1, 0), (1, 0), dev)
    assert int(cnt.sum()) == N
    gx, gy = oracle.pos2gridpos(x, y)
    ref_c = np.zeros((242, 402), np.int64)
    ref_m = np.zeros((242, 402))
    np.add.at(ref_c, (np.clip(gx, 0, 241), np.clip(gy, 0, 401)), 1)
    np.add.at(ref_m, (np.clip(gx, 0, 241), np.clip(gy, 0, 401)), w.astype(np.float64))
    assert (cnt[0].cpu().numpy() == ref_c).all()
    assert np.abs(mass[0].cpu().numpy() / 2.0 ** spec.mass_scale_log2 - ref_m).max() < 1e-6


def test_binning_fused_occupancy_dot(dp, args):
    """dp_bin2d_occ: time-sliced binning (shared-memory window + global fall-back) equals the slice-by-slice binning, and
    the fused obstacle-grid dot product equals sum_cells mass * grid computed from the grids afterwards."""
    from density_planner_b200 import _lib, grid as G
    e, denv, _, _ = _env(dp, args, 2)
    Ts, N = 7, 200_003
    g = torch.Generator(device="cuda").manual_seed(5)
    tt = torch.linspace(0, 1, Ts, device="cuda")
    # slice-dependent spread: tight clouds (all in the window), wide clouds (fall-back path) and samples beyond the grid
    sig = torch.tensor([0.3, 0.6, 1.5, 4.0, 9.0, 0.1, 20.0], device="cuda")[:, None]
    x = (0.5 * torch.sin(3 * tt))[:, None] + sig * torch.randn(Ts, N, device="cuda", generator=g)
    y = (-28 + 36 * tt)[:, None] + sig * torch.randn(Ts, N, device="cuda", generator=g)
    w = torch.rand(Ts, N, device="cuda", generator=g)
    w = (w / w.sum(dim=1, keepdim=True)).contiguous()
    mass, cnt, dot, sc = dp.bin_occupancy(denv, x.contiguous(), y.contiguous(), w, args)
    assert int(cnt.sum()) == Ts * N
    spec = _lib.BinSpec()
    spec.mode = 0
    spec.geom = G._geom(args)
    spec.mass_scale_log2 = sc
    dev = x.device
    for t in range(Ts):
        m1, c1 = G.bin2d_raw(spec, x[t].contiguous(), y[t].contiguous(), w[t].contiguous(), N, 1, (1, 0), (1, 0), dev)
        assert bool((m1[0] == mass[t]).all()) and bool((c1[0] == cnt[t]).all())
    grid = torch.from_numpy(e["grid"]).cuda().double()  # (241, 401, 101)
    ref = (mass[:, :241, :401].double() * grid[:, :, :Ts].permute(2, 0, 1)).sum(dim=(1, 2)) / 2.0 ** sc
    assert float(((dot - ref).abs() / ref.abs().clamp(min=1e-30)).max()) < 1e-10
    assert float(ref.max()) > 0


def _env(dp, args, seed):
    e = golden("env_seed%d.npz" % seed)
    import dp_oracle
    gx, gy = dp_oracle.env_gradient(e["grid"])
    denv = dp.DeviceEnv(torch.from_numpy(e["grid"]), torch.from_numpy(gx), torch.from_numpy(gy), args)
    return e, denv, gx, gy


@pytest.mark.parametrize("seed", list(range(10)))
def test_cost_coll_vs_reference_golden(dp, args, seed):
    """Collision cost on every golden environment (plan_motion's 10 seeds): synthetic 20k weighted samples."""
    from make_golden import synth_samples
    e, denv, _, _ = _env(dp, args, seed)
    Ts = 101
    tt = np.linspace(0, 1, Ts, dtype=np.float32)
    path = np.stack([0.5 * np.sin(3 * tt), -28 + 36 * tt], 0)[None].astype(np.float32)
    xy, rho = synth_samples(seed, 20000, Ts, path)
    x5 = torch.zeros(xy.shape[0], 5, Ts)
    x5[:, :2, :] = torch.from_numpy(xy)
    x5.requires_grad_(True)
    r = torch.from_numpy(rho).unsqueeze(1).clone().requires_grad_(True)
    c = dp.get_cost_coll(x5, r, denv)
    assert c.shape == (1,)
    assert abs(float(c) - e["syn_cost"][0]) <= COST_RTOL * abs(e["syn_cost"][0])
    c.sum().backward()
    assert (x5.grad[:64, :2, :].numpy() == e["syn_gx_head"]).all()
    assert (r.grad[:64, 0, :].numpy() == e["syn_grho_head"]).all()
    assert float(x5.grad[:, 2:, :].abs().max()) == 0.0
    assert np.abs(x5.grad[:, :2, :].double().sum(0).numpy() - e["syn_gx_sum"]).max() < 1e-9
    assert np.abs(r.grad.double().sum(0).numpy()[0] - e["syn_grho_sum"]).max() < 1e-6
    cm = dp.get_cost_coll(x5.detach(), r.detach(), denv, get_max=True)
    assert abs(float(cm) - e["syn_cost_max"][0]) <= COST_RTOL * abs(e["syn_cost_max"][0])


@pytest.mark.parametrize("seed", [0, 1])
def test_cost_coll_on_le_trajectories(dp, args, seed):
    """The planner's own call: LE trajectories of 500 samples, reference layouts (N,5,101)/(N,1,1001), CPU tensors."""
    e, denv, _, _ = _env(dp, args, seed)
    x5 = torch.zeros(e["le_x"].shape[0], 5, 101)
    x5[:, :2, :] = torch.from_numpy(e["le_x"])
    x5.requires_grad_(True)
    r = torch.from_numpy(e["le_rho"]).unsqueeze(1).clone().requires_grad_(True)
    c = dp.get_cost_coll(x5, r, denv)
    assert abs(float(c) - e["le_cost"][0]) <= COST_RTOL * abs(e["le_cost"][0])
    (2.0 * c.sum()).backward()
    assert (x5.grad[:, :2, :].numpy() == 2 * e["le_gx"]).all()
    assert (r.grad[:, 0, :101].numpy() == 2 * e["le_grho"]).all()
    assert float(r.grad[:, 0, 101:].abs().max()) == 0.0
    cm = dp.get_cost_coll(x5.detach(), r.detach(), denv, get_max=True)
    assert abs(float(cm) - e["le_cost_max"][0]) <= COST_RTOL * abs(e["le_cost_max"][0])
    ci = dp.get_cost_coll_initialize(x5.detach()[:100], denv)
    assert ci.shape == (100,)
    assert np.allclose(ci.numpy(), e["le_cost_init"], rtol=COST_RTOL)


def test_cost_layouts_agree_and_full_size(dp, oracle, args):
    """1M samples x 101 slices in the kernel's native [t][n] layout: equals the sum over shards (linearity in the
    sample set), equals the strided-layout path on a subset, and matches the oracle on that subset."""
    e, denv, gx, gy = _env(dp, args, 3)
    N, Ts = 1_000_000, 101
    g = torch.Generator(device="cuda").manual_seed(1)
    tt = torch.linspace(0, 1, Ts, device="cuda")
    x = (0.5 * torch.sin(3 * tt))[:, None] + 0.6 * torch.randn(Ts, N, device="cuda", generator=g)
    y = (-28 + 36 * tt)[:, None] + 0.6 * torch.randn(Ts, N, device="cuda", generator=g)
    rho = torch.rand(Ts, N, device="cuda", generator=g)
    rho = rho / rho.sum(dim=1, keepdim=True)
    from density_planner_b200 import grid as G
    tot, _, _, _, _ = G._cost_call(denv, x.t(), y.t(), rho.t(), False, False, False)
    parts = 0.0
    for lo, hi in ((0, 333_333), (333_333, 700_001), (700_001, N)):
        p, _, _, _, _ = G._cost_call(denv, x[:, lo:hi].t(), y[:, lo:hi].t(), rho[:, lo:hi].t(), False, False, False)
        parts += float(p)
    assert abs(parts - float(tot)) <= 1e-9 * abs(float(tot))
    n = 5000
    xs, ys, rs = x[:, :n].t().contiguous(), y[:, :n].t().contiguous(), rho[:, :n].t().contiguous()  # (n, Ts) time-fast
    a, _, gxa, gya, gra = G._cost_call(denv, xs, ys, rs, False, True, False)
    b, _, gxb, gyb, grb = G._cost_call(denv, x[:, :n].contiguous().t(), y[:, :n].contiguous().t(),
                                       rho[:, :n].contiguous().t(), False, True, False)
    assert abs(float(a) - float(b)) <= 1e-12 * abs(float(a))
    assert bool((gxa == gxb).all()) and bool((gya == gyb).all()) and bool((gra == grb).all())
    oc, odx, ody, odr = oracle.cost_coll(xs.cpu().numpy(), ys.cpu().numpy(), rs.cpu().numpy(), e["grid"], gx, gy,
                                         want_grad=True)
    assert abs(float(a) - oc) <= 1e-6 * abs(oc)
    assert (gxa.cpu().numpy() == odx).all() and (gya.cpu().numpy() == ody).all() and (gra.cpu().numpy() == odr).all()


def test_normalizer_and_planner_le_call(dp, car, oracle, blob, args):
    """EgoPredictor.predict_density(use_nn=False) against the reference golden (env seed 0) incl. the quirk that the
    density keeps 1001 time points while the states keep 101 (simulation_objects.py:293-298)."""
    e = golden("env_seed0.npz")
    xe0 = torch.from_numpy(e["le_xe0"]).unsqueeze(-1)
    rho0 = torch.from_numpy(e["le_rho0"]).reshape(-1, 1, 1)
    xref0 = torch.from_numpy(e["le_xref0"]).reshape(1, 5, 1)
    up = torch.from_numpy(e["le_up"])
    pred = dp.EgoPredictor(car, args, xref0, xe0, rho0)
    uref_traj, xref_traj = car.up2ref_traj(xref0, up, args, short=True)
    x_traj, rho_traj = pred.predict_density(up, xref_traj, use_nn=False)
    assert x_traj.shape == (500, 5, 101) and rho_traj.shape == (500, 1, 1001)
    assert np.abs(x_traj[:, :2, :].numpy() - e["le_x"]).max() < 1e-4
    ref = e["le_rho"]
    got = rho_traj[:, 0, :].numpy()
    big = ref > 1e-6
    # The normalised density is rho_s = exp(l_s) rho0_s / sum_j exp(l_j) rho0_j, so
    #     d(log rho_s) = d(l_s) - sum_j rho_j d(l_j):
    # an error of the log-density l (contract: rtol 1e-3 + 1e-4 on |l|, |l| reaches ~40 over these 1000 steps) moves log rho
    # by the sample's own error plus the DENSITY-WEIGHTED MEAN of everybody's error (the shift of the normaliser).  The bound
    # below is that first-order propagation of the contract, per sample and time index -- no slack factor.  The oracle
    # (pinned to the reference at 8e-7 rel) supplies l.
    uref_long, _ = car.sample_uref_traj(args, up=up)
    xref_long = car.compute_xref_traj(xref0, uref_long, args)
    o_rho = oracle.rollout(blob, e["le_xe0"], xref_long[0].numpy(), uref_long[0].numpy(), args.dt_sim, cutting=False)[1]
    tol_l = RHO_RTOL * np.abs(o_rho) + RHO_ATOL                      # (N, 1001): the contract on l
    tol = tol_l + (ref * tol_l).sum(axis=0, keepdims=True)            # + normaliser shift (ref sums to 1 per time index)
    err = np.abs(np.log(got.clip(1e-30)) - np.log(ref.clip(1e-30)))
    assert (err <= tol)[big].all(), "normalised density: max log error %.3g (tol %.3g)" % (err[big].max(), tol[big].max())
    assert np.abs(got.sum(axis=0) - 1).max() < 1e-4
    # and the cost the planner computes from it.  (a) on IDENTICAL inputs (the reference's own x_traj and rho_traj) the
    # contract's 1e-4 holds; (b) end to end the cost inherits the density's deviation: it is linear in rho, so the bound is
    # the tolerance-weighted sum of the per-sample terms -- asserted, and the measured deviation is far inside it (< 2e-3)
    gx, gy = oracle.env_gradient(e["grid"])
    cc = dp.CollisionCost(torch.from_numpy(e["grid"]), torch.from_numpy(gx), torch.from_numpy(gy), args)
    x_ref5 = torch.zeros(500, 5, 101)
    x_ref5[:, :2, :] = torch.from_numpy(e["le_x"])
    c_same = cc.get_cost_coll(x_ref5, torch.from_numpy(ref).unsqueeze(1))
    assert abs(float(c_same) - e["le_cost"][0]) <= COST_RTOL * abs(e["le_cost"][0])
    c = cc.get_cost_coll(x_traj, rho_traj)
    xs, ys = x_ref5[:, 0, :].numpy(), x_ref5[:, 1, :].numpy()
    _, _, _, odr = oracle.cost_coll(xs, ys, ref[:, :101].copy(), e["grid"], gx, gy, want_grad=True)   # d cost / d rho = p * sq
    bound = float((np.abs(odr) * ref[:, :101] * np.expm1(tol[:, :101])).sum()) + COST_RTOL * abs(e["le_cost"][0])
    dev = abs(float(c) - e["le_cost"][0])
    print("LE cost end to end: deviation %.3g rel, propagated contract bound %.3g rel" % (dev / abs(e["le_cost"][0]), bound / abs(e["le_cost"][0])))
    assert dev <= bound and dev <= 2e-3 * abs(e["le_cost"][0])
