"""CPU suite: the oracle restatement (oracle/dp_oracle.{c,py}) against the golden vectors minted from the unmodified
reference (tests/golden/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

from conftest import golden, assert_states_close, assert_rho_close, COST_RTOL


@pytest.mark.parametrize("name", ["rollout_cfg1", "rollout_n2000_t300", "rollout_n4000_t100"])
def test_rollout_matches_reference(oracle, blob, name):
    g = golden(name + ".npz")
    stride = int(g["stride"]) if "stride" in g.files else 1
    xe, rho, margin, cut = oracle.rollout(blob, g["xe0"], g["xref"], g["uref"], float(g["dt"]), stride=stride,
                                          want_margin=True)
    xref_s = g["xref"][:, ::stride]
    assert_states_close(xe, g["xe_traj"], xref_s, name)
    assert_rho_close(rho, g["rholog_traj"], g["clip_margin"], name)
    # the oracle's own clip margin agrees with the one computed from the reference trajectory
    assert np.abs(margin - g["clip_margin"]).max() < 1e-3
    assert cut == (0, 0)


def test_rollout_modes(oracle, blob):
    g = golden("rollout_modes.npz")
    s = int(g["stride"])
    dt = float(g["dt"])
    xref_s = g["xref"][:, ::s]
    xe, rho, cut = oracle.rollout(blob, g["xe0"], g["xref"], g["uref"], dt, rho0=g["rho0"], log_density=False, stride=s)
    assert_states_close(xe, g["xe_lin"], xref_s, "linear")
    assert np.allclose(rho, g["rho_lin"], rtol=1e-3, atol=1e-6)
    xe, rho, cut = oracle.rollout(blob, g["xe0"], g["xref"], g["uref"], dt, log_density=True, cutting=False, stride=s)
    assert_rho_close(rho, g["rholog"], None, "log nocut")
    xe, rho, cut = oracle.rollout(blob, g["xe0"], g["xref"], g["uref"], dt, compute_density=False, stride=s)
    assert rho is None
    assert_states_close(xe, g["xe_nodens"], xref_s, "no density")


def test_step_api(oracle, blob):
    g = golden("step_api.npz")
    r = oracle.step(blob, g["x"], g["xref"], g["uref"])
    assert np.allclose(r["u_raw"], g["u_raw"], rtol=1e-5, atol=2e-5)
    assert np.allclose(r["u"], g["u"], rtol=1e-5, atol=2e-5)
    assert np.allclose(r["f"], g["f"], rtol=1e-5, atol=2e-5)
    # Jacobian rows vanish where the input is clipped; compare where reference and oracle agree on the mask
    safe = (np.abs(np.abs(g["u_raw"]) - 3) > 1e-4).all(axis=1)
    assert safe.mean() > 0.99
    scale = np.abs(g["dudx"]).max()
    assert np.abs(r["dudx"][safe] - g["dudx"][safe]).max() < 1e-4 * scale
    assert np.abs(r["div"][safe] - g["div"][safe]).max() < 1e-4 * np.abs(g["div"]).max()
    dfdx = g["dfdx"]
    assert np.abs(np.trace(dfdx, axis1=1, axis2=2)[safe] - r["div"][safe]).max() < 1e-4 * np.abs(g["div"]).max()
    dt = float(g["dt"])
    xn = g["x"] + r["f"] * np.float32(dt)
    assert np.allclose(xn, g["x_next"], rtol=1e-6, atol=1e-6)
    rn = g["rho"] + (-r["div"] * g["rho"]) * np.float32(dt)
    assert np.allclose(rn[safe], g["rho_next"][safe], rtol=1e-4, atol=1e-5)
    rln = g["rho"] + np.log(1 - r["div"] * np.float32(dt))
    assert np.allclose(rln[safe], g["rholog_next"][safe], rtol=1e-4, atol=1e-5)


def test_fp64_build_agrees(oracle, blob):
    g = golden("rollout_n4000_t100.npz")
    n = 256
    xe32, rho32, m32, _ = oracle.rollout(blob, g["xe0"][:n], g["xref"], g["uref"], float(g["dt"]), want_margin=True)
    xe64, rho64, m64, _ = oracle.rollout(blob, g["xe0"][:n], g["xref"], g["uref"], float(g["dt"]), dtype=np.float64,
                                         want_margin=True)
    safe = np.minimum(m32, m64) > 1e-3
    assert safe.mean() > 0.8
    assert np.abs(xe32[safe] - xe64[safe]).max() < 1e-4
    assert np.abs(rho32[safe] - rho64[safe]).max() < 1e-2


def test_pos2gridpos_bit_exact(oracle):
    from make_golden import grid_inputs
    g = golden("pos2gridpos.npz")
    px, py = grid_inputs(int(g["n"]), int(g["seed"]))
    gx, gy = oracle.pos2gridpos(px, py)
    assert (gx == g["gx"]).all() and (gy == g["gy"]).all()
    ghx, _ = oracle.pos2gridpos(g["hx"], None)
    _, ghy = oracle.pos2gridpos(None, g["hy"])
    assert (ghx == g["ghx"]).all() and (ghy == g["ghy"]).all()
    bx, by = oracle.gridpos2pos(g["fgx"], g["fgy"])
    assert (bx == g["bx"]).all() and (by == g["by"]).all()


def test_pred2grid_and_density_map(oracle):
    g = golden("pred2grid.npz")
    grid, gp, cnt = oracle.pred2grid(g["x"][:, 0], g["x"][:, 1], g["rho"])
    ref = np.zeros((241, 401), np.float32)
    ref[g["nz"][:, 0], g["nz"][:, 1]] = g["val"]
    assert (gp == g["gridpos"]).all()
    assert ((grid > 0) == (ref > 0)).all()
    assert np.allclose(grid, ref, rtol=1e-5, atol=1e-10)
    assert cnt.sum() == len(g["rho"])
    d = golden("density_map.npz")
    dm, rb = oracle.density_map(d["xe"], d["rholog"], log_density=True)
    assert np.allclose(dm, d["dm"], rtol=1e-6, atol=1e-9)
    assert np.allclose(np.array(rb), d["range_bins"])
    dm, _ = oracle.density_map(d["xe"], np.exp(d["rholog"]), log_density=False)
    assert np.allclose(dm, d["dm_lin"], rtol=1e-6, atol=1e-9)
    dm, _ = oracle.density_map(d["xe"], np.ones(len(d["xe"])), kind="MC")
    assert np.allclose(dm, d["dm_mc"], rtol=1e-12)


@pytest.mark.parametrize("seed", [0, 1, 4, 9])
def test_env_gradient_and_cost(oracle, seed):
    from make_golden import synth_samples
    e = golden("env_seed%d.npz" % seed)
    grid = e["grid"]
    gx, gy = oracle.env_gradient(grid)
    ci = e["chk_idx"].astype(int)
    assert (gx[ci[:, 0], ci[:, 1], ci[:, 2]] == e["chk_gx"]).all()
    assert (gy[ci[:, 0], ci[:, 1], ci[:, 2]] == e["chk_gy"]).all()
    assert abs(gx.astype(np.float64).sum() - float(e["gx_sum"])) < 1e-6
    assert abs(gy.astype(np.float64).sum() - float(e["gy_sum"])) < 1e-6
    Ts = 101
    tt = np.linspace(0, 1, Ts, dtype=np.float32)
    path = np.stack([0.5 * np.sin(3 * tt), -28 + 36 * tt], 0)[None].astype(np.float32)
    xy, rho = synth_samples(seed, 20000, Ts, path)
    c, dx, dy, dr = oracle.cost_coll(xy[:, 0], xy[:, 1], rho, grid, gx, gy, want_grad=True)
    assert abs(c - e["syn_cost"][0]) <= COST_RTOL * abs(e["syn_cost"][0])
    assert abs(oracle.cost_coll(xy[:, 0], xy[:, 1], rho, grid, gx, gy, get_max=True) - e["syn_cost_max"][0]) \
        <= COST_RTOL * abs(e["syn_cost_max"][0])
    assert (dx[:64] == e["syn_gx_head"][:, 0]).all() and (dy[:64] == e["syn_gx_head"][:, 1]).all()
    assert (dr[:64] == e["syn_grho_head"]).all()
    assert np.abs(dx.astype(np.float64).sum(0) - e["syn_gx_sum"][0]).max() < 1e-9
    if "le_x" in e.files:
        x, r = e["le_x"], e["le_rho"]
        c, dx, dy, dr = oracle.cost_coll(x[:, 0], x[:, 1], r, grid, gx, gy, want_grad=True)
        assert abs(c - e["le_cost"][0]) <= COST_RTOL * abs(e["le_cost"][0])
        assert (dx == e["le_gx"][:, 0]).all() and (dy == e["le_gx"][:, 1]).all() and (dr == e["le_grho"]).all()
        ci_ = oracle.cost_coll(x[:100, 0], x[:100, 1], None, grid, gx, gy, per_sample=True)
        assert np.allclose(ci_, e["le_cost_init"], rtol=COST_RTOL)


def test_normalizer_matches_reference_planner_call(oracle, blob):
    """EgoVehicle.predict_density LE branch (simulation_objects.py:288-299) on the golden of env seed 0."""
    e = golden("env_seed0.npz")
    import hashlib  # noqa: F401  (keep the import list of the CPU suite light)
    # rebuild the reference trajectory from up with the package's host logic, then the oracle rollout + normaliser
    import torch
    import density_planner_b200 as dp
    a = dp.default_args()
    car = dp.Car(a)
    up = torch.from_numpy(e["le_up"])
    uref, _ = car.sample_uref_traj(a, up=up)
    xref0 = torch.from_numpy(e["le_xref0"]).reshape(1, 5, 1)
    xref = car.compute_xref_traj(xref0, uref, a)
    xe, rl, _ = oracle.rollout(blob, e["le_xe0"], xref[0].numpy(), uref[0].numpy(), a.dt_sim, cutting=False)
    rho = oracle.normalize_density(rl, e["le_rho0"])
    x = xe[:, :2, ::10] + xref[0, :2, ::10].numpy()[None]
    assert np.abs(x - e["le_x"]).max() < 1e-4
    # density weights: relative agreement wherever the weight matters
    ref = e["le_rho"]
    big = ref > 1e-6
    assert np.abs(rho[big] / ref[big] - 1).max() < 5e-3
    assert np.abs(rho.sum(axis=0) - 1).max() < 1e-4


def test_oracle_reference_integrator_vs_golden(oracle):
    """compute_xref_traj / up2ref_traj and its gradient w.r.t. the input parameters."""
    g = golden("planner_terms.npz")
    xref = oracle.xref_traj(g["int_xref0"], g["int_uref"], 0.01)
    assert xref.shape == g["int_xref"].shape
    assert np.abs(xref - g["int_xref"]).max() < 2e-4  # torch's and numpy's cos/sin differ in the last bit over 1000 steps
    assert np.abs(xref[:, :, ::10] - g["int_xref_short"]).max() < 2e-4
    # loss = sum(w * xref) + 0.1 * sum(uref^2); uref = repeat_interleave(up, 100)[:, :, :999]
    gu, _ = oracle.xref_traj_grad(g["int_xref"], g["int_w"], 0.01)
    gu = gu + 0.2 * g["int_uref"]
    gup = np.zeros((6, 2, 10))
    for k in range(gu.shape[2]):
        gup[:, :, k // 100] += gu[:, :, k]
    assert np.abs(gup - g["int_gup"]).max() < 2e-4 * np.abs(g["int_gup"]).max()


def test_oracle_goal_bounds_costs_vs_golden(oracle):
    """get_cost_goal / get_cost_bounds forward (sum and max) and gradients."""
    g = golden("planner_terms.npz")
    for ev in (0, 1):
        lo, hi = (g["x_min"], g["x_max"]) if ev else (g["x_min_mp"], g["x_max_mp"])
        for mx in (0, 1):
            tag = "e%d_m%d" % (ev, mx)
            o = oracle.cost_goal_bounds(g["cost_x"], g["cost_rho"], lo, hi, g["xrefN"][:2], get_max=bool(mx), want_grad=not mx)
            goal = o["goal"]
            close = goal < float(g["close2goal_thr"])
            assert close == bool(g["close_" + tag])
            if not ev and not close:
                goal *= float(g["weight_goal_far"])
            assert abs(goal - float(g["goal_" + tag])) <= 2e-5 * abs(float(g["goal_" + tag]))
            bounds = o["below"] * o["any_lo4"] + o["above"] * o["any_hi4"]
            assert abs(bounds - float(g["bounds_" + tag][0])) <= 2e-5 * abs(float(g["bounds_" + tag][0]))
            assert (not (o["any_lo4"] or o["any_hi4"])) == bool(g["inb_" + tag])
            if not mx:
                s_goal = 1.0 if (ev or close) else float(g["weight_goal_far"])
                gx = o["gx"].copy()
                gx[:, :2, -1] *= 1.0  # goal part already has unit factor; rescale below
                e = g["cost_x"][:, :2, -1] - g["xrefN"][None, :2]
                gx[:, :2, -1] += (s_goal - 1.0) * 2.0 * g["cost_rho"][:, -1:] * e
                gr = o["grho"].copy()
                gr[:, -1] += (s_goal - 1.0) * (e ** 2).sum(axis=1)
                assert np.abs(gx - g["gx_" + tag]).max() <= 1e-5 * np.abs(g["gx_" + tag]).max()
                assert np.abs(gr - g["grho_" + tag]).max() <= 1e-5 * np.abs(g["grho_" + tag]).max()
