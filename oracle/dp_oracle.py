"""CPU oracle for the density_planner hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this
module.  The product package never does, and has no CPU fallback.

Parity is pinned on tests/golden/*.npz, outputs of the unmodified reference minted by
tests/golden/make_golden.py (the reference has no tests or golden vectors of its own).

* rollout / single step: C restatement in dp_oracle.c (fp32 and fp64 builds), loaded with ctypes.
* grid functions: numpy restatements below, each citing the reference lines it follows.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

DPO_LOG, DPO_DENSITY, DPO_CUT = 1, 2, 4

# hyperparams.py:124,154-155
ENV_SIZE = (-12.0, 12.0, -30.0, 10.0)
GRID_SIZE = (241, 401)


def build(force=False):
    so = os.path.join(_HERE, "libdp_oracle.so")
    src = os.path.join(_HERE, "dp_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libdp_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libdp_oracle.so")
        if not os.path.exists(so):
            build()
        _LIB = ctypes.CDLL(so)
    return _LIB


def set_threads(n):
    """Set the OpenMP thread count of the C oracle; returns the count in effect."""
    return int(lib().dpo_set_threads_f32(int(n)))


def pack_weights(sd):
    """state_dict (numpy or torch values, keys of U_FUNC: model_u_w{1,2}.{0,2,4}.{weight,bias}) -> fp32 blob.

    Layout documented in dp_oracle.c / include/density_planner_b200.h."""
    parts = []
    for net in ("model_u_w1", "model_u_w2"):
        for layer in ("0", "2", "4"):
            for kind in ("weight", "bias"):
                v = sd["%s.%s.%s" % (net, layer, kind)]
                v = v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
                parts.append(np.ascontiguousarray(v, dtype=np.float32).ravel())
    blob = np.concatenate(parts)
    assert blob.size == 43592, blob.size
    return blob


def load_weights(path=None):
    if path is None:
        path = os.path.join(os.path.dirname(_HERE), "density_planner_b200", "data", "controller_CAR_3layers.npz")
    z = np.load(path)
    return pack_weights({k: z[k] for k in z.files})


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def rollout(blob, xe0, xref, uref, dt, rho0=None, log_density=True, compute_density=True, cutting=True, stride=1,
            dtype=np.float32, want_margin=False):
    """compute_density (systems/ControlAffineSystem.py:409-463).

    xe0 (N,5), xref (5,T+1), uref (2,T) -> xe_traj (N,5,Ts), rho (N,Ts) or None, [clip_margin (N,)], cut_flags."""
    L = lib()
    dtype = np.dtype(dtype)
    fn = L.dpo_rollout_f32 if dtype == np.float32 else L.dpo_rollout_f64
    blob = np.ascontiguousarray(blob, np.float32)
    xe0 = np.ascontiguousarray(xe0, dtype)
    xref = np.ascontiguousarray(xref, dtype)
    uref = np.ascontiguousarray(uref, dtype)
    N, T = xe0.shape[0], uref.shape[1]
    assert xref.shape == (5, T + 1) and xe0.shape == (N, 5)
    Ts = T // stride + 1
    xe_out = np.empty((N, 5, Ts), dtype)
    rho_out = np.empty((N, Ts), dtype) if compute_density else None
    margin = np.empty(N, dtype) if want_margin else None
    r0 = None if rho0 is None else np.ascontiguousarray(rho0, dtype).reshape(N)
    flags = (DPO_LOG if log_density else 0) | (DPO_DENSITY if compute_density else 0) | (DPO_CUT if cutting else 0)
    cut = (ctypes.c_int * 2)()
    dt_c = ctypes.c_float(dt) if dtype == np.float32 else ctypes.c_double(float(np.float32(dt)))
    fn(_ptr(blob), _ptr(xe0), _ptr(xref), _ptr(uref), _ptr(r0), dt_c, ctypes.c_long(N), ctypes.c_int(T),
       ctypes.c_uint(flags), ctypes.c_int(stride), _ptr(xe_out), _ptr(rho_out), _ptr(margin), cut)
    res = [xe_out, rho_out]
    if want_margin:
        res.append(margin)
    res.append((int(cut[0]), int(cut[1])))
    return tuple(res)


def step(blob, x, xref5, uref2, dtype=np.float32, jac=True):
    """Single-step API: u_raw, u (clipped), dudx (N,2,5) incl. clip mask, f (N,5), div (N,).

    u_func :26 / f_func :43 / dudx_func :69 / divf_func :116 of systems/ControlAffineSystem.py."""
    L = lib()
    dtype = np.dtype(dtype)
    fn = L.dpo_step_f32 if dtype == np.float32 else L.dpo_step_f64
    x = np.ascontiguousarray(x, dtype)
    N = x.shape[0]
    xr = np.ascontiguousarray(xref5, dtype).reshape(5)
    ur = np.ascontiguousarray(uref2, dtype).reshape(2)
    u_raw = np.empty((N, 2), dtype)
    u = np.empty((N, 2), dtype)
    dudx = np.empty((N, 2, 5), dtype) if jac else None
    f = np.empty((N, 5), dtype)
    div = np.empty(N, dtype)
    fn(_ptr(np.ascontiguousarray(blob, np.float32)), _ptr(x), _ptr(xr), _ptr(ur), ctypes.c_long(N), _ptr(u_raw), _ptr(u),
       _ptr(dudx), _ptr(f), _ptr(div))
    return dict(u_raw=u_raw, u=u, dudx=dudx, f=f, div=div)


# ----------------------------------------------------------------------------------------------------------
# grid functions (numpy restatements)
# ----------------------------------------------------------------------------------------------------------
_F = np.float32


def pos2gridpos(pos_x=None, pos_y=None, env=ENV_SIZE, grid=GRID_SIZE):
    """motion_planning/utils.py:70-97.  fp32 op for op: ((p - e0) / (e1 - e0)) * (G-1), + 0.001, round-half-even."""
    out = []
    for p, lo, hi, g in ((pos_x, env[0], env[1], grid[0]), (pos_y, env[2], env[3], grid[1])):
        if p is None:
            out.append(None)
            continue
        p = np.asarray(p, _F)
        q = ((p - _F(lo)) / _F(hi - lo)) * _F(g - 1)
        out.append(np.rint(q + _F(0.001)).astype(np.int64))
    return out[0], out[1]


def gridpos2pos(pos_x=None, pos_y=None, env=ENV_SIZE, grid=GRID_SIZE):
    """motion_planning/utils.py:100-116 on float grid positions."""
    out = []
    for p, lo, hi, g in ((pos_x, env[0], env[1], grid[0]), (pos_y, env[2], env[3], grid[1])):
        if p is None:
            out.append(None)
            continue
        p = np.asarray(p, _F)
        out.append((p / _F(g - 1)) * _F(hi - lo) + _F(lo))
    return out[0], out[1]


def _shift(grid, step_x=0, step_y=0, fill=0):
    """shift_array (motion_planning/utils.py:136-164) on the first two axes."""
    res = np.full_like(grid, fill)
    G0, G1 = grid.shape[0], grid.shape[1]
    sx0, sx1 = max(0, -step_x), min(G0, G0 - step_x)
    sy0, sy1 = max(0, -step_y), min(G1, G1 - step_y)
    if sx1 > sx0 and sy1 > sy0:
        res[sx0 + step_x:sx1 + step_x, sy0 + step_y:sy1 + step_y] = grid[sx0:sx1, sy0:sy1]
    return res


def compute_gradient(grid, step=1):
    """motion_planning/utils.py:188-199."""
    gx = (_shift(grid, step_x=step, fill=1) - _shift(grid, step_x=-step, fill=1)) / _F(2 * step)
    gy = (_shift(grid, step_y=step, fill=1) - _shift(grid, step_y=-step, fill=1)) / _F(2 * step)
    return gx.astype(_F), gy.astype(_F)


def env_gradient(grid):
    """Environment.get_gradient (motion_planning/simulation_objects.py:77-94): centred differences, multi-scale fill."""
    grid = np.asarray(grid, _F)
    gx, gy = compute_gradient(grid, 1)
    s = 5
    missing = (grid != 0) & (gx == 0) & (gy == 0)
    while missing.any():
        gxn, gyn = compute_gradient(grid, s)
        gx[missing] += _F(s) * gxn[missing]
        gy[missing] += _F(s) * gyn[missing]
        s += 10
        missing = (grid != 0) & (gx == 0) & (gy == 0)
    return gx, gy


def cost_coll(x, y, rho, grid, gradx, grady, get_max=False, want_grad=False, per_sample=False,
              env=ENV_SIZE, gsize=GRID_SIZE):
    """MotionPlanner.get_cost_coll (motion_planning/MotionPlanner.py:192-224) and, with per_sample=True and rho=None,
    MotionPlannerGrad.get_cost_coll_initialize (:309-336).

    x, y: (N,Ts) fp32 world positions; rho: (N, >=Ts); grids (G0,G1,>=Ts) fp32 time-innermost.
    Returns cost (float64 accumulation of the fp32 per-sample terms) [, dcost/dx, dcost/dy, dcost/drho]."""
    x = np.asarray(x, _F)
    y = np.asarray(y, _F)
    N, Ts = x.shape
    gx, gy = pos2gridpos(x, y, env, gsize)
    gx = np.clip(gx, 0, gsize[0] - 1)
    gy = np.clip(gy, 0, gsize[1] - 1)
    ti = np.broadcast_to(np.arange(Ts), (N, Ts))
    GX = gradx[gx, gy, ti]
    GY = grady[gx, gy, ti]
    active = (GX != 0) | (GY != 0)
    des_gx = gx.astype(_F) + _F(100) * GX
    des_gy = gy.astype(_F) + _F(100) * GY
    dpx, dpy = gridpos2pos(des_gx, des_gy, env, gsize)
    ex = dpx - x
    ey = dpy - y
    sq = ex * ex + ey * ey
    p = grid[gx, gy, ti]
    if rho is None:
        term = p * sq
    else:
        r = np.asarray(rho, _F)[:, :Ts]
        term = r * p * sq
    term = np.where(active, term, _F(0)).astype(_F)
    if per_sample:
        cost = term.astype(np.float64).sum(axis=1)
    elif get_max:
        # cost += (...)[idx].max() per slice that has any active sample (:221)
        cost = 0.0
        for i in range(Ts):
            if active[:, i].any():
                cost += float(term[active[:, i], i].max())
    else:
        cost = float(term.astype(np.float64).sum())
    if not want_grad:
        return cost
    # backward of the sum: indices carry no grad (:206); d/dx of (des - x)^2 = -2 (des - x)
    w = p if rho is None else r * p
    dx = np.where(active, _F(-2) * w * ex, _F(0)).astype(_F)
    dy = np.where(active, _F(-2) * w * ey, _F(0)).astype(_F)
    drho = None if rho is None else np.where(active, p * sq, _F(0)).astype(_F)
    return cost, dx, dy, drho


def pred2grid(x, y, rho, env=ENV_SIZE, gsize=GRID_SIZE):
    """motion_planning/utils.py:255-280 for one time slice: mean rho per occupied cell, normalised to sum 1.

    Returns (grid (G0,G1) fp32, gridpos (N,2) int64, counts (G0,G1) int64).  Indices clamp to [0, G] (sic, :261-262);
    the reference fails on the slice assignment when a sample reaches G, so callers keep samples below it."""
    gx, gy = pos2gridpos(x, y, env, gsize)
    gx = np.clip(gx, 0, gsize[0])
    gy = np.clip(gy, 0, gsize[1])
    cnt = np.zeros((gsize[0] + 1, gsize[1] + 1), np.int64)
    np.add.at(cnt, (gx, gy), 1)
    mass = np.zeros((gsize[0] + 1, gsize[1] + 1), np.float64)
    np.add.at(mass, (gx, gy), np.asarray(rho, np.float64))
    mean = np.zeros_like(mass)
    m = cnt > 0
    mean[m] = mass[m] / cnt[m]
    grid = (mean / mean.sum()).astype(_F)
    return grid[:gsize[0], :gsize[1]], np.stack([gx, gy], 1), cnt[:gsize[0], :gsize[1]]


def density_map(xe, rho, bins=(20, 20), xe_range=(4.0, 4.0), log_density=False, kind="LE"):
    """get_density_map (systems/utils.py:105-129): 2-D histogram of error states, mean density per bin, normalised.

    bin_width = (XE_MAX - XE_MIN)[:2] / bins = 4/20; range = +-bins*bin_width/2 = +-2."""
    xe = np.asarray(xe, _F)
    rho = np.asarray(rho, _F).copy()
    if kind != "MC" and log_density:
        rho = np.exp(rho - rho.max())
    else:
        rho = rho / rho.max()
    rng = [[-bins[i] * (xe_range[i] / bins[i]) / 2, bins[i] * (xe_range[i] / bins[i]) / 2] for i in range(2)]
    mean, _ = np.histogramdd(xe, bins=list(bins), weights=rho, range=rng)
    mask = mean > 0
    if kind != "MC":
        num, _ = np.histogramdd(xe, bins=list(bins), range=rng)
        mean[mask] /= num[mask]
    mean[mask] = mean[mask] / mean[mask].sum()
    return mean, rng


def normalize_density(rholog, rho0):
    """EgoVehicle.predict_density normaliser (motion_planning/simulation_objects.py:295-299).

    rholog (N,T) log densities, rho0 (N,) -> rho (N,T), each time column summing to 1."""
    rholog = np.asarray(rholog, _F)
    m = rholog.max(axis=0, keepdims=True)
    un = np.exp(rholog - m) * np.asarray(rho0, _F).reshape(-1, 1)
    return (un / un.sum(axis=0, keepdims=True, dtype=_F)).astype(_F)


# ---- reference-trajectory integrator and goal / bounds cost terms ---------------------
def xref_traj(xref0, uref, dt):
    """compute_xref_traj (systems/ControlAffineSystem.py:341-354; CAR dynamics sytem_CAR.py:128-176), fp32 op for op.
    xref0 (B,5) or (5,), uref (B,2,T) -> xref (B,5,T+1)."""
    uref = np.asarray(uref, np.float32)
    B, _, T = uref.shape
    x = np.broadcast_to(np.asarray(xref0, np.float32).reshape(-1, 5), (B, 5)).copy()
    dt = np.float32(dt)
    out = np.empty((B, 5, T + 1), np.float32)
    out[:, :, 0] = x
    for k in range(T):
        cs, sn = np.cos(x[:, 2]).astype(np.float32), np.sin(x[:, 2]).astype(np.float32)
        x = x.copy()
        x[:, 0] = x[:, 0] + (x[:, 3] * cs) * dt
        x[:, 1] = x[:, 1] + (x[:, 3] * sn) * dt
        x[:, 2] = x[:, 2] + uref[:, 0, k] * dt
        x[:, 3] = x[:, 3] + uref[:, 1, k] * dt
        out[:, :, k + 1] = x
    return out


def xref_traj_grad(xref, gxref, dt):
    """Discrete adjoint of xref_traj in float64: upstream gxref (B,5,T+1) -> (guref (B,2,T), gxref0 (B,5))."""
    xref = np.asarray(xref, np.float64)
    g = np.asarray(gxref, np.float64)
    B, _, T1 = xref.shape
    T = T1 - 1
    lam = np.zeros((B, 5))
    gu = np.zeros((B, 2, T))
    for k in range(T, 0, -1):
        lam = lam + g[:, :, k]
        gu[:, 0, k - 1] = lam[:, 2] * dt
        gu[:, 1, k - 1] = lam[:, 3] * dt
        th, v = xref[:, 2, k - 1], xref[:, 3, k - 1]
        cs, sn = np.cos(th), np.sin(th)
        lth = lam[:, 2] + dt * v * (cs * lam[:, 1] - sn * lam[:, 0])
        lv = lam[:, 3] + dt * (cs * lam[:, 0] + sn * lam[:, 1])
        lam = lam.copy()
        lam[:, 2], lam[:, 3] = lth, lv
    return gu, lam + g[:, :, 0]


def cost_goal_bounds(x, rho, lo, hi, goal_xy, get_max=False, want_grad=False):
    """Raw terms of get_cost_goal / get_cost_bounds (motion_planning/MotionPlanner.py:121-190): x (N,5,Ts), rho (N,Ts).
    Returns dict(goal, below, above, any_lo4, any_hi4[, gx, grho for unit upstream factors on all three terms])."""
    x = np.asarray(x, np.float32)
    rho = np.asarray(rho, np.float32)
    lo = np.asarray(lo, np.float32).reshape(1, 5, 1)
    hi = np.asarray(hi, np.float32).reshape(1, 5, 1)
    red = (lambda a: float(a.max()) if a.size else 0.0) if get_max else (lambda a: float(a.astype(np.float64).sum()))
    mlo, mhi = x < lo, x > hi
    elo = (x - lo).astype(np.float32)
    ehi = (x - hi).astype(np.float32)
    r3 = rho[:, None, :]
    tlo = (r3 * (elo * elo))[mlo]
    thi = (r3 * (ehi * ehi))[mhi]
    e = x[:, :2, -1] - np.asarray(goal_xy, np.float32)[None]
    sq = (e[:, 0] * e[:, 0] + e[:, 1] * e[:, 1]).astype(np.float32)
    out = dict(goal=red(rho[:, -1] * sq), below=red(tlo), above=red(thi), any_lo4=bool(mlo[:, :4].any()),
               any_hi4=bool(mhi[:, :4].any()))
    if want_grad:
        gx = np.zeros(x.shape, np.float64)
        gx += np.where(mlo, 2.0 * r3 * np.where(mlo, elo, 0), 0.0) + np.where(mhi, 2.0 * r3 * np.where(mhi, ehi, 0), 0.0)
        gx[:, :2, -1] += 2.0 * rho[:, -1:] * e
        gr = (np.where(mlo, elo.astype(np.float64) ** 2, 0.0) + np.where(mhi, ehi.astype(np.float64) ** 2, 0.0)).sum(axis=1)
        gr[:, -1] += sq
        out.update(gx=gx, grho=gr)
    return out
