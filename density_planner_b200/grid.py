"""Grid layer: world<->grid transforms, occupancy binning and the collision cost, backed by the CUDA library.

Reference surface mirrored here:
  motion_planning/utils.py     pos2gridpos:70 gridpos2pos:100 pred2grid:255 check_collision:231
  systems/utils.py             get_density_map:105
  motion_planning/MotionPlanner.py       get_cost_coll:192        (forward + analytic backward)
  motion_planning/MotionPlannerGrad.py   get_cost_coll_initialize:309
  motion_planning/simulation_objects.py  EgoVehicle.predict_density normaliser :295-299

`args` is the reference's hyper-parameter namespace (hyperparams.py): only `environment_size` and `grid_size`
(and `bin_number` for get_density_map) are read.
"""
import ctypes
import sys
import math

import numpy as np
import torch

from . import _lib


def _geom(args):
    es, gs = args.environment_size, args.grid_size
    return _lib.GridGeom(float(es[0]), float(es[1]), float(es[2]), float(es[3]), int(gs[0]), int(gs[1]))


def _cuda_device(t=None):
    if not torch.cuda.is_available():
        raise _lib.DensityPlannerError("density_planner_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    if t is not None and t.is_cuda:
        return t.device
    return torch.device("cuda", torch.cuda.current_device())


def _pos2gridpos_one(args, p, which):
    """One coordinate of pos2gridpos with the reference's dtype rules (motion_planning/utils.py:79-97):
    float32 tensors -> fp32 ops (the hot case); python lists are `torch.from_numpy(np.array(list))` there, i.e. float64
    (or int64) tensors; numpy arrays and python scalars stay numpy (float64 arithmetic) and return numpy ints."""
    L = _lib.lib()
    as_numpy = False
    if isinstance(p, list):
        p = torch.from_numpy(np.array(p))
    elif not torch.is_tensor(p):
        as_numpy = True
        scalar = np.ndim(p) == 0
        a = np.asarray(p)
        # numpy: float32 arrays stay float32 next to python ints; everything else is evaluated in float64
        p = torch.from_numpy(np.ascontiguousarray(a if a.dtype == np.float32 else a.astype(np.float64)).reshape(-1))
        shape = a.shape
    dev = _cuda_device(p)
    g = _geom(args)
    lo = float(args.environment_size[0 if which == 0 else 2])
    if p.dtype == torch.float64:
        pd = p.detach().to(device=dev).contiguous()
        idx = torch.empty(pd.shape, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            a0, a1 = (_lib.ptr(pd), None) if which == 0 else (None, _lib.ptr(pd))
            i0, i1 = (_lib.ptr(idx), None) if which == 0 else (None, _lib.ptr(idx))
            _lib.check(L.dp_pos2gridpos_f64(ctypes.byref(g), a0, a1, pd.numel(), i0, i1, _lib.current_stream(dev)),
                       "dp_pos2gridpos_f64")
    else:
        if not p.dtype.is_floating_point:
            # integer tensors: (p - lo) stays integer, the division promotes it to float32 (torch true division)
            pd = (p.detach().to(dev) - int(lo)).to(torch.float32).contiguous()
            if which == 0:
                g = _lib.GridGeom(0.0, g.x_max - g.x_min, g.y_min, g.y_max, g.gx, g.gy)
            else:
                g = _lib.GridGeom(g.x_min, g.x_max, 0.0, g.y_max - g.y_min, g.gx, g.gy)
        else:
            pd = p.detach().to(device=dev, dtype=torch.float32).contiguous()
        idx = torch.empty(pd.shape, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            a0, a1 = (_lib.ptr(pd), None) if which == 0 else (None, _lib.ptr(pd))
            i0, i1 = (_lib.ptr(idx), None) if which == 0 else (None, _lib.ptr(idx))
            _lib.check(L.dp_pos2gridpos(ctypes.byref(g), a0, a1, pd.numel(), i0, i1, -1, -1, _lib.current_stream(dev)),
                       "dp_pos2gridpos")
    if as_numpy:
        out = idx.cpu().numpy().astype(int)
        return out.reshape(())[()] if scalar else out.reshape(shape)
    return idx.to(device=p.device, dtype=torch.int64)


def pos2gridpos(args, pos_x=None, pos_y=None):
    """World -> grid coordinates (motion_planning/utils.py:70-97): int64 tensors on the input's device for tensors and
    lists, numpy integers for numpy / scalar inputs, evaluated in the input's precision like the reference."""
    return (None if pos_x is None else _pos2gridpos_one(args, pos_x, 0),
            None if pos_y is None else _pos2gridpos_one(args, pos_y, 1))


def gridpos2pos(args, pos_x=None, pos_y=None):
    """Grid -> world coordinates (motion_planning/utils.py:100-116)."""
    L = _lib.lib()
    outs = []
    for p, which in ((pos_x, 0), (pos_y, 1)):
        if p is None:
            outs.append(None)
            continue
        p = torch.as_tensor(p)
        dev = _cuda_device(p)
        pd = p.detach().to(device=dev, dtype=torch.float32).contiguous()
        o = torch.empty_like(pd)
        g = _geom(args)
        with torch.cuda.device(dev):
            if which == 0:
                rc = L.dp_gridpos2pos(ctypes.byref(g), _lib.ptr(pd), None, pd.numel(), _lib.ptr(o), None,
                                      _lib.current_stream(dev))
            else:
                rc = L.dp_gridpos2pos(ctypes.byref(g), None, _lib.ptr(pd), pd.numel(), None, _lib.ptr(o),
                                      _lib.current_stream(dev))
        _lib.check(rc, "dp_gridpos2pos")
        outs.append(o.to(p.device))
    return outs[0], outs[1]


def mass_scale_log2(n_total, max_weight=1.0):
    """Fixed-point scale for the int64 mass accumulator: n_total weights of at most max_weight cannot overflow."""
    bits = 62 - math.ceil(math.log2(max(n_total, 2))) - max(0, math.ceil(math.log2(max(max_weight, 1e-30))))
    return int(max(0, min(bits, 52)))


def bin2d_raw(spec, x, y, w, N, Ts, sx, sw, device):
    """Integer binning: returns (mass int64 [Ts][cx][cy], count int32 [Ts][cx][cy]) on `device`."""
    L = _lib.lib()
    if spec.mode == 0:
        cx, cy = spec.geom.gx + 1, spec.geom.gy + 1
    else:
        cx, cy = spec.bx, spec.by
    mass = torch.zeros((Ts, cx, cy), dtype=torch.int64, device=device)
    count = torch.zeros((Ts, cx, cy), dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        _lib.check(L.dp_bin2d(ctypes.byref(spec), _lib.ptr(x), _lib.ptr(y), sx[0], sx[1], _lib.ptr(w), sw[0], sw[1], N, Ts,
                              _lib.ptr(mass), _lib.ptr(count), _lib.current_stream(device)), "dp_bin2d")
    return mass, count


def bin_occupancy(denv, x, y, w, args, mass_scale=None):
    """Time-sliced occupancy binning fused with the obstacle-grid dot product (dp_bin2d_occ).

    x, y, w: CUDA float32 tensors [Ts][N] (this library's time-major layout; w = normalised density per slice).
    Returns (mass int64 [Ts][G0+1][G1+1], count int32 [Ts][G0+1][G1+1], occ_dot float64 [Ts], scale_log2) where
    occ_dot[t] = sum_s w[t,s] * grid[cell(s), t] = sum_cells mass[t] * grid[:, :, t] / 2**scale_log2, the collision
    probability mass `check_collision` (motion_planning/utils.py:231-252) thresholds for a sum-binned density."""
    L = _lib.lib()
    Ts, N = x.shape
    dev = x.device
    spec = _lib.BinSpec()
    spec.mode = 0
    spec.geom = _geom(args)
    spec.mass_scale_log2 = mass_scale if mass_scale is not None else mass_scale_log2(N, float(w.abs().max()) if N > 0 else 1.0)
    cx, cy = spec.geom.gx + 1, spec.geom.gy + 1
    mass = torch.zeros((Ts, cx, cy), dtype=torch.int64, device=dev)
    count = torch.zeros((Ts, cx, cy), dtype=torch.int32, device=dev)
    dot = torch.empty(Ts, dtype=torch.float64, device=dev)
    assert x.is_contiguous() and y.is_contiguous() and w.is_contiguous()
    with torch.cuda.device(dev):
        _lib.check(L.dp_bin2d_occ(ctypes.byref(spec), denv.handle, _lib.ptr(x), _lib.ptr(y), 1, N, _lib.ptr(w), 1, N, N, Ts,
                                  _lib.ptr(mass), _lib.ptr(count), _lib.ptr(dot), _lib.current_stream(dev)), "dp_bin2d_occ")
    return mass, count, dot, spec.mass_scale_log2


def pred2grid(x, rho, args, return_gridpos=False):
    """Mean density per occupied cell, normalised to 1 (motion_planning/utils.py:255-280).

    x (N,>=2,1), rho (N,1,1) -> grid (G0,G1,1).  Cell indices and per-cell counts are exact integers; the density mass
    is accumulated in int64 fixed point (order-independent)."""
    dev = _cuda_device(x)
    N = x.shape[0]
    xs = x.detach()[:, 0, 0].to(device=dev, dtype=torch.float32).contiguous()
    ys = x.detach()[:, 1, 0].to(device=dev, dtype=torch.float32).contiguous()
    w = rho.detach()[:, 0, 0].to(device=dev, dtype=torch.float32).contiguous()
    spec = _lib.BinSpec()
    spec.mode = 0
    spec.geom = _geom(args)
    wmax = float(w.abs().max()) if N > 0 else 1.0
    spec.mass_scale_log2 = mass_scale_log2(N, wmax)
    mass, count = bin2d_raw(spec, xs, ys, w, N, 1, (1, 0), (1, 0), dev)
    G0, G1 = int(args.grid_size[0]), int(args.grid_size[1])
    if bool((count[0, G0, :] > 0).any()) or bool((count[0, :, G1] > 0).any()):
        # the reference clamps to G (not G-1) and then fails on the slice assignment (:262,:277); report it the same way
        raise RuntimeError("pred2grid: samples beyond the upper grid edge (the reference raises here as well)")
    cnt = count[0].to(torch.float64)
    mean = torch.where(count[0] > 0, mass[0].to(torch.float64) / (2.0 ** spec.mass_scale_log2) / cnt.clamp(min=1),
                       torch.zeros_like(cnt))
    grid = (mean / mean.sum()).to(torch.float32)[:G0, :G1].unsqueeze(-1).to(x.device)
    if return_gridpos:
        gx, gy = pos2gridpos(args, pos_x=x[:, 0, 0], pos_y=x[:, 1, 0])
        gridpos = torch.stack((gx.clamp(0, G0), gy.clamp(0, G1)), dim=1)
        return grid, gridpos
    return grid


def check_collision(grid_ego, grid_env, max_coll_sum, return_coll_pos=False):
    """Occupancy dot product (motion_planning/utils.py:231-252); plain tensor arithmetic on the caller's device."""
    coll_grid = grid_ego * grid_env
    coll_sum = coll_grid.sum()
    collision = bool(coll_sum > max_coll_sum)
    coll_pos_prob = None
    if collision and return_coll_pos:
        pos = coll_grid.nonzero(as_tuple=True)
        coll_pos_prob = torch.stack((pos[0], pos[1], coll_grid[pos[0], pos[1]]))
    return collision, coll_sum, coll_pos_prob


def get_density_map(x, rho, args, log_density=False, type="LE", bins=None, bin_width=None, system=None):
    """Error-state density histogram (systems/utils.py:105-129) -> (float64 numpy (bx,by), range_bins).

    Like the reference, mutates `rho` in place (rho -= max / rho /= max)."""
    if x.shape[1] != 2:
        raise NotImplementedError("only the 2-D map used by the reference's callers is on the hot path")
    if bins is None:
        bins = args.bin_number[:x.shape[1]]
    if bin_width is None:
        bin_width = (system.XE_MAX - system.XE_MIN)[0, :x.shape[1], 0] / bins
    if type != "MC" and log_density:
        rho -= rho.max()
        w = torch.exp(rho)
    else:
        rho /= rho.max()
        w = rho
    range_bins = [[float(-bins[i] * bin_width[i] / 2), float(bins[i] * bin_width[i] / 2)] for i in range(len(bins))]
    dev = _cuda_device(x)
    N = x.shape[0]
    xs = x.detach()[:, 0].to(device=dev, dtype=torch.float32).contiguous()
    ys = x.detach()[:, 1].to(device=dev, dtype=torch.float32).contiguous()
    wd = w.detach().reshape(N).to(device=dev, dtype=torch.float32).contiguous()
    spec = _lib.BinSpec()
    spec.mode = 1
    spec.lo_x, spec.hi_x = range_bins[0]
    spec.lo_y, spec.hi_y = range_bins[1]
    spec.bx, spec.by = int(bins[0]), int(bins[1])
    spec.mass_scale_log2 = mass_scale_log2(N, 1.0)
    mass, count = bin2d_raw(spec, xs, ys, wd, N, 1, (1, 0), (1, 0), dev)
    density = (mass[0].to(torch.float64) / (2.0 ** spec.mass_scale_log2)).cpu().numpy()
    mask = density > 0
    if type != "MC":
        density[mask] /= count[0].cpu().numpy()[mask]
    density[mask] = density[mask] / density[mask].sum()
    return density, range_bins


def normalize_density(rho_log_unnorm, rho0):
    """rho = exp(l - max_s l) * rho0 / sum_s(...) per time index (simulation_objects.py:295-299).

    rho_log_unnorm (N,1,T) (any strides; the kernel's [T][N] storage is used without a copy), rho0 (N,1,1)."""
    L = _lib.lib()
    dev = _cuda_device(rho_log_unnorm)
    N, T = rho_log_unnorm.shape[0], rho_log_unnorm.shape[2]
    lt = rho_log_unnorm.detach()[:, 0, :].to(device=dev, dtype=torch.float32).t()  # (T, N)
    if not lt.is_contiguous():
        lt = lt.contiguous()
    r0 = rho0.detach().reshape(N).to(device=dev, dtype=torch.float32).contiguous()
    lmax = torch.empty(T, dtype=torch.float32, device=dev)
    lsum = torch.empty(T, dtype=torch.float64, device=dev)
    out = torch.empty((T, N), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        s = _lib.current_stream(dev)
        _lib.check(L.dp_normalize_stats(_lib.ptr(lt), _lib.ptr(r0), N, T, N, _lib.ptr(lmax), _lib.ptr(lsum), s),
                   "dp_normalize_stats")
        _lib.check(L.dp_normalize_apply(_lib.ptr(lt), _lib.ptr(r0), N, T, N, _lib.ptr(lmax), _lib.ptr(lsum),
                                        _lib.ptr(out), s), "dp_normalize_apply")
    return out.t().unsqueeze(1).to(rho_log_unnorm.device)


class DeviceEnv:
    """Device copy of an environment's occupancy grid and its gradients (simulation_objects.py:29-44,77-94), repacked
    into the time-major gather table of the cost kernel.  Build once per environment."""

    def __init__(self, grid, grid_gradientX, grid_gradientY, args, device=None):
        L = _lib.lib()
        dev = torch.device(device) if device is not None else _cuda_device(grid)
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        self.args = args
        G0, G1, Tg = grid.shape
        if (G0, G1) != (int(args.grid_size[0]), int(args.grid_size[1])):
            raise _lib.DensityPlannerError("grid shape %s does not match args.grid_size %s" % ((G0, G1), args.grid_size))
        gs = [t.detach().to(device=dev, dtype=torch.float32).contiguous() for t in (grid, grid_gradientX, grid_gradientY)]
        g = _geom(args)
        h = ctypes.c_void_p()
        with torch.cuda.device(dev):
            torch.cuda.current_stream(dev).synchronize()
            _lib.check(L.dp_env_create(ctypes.byref(g), Tg, _lib.ptr(gs[0]), _lib.ptr(gs[1]), _lib.ptr(gs[2]), 1, dev.index,
                                       ctypes.byref(h)), "dp_env_create")
        self.handle = h
        self.Tg = Tg

    @classmethod
    def from_env(cls, env, args, device=None):
        """From a reference `Environment` object after get_gradient()."""
        return cls(env.grid, env.grid_gradientX, env.grid_gradientY, args, device)

    def close(self):
        """Release the device table (errors surface here; __del__ only guards against interpreter shutdown)."""
        if getattr(self, "handle", None):
            h, self.handle = self.handle, None
            _lib.lib().dp_env_destroy(h)

    def __del__(self, _finalizing=sys.is_finalizing):
        # at interpreter shutdown the driver context is going away with the process: nothing to release, nothing to import
        if _finalizing():
            return
        self.close()


def _cost_call(denv, x, y, rho, get_max, want_grad, per_sample):
    """x, y: 2-D (N, Ts) fp32 device views (any strides); rho likewise or None."""
    L = _lib.lib()
    dev = denv.device
    N, Ts = x.shape
    total = torch.zeros(1, dtype=torch.float64, device=dev)
    samples = torch.zeros(N, dtype=torch.float64, device=dev) if per_sample else None
    gx = gy = gr = None
    sg = (0, 0)
    if want_grad:
        # gradients in the layout of x (sample-contiguous if x is, else time-contiguous)
        if x.stride(0) == 1:
            gx = torch.empty((Ts, N), dtype=torch.float32, device=dev).t()
            gy = torch.empty((Ts, N), dtype=torch.float32, device=dev).t()
            gr = torch.empty((Ts, N), dtype=torch.float32, device=dev).t() if rho is not None else None
        else:
            gx = torch.empty((N, Ts), dtype=torch.float32, device=dev)
            gy = torch.empty((N, Ts), dtype=torch.float32, device=dev)
            gr = torch.empty((N, Ts), dtype=torch.float32, device=dev) if rho is not None else None
        sg = (gx.stride(0), gx.stride(1))
    assert x.stride() == y.stride()
    with torch.cuda.device(dev):
        _lib.check(L.dp_cost_coll(denv.handle, _lib.ptr(x), _lib.ptr(y), x.stride(0), x.stride(1), _lib.ptr(rho),
                                  rho.stride(0) if rho is not None else 0, rho.stride(1) if rho is not None else 0,
                                  N, Ts, 1 if get_max else 0, None, _lib.ptr(samples), _lib.ptr(gx), _lib.ptr(gy),
                                  _lib.ptr(gr), sg[0], sg[1], _lib.ptr(total), _lib.current_stream(dev)), "dp_cost_coll")
    return total, samples, gx, gy, gr


def _as_dev_2d(t, dev):
    """(N, Ts) fp32 view on `dev`, keeping strides when no copy is needed and one of them is 1."""
    t = t.detach()
    if t.device != dev or t.dtype != torch.float32:
        t = t.to(device=dev, dtype=torch.float32)
    if t.stride(0) != 1 and t.stride(1) != 1:
        t = t.contiguous()
    return t


class _CollisionCost(torch.autograd.Function):
    """cost = sum_i sum_s rho[s,0,i] * p(cell) * |des(cell) - x[s,:2,i]|^2 with grid indices constant (no_grad, :206)."""

    @staticmethod
    def forward(ctx, x_traj, rho_traj, denv, get_max):
        dev = denv.device
        Ts = x_traj.shape[2]
        x = _as_dev_2d(x_traj[:, 0, :], dev)
        y = _as_dev_2d(x_traj[:, 1, :], dev)
        if x.stride() != y.stride():
            x, y = x.contiguous(), y.contiguous()
        rho = _as_dev_2d(rho_traj[:, 0, :Ts], dev)
        need = (x_traj.requires_grad or rho_traj.requires_grad) and not get_max
        total, _, gx, gy, gr = _cost_call(denv, x, y, rho, get_max, need, False)
        if need:
            ctx.save_for_backward(gx, gy, gr)
        ctx.meta = (x_traj.shape, rho_traj.shape, x_traj.device, rho_traj.device, Ts, need)
        return total.to(dtype=torch.float32, device=x_traj.device)

    @staticmethod
    def backward(ctx, g):
        xs, rs, xdev, rdev, Ts, need = ctx.meta
        if not need:
            raise RuntimeError("backward through get_cost_coll(get_max=True) is not supported by the kernel path")
        gx, gy, gr = ctx.saved_tensors
        gxt = grt = None
        if ctx.needs_input_grad[0]:
            gxt = torch.zeros(xs, dtype=torch.float32, device=xdev)
            gxt[:, 0, :] = (gx * g.to(gx.device)).to(xdev)
            gxt[:, 1, :] = (gy * g.to(gy.device)).to(xdev)
        if ctx.needs_input_grad[1]:
            grt = torch.zeros(rs, dtype=torch.float32, device=rdev)
            grt[:, 0, :Ts] = (gr * g.to(gr.device)).to(rdev)
        return gxt, grt, None, None


def get_cost_coll(x_traj, rho_traj, denv, get_max=False):
    """MotionPlanner.get_cost_coll (motion_planning/MotionPlanner.py:192-224): returns a (1,) fp32 tensor on
    x_traj's device, differentiable w.r.t. x_traj[:, :2, :] and rho_traj (get_max=False)."""
    return _CollisionCost.apply(x_traj, rho_traj, denv, bool(get_max))


class _CollisionCostInit(torch.autograd.Function):
    """cost[n] = sum_i p(cell) * |des(cell) - x[n,:2,i]|^2 per trajectory (MotionPlannerGrad.py:309-336), grid indices
    constant (no_grad, :321-324): d cost[n] / d x[n,:2,i] = -2 p (des - x), evaluated by the same kernel pass."""

    @staticmethod
    def forward(ctx, x_traj, denv):
        dev = denv.device
        x = _as_dev_2d(x_traj[:, 0, :], dev)
        y = _as_dev_2d(x_traj[:, 1, :], dev)
        if x.stride() != y.stride():
            x, y = x.contiguous(), y.contiguous()
        need = x_traj.requires_grad
        _, samples, gx, gy, _ = _cost_call(denv, x, y, None, False, need, True)
        if need:
            ctx.save_for_backward(gx, gy)
        ctx.meta = (x_traj.shape, x_traj.device, need)
        return samples.to(dtype=torch.float32, device=x_traj.device)

    @staticmethod
    def backward(ctx, g):
        shape, xdev, need = ctx.meta
        if not need:
            return None, None
        gx, gy = ctx.saved_tensors
        gd = g.to(device=gx.device, dtype=torch.float32).reshape(-1, 1)
        gxt = torch.zeros(shape, dtype=torch.float32, device=xdev)
        gxt[:, 0, :] = (gx * gd).to(xdev)
        gxt[:, 1, :] = (gy * gd).to(xdev)
        return gxt, None


def get_cost_coll_initialize(x_traj, denv):
    """MotionPlannerGrad.get_cost_coll_initialize (motion_planning/MotionPlannerGrad.py:309-336): per-trajectory cost
    (num_traj,), no density weights; differentiable w.r.t. x_traj[:, :2, :] (the initialisation phase of the planner
    differentiates it w.r.t. the up parameters through up2ref_traj, :56-96,138)."""
    if x_traj.shape[0] == 0:
        return torch.zeros(0, dtype=torch.float32, device=x_traj.device)
    return _CollisionCostInit.apply(x_traj, denv)


# ---- goal / bounds cost terms of MotionPlanner.get_cost --------------------------------------
class _StateCost(torch.autograd.Function):
    """dp_cost_state / dp_cost_state_bwd on (N,5,Ts) x and (N,1,Ts) rho (any strides, CUDA tensors)."""

    @staticmethod
    def forward(ctx, x, rho, lo5, hi5, goal_xy, get_max):
        L = _lib.lib()
        dev = x.device
        N, _, Ts = x.shape
        xd = x.detach()
        rd = rho.detach()
        out = torch.empty(3, dtype=torch.float64, device=dev)
        flags = torch.empty(2, dtype=torch.int32, device=dev)
        lo = (ctypes.c_float * 5)(*[float(v) for v in lo5])
        hi = (ctypes.c_float * 5)(*[float(v) for v in hi5])
        goal = (ctypes.c_float * 2)(float(goal_xy[0]), float(goal_xy[1]))
        with torch.cuda.device(dev):
            _lib.check(L.dp_cost_state(_lib.ptr(xd), xd.stride(0), xd.stride(1), xd.stride(2), _lib.ptr(rd), rd.stride(0),
                                       rd.stride(2), N, Ts, lo, hi, goal, 1 if get_max else 0, _lib.ptr(out), _lib.ptr(flags),
                                       _lib.current_stream(dev)), "dp_cost_state")
        ctx.save_for_backward(xd, rd)
        ctx.consts = (lo, hi, goal, get_max)
        ctx.mark_non_differentiable(flags)
        return out.to(torch.float32), flags

    @staticmethod
    def backward(ctx, g_out, _g_flags):
        xd, rd = ctx.saved_tensors
        lo, hi, goal, get_max = ctx.consts
        if get_max:
            raise RuntimeError("the max variants of the goal / bounds costs are evaluation-only (as in the reference)")
        L = _lib.lib()
        dev = xd.device
        N, _, Ts = xd.shape
        gx = torch.empty((N, 5, Ts), dtype=torch.float32, device=dev)
        gr = torch.empty((N, 1, Ts), dtype=torch.float32, device=dev)
        s = [float(v) for v in g_out.detach().cpu()]
        with torch.cuda.device(dev):
            _lib.check(L.dp_cost_state_bwd(_lib.ptr(xd), xd.stride(0), xd.stride(1), xd.stride(2), _lib.ptr(rd), rd.stride(0),
                                           rd.stride(2), N, Ts, lo, hi, goal, s[0], s[1], s[2], _lib.ptr(gx), gx.stride(0),
                                           gx.stride(1), gx.stride(2), _lib.ptr(gr), gr.stride(0), gr.stride(2),
                                           _lib.current_stream(dev)), "dp_cost_state_bwd")
        return gx, gr, None, None, None, None


def state_costs(x_traj, rho_traj, lo5, hi5, goal_xy, get_max=False):
    """Raw goal / bounds terms: returns (terms, flags) with terms = [goal, below-bound part, above-bound part] (float32,
    differentiable w.r.t. x_traj and rho_traj in sum mode) and flags = [any x < lo, any x > hi] over state dims 0..3."""
    dev = _cuda_device(x_traj)
    if x_traj.dim() != 3 or x_traj.shape[1] != 5:
        raise ValueError("x_traj must be (N, 5, Ts): the kernels read the five CAR state dimensions (got %s)" % (tuple(x_traj.shape),))
    if rho_traj.dim() != 3 or rho_traj.shape[1] != 1 or rho_traj.shape[0] != x_traj.shape[0]:
        raise ValueError("rho_traj must be (N, 1, >=Ts) (got %s)" % (tuple(rho_traj.shape),))
    x = x_traj.to(device=dev, dtype=torch.float32)
    rho = rho_traj.to(device=dev, dtype=torch.float32)
    if rho.shape[2] < x.shape[2]:
        raise ValueError("rho_traj has fewer time points than x_traj")
    return _StateCost.apply(x, rho[:, :, :x.shape[2]] if rho.shape[2] != x.shape[2] else rho, lo5, hi5, goal_xy, bool(get_max))


def get_cost_goal(x_traj, rho_traj, xrefN, args, evaluate=False, get_max=False):
    """MotionPlanner.get_cost_goal (motion_planning/MotionPlanner.py:121-147) on the kernels."""
    big = [float("inf")] * 5
    # the reference weights the LAST state with the LAST density sample, rho_traj[:, 0, -1] (:139-141) -- rho_traj may be
    # longer than x_traj (the LE branch returns 101 states and 1001 densities), so the goal term is its own one-slice call
    terms, _ = state_costs(x_traj[:, :, -1:], rho_traj[:, :, -1:], [-v for v in big], big,
                           (float(xrefN[0, 0, 0]), float(xrefN[0, 1, 0])), get_max)
    cost_goal = terms[0]
    close = cost_goal < args.close2goal_thr
    if not evaluate and not bool(close):
        cost_goal = cost_goal * args.weight_goal_far
    return cost_goal.to(x_traj.device), close.to(x_traj.device)


def get_cost_bounds(x_traj, rho_traj, system, evaluate=False, get_max=False):
    """MotionPlanner.get_cost_bounds (motion_planning/MotionPlanner.py:149-190) on the kernels: returns (cost (1,), in_bounds)."""
    x_min, x_max = (system.X_MIN, system.X_MAX) if evaluate else (system.X_MIN_MP, system.X_MAX_MP)
    terms, flags = state_costs(x_traj, rho_traj, x_min[0, :, 0].tolist(), x_max[0, :, 0].tolist(), (0.0, 0.0), get_max)
    f = flags.cpu()
    # the reference enters a part only if a bound is violated in the first four state dimensions (:173,:181)
    cost = terms[1] * float(f[0] != 0) + terms[2] * float(f[1] != 0)
    in_bounds = torch.tensor(bool(f[0] == 0 and f[1] == 0))
    return cost.reshape(1).to(x_traj.device), in_bounds
