"""System layer: call-compatible mirror of the reference's `ControlAffineSystem` / `Car`, backed by the CUDA library.

Reference surface mirrored here (names, argument meaning, shapes, prints, `None` returns):
  systems/ControlAffineSystem.py  u_func:26 f_func:43 fref_func:66 dudx_func:69 dfdx_func:93 divf_func:116
                                  get_next_x:124 get_next_xref:130 get_next_rho:136 get_next_rholog:147
                                  cut_xref_traj:159 sample_uref_traj:221 sample_xref0:335 compute_xref_traj:341
                                  up2ref_traj:370 sample_xe:381 sample_xe0:389 sample_x0:397 compute_density:409
                                  get_valid_ref:465 get_valid_trajectories:477
  systems/sytem_CAR.py            Controller:7 Car:40 (init_limits:57 a_func:128 dadx_func:145 b_func:163 dbdx_func:178)

What runs where: the closed-loop rollout with Liouville density (`compute_density`) and the single-step functions
run in hand-written sm_100a kernels through the C ABI (include/density_planner_b200.h).  Reference sampling and the
open-loop reference integration stay in torch on the CPU, consuming the torch global RNG in exactly the reference's
order, so seeded runs of `get_valid_trajectories` / `compute_data` reproduce the reference's samples.

Tensors may live on the CPU (like the reference's) or on the GPU; results come back on the device of `xe0` / `x`.
Trajectory results are (N, 5, T+1) / (N, 1, T+1) *views* of the kernel's time-major [T+1][5][N] storage.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .controller import DeviceController, load_blob


def _f32(t, device):
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


class Controller:
    """Clipped C3M tracking controller (systems/sytem_CAR.py:7-37)."""

    def __init__(self, device_controller, u_min, u_max):
        self.controller = device_controller
        self.dist_state = 2
        self.U_MIN = u_min
        self.U_MAX = u_max

    def __call__(self, x, xref, uref):
        return _step(self.controller, x, xref, uref, want=("u",))["u"]


class _XrefTraj(torch.autograd.Function):
    """dp_xref_traj / dp_xref_traj_bwd: xref0 (B or 1, 5, 1), uref (B, 2, T) on the GPU -> xref (B, 5, T+1)."""

    @staticmethod
    def forward(ctx, xref0, uref, dt):
        L = _lib.lib()
        B, _, T = uref.shape
        dev = uref.device
        batched = xref0.shape[0] == B and B > 1
        x0 = xref0.detach().to(torch.float32).reshape(-1, 5).contiguous()
        if not batched:
            x0 = x0[:1].contiguous()
        u = uref.detach().to(torch.float32).contiguous()
        out = torch.empty((B, 5, T + 1), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.dp_xref_traj(_lib.ptr(x0), 1 if batched else 0, _lib.ptr(u), dt, B, T, _lib.ptr(out),
                                      _lib.current_stream(dev)), "dp_xref_traj")
        ctx.save_for_backward(out)
        ctx.dt, ctx.x0_shape, ctx.batched = dt, xref0.shape, batched
        return out

    @staticmethod
    def backward(ctx, g):
        (xref,) = ctx.saved_tensors
        L = _lib.lib()
        B, _, T1 = xref.shape
        T = T1 - 1
        dev = xref.device
        g = g.to(torch.float32).contiguous()
        gu = torch.empty((B, 2, T), dtype=torch.float32, device=dev)
        gx0 = torch.empty((B, 5), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.dp_xref_traj_bwd(_lib.ptr(xref), _lib.ptr(g), ctx.dt, B, T, _lib.ptr(gu), _lib.ptr(gx0),
                                          _lib.current_stream(dev)), "dp_xref_traj_bwd")
        gx0 = gx0.reshape(ctx.x0_shape) if ctx.batched else gx0.sum(dim=0).reshape(ctx.x0_shape)
        return gx0, gu, None


def _step(dc, x, xref, uref, want):
    """One evaluation of the closed loop for N states sharing (xref, uref) -> dict of tensors on x's device."""
    L = _lib.lib()
    dev = dc.device
    out_dev = x.device
    N = x.shape[0]
    if xref.shape[0] != 1 or uref.shape[0] != 1:
        raise _lib.DensityPlannerError("only a shared reference (batch dimension 1) is supported")
    xd = _f32(x.reshape(N, 5), dev)
    xr = _f32(xref.reshape(5), dev)
    ur = _f32(uref.reshape(2), dev)
    bufs = {}
    shapes = {"u_raw": (N, 2), "u": (N, 2), "dudx": (N, 2, 5), "f": (N, 5), "div": (N,)}
    for k in shapes:
        bufs[k] = torch.empty(shapes[k], dtype=torch.float32, device=dev) if k in want else None
    with torch.cuda.device(dev):
        _lib.check(L.dp_step(dc.handle, _lib.ptr(xd), _lib.ptr(xr), _lib.ptr(ur), N, _lib.ptr(bufs["u_raw"]),
                             _lib.ptr(bufs["u"]), _lib.ptr(bufs["dudx"]), _lib.ptr(bufs["f"]), _lib.ptr(bufs["div"]),
                             _lib.current_stream(dev)), "dp_step")
    res = {}
    for k in want:
        v = bufs[k].to(out_dev)
        if k in ("u", "u_raw", "f"):
            v = v.unsqueeze(-1)
        res[k] = v
    return res


class ControlAffineSystem:
    """dx/dt = a(x) + b(x) u(x, xref, uref) with the learned contraction controller."""

    DIM_X = 5
    DIM_U = 2

    # ---- single-step functions -------------------------------------------------------------------------
    def u_func(self, x, xref, uref):
        return _step(self.controller.controller, x, xref, uref, ("u",))["u"]

    def f_func(self, x, xref, uref, noise=False):
        if noise:
            raise NotImplementedError("process noise is outside the hot path (no DIST is defined for CAR)")
        return _step(self.controller.controller, x, xref, uref, ("f",))["f"]

    def fref_func(self, xref, uref):
        return self.a_func(xref) + torch.bmm(self.b_func(xref), uref.type(torch.float32))

    def dudx_func(self, x, xref, uref):
        return _step(self.controller.controller, x, xref, uref, ("dudx",))["dudx"]

    def dfdx_func(self, x, xref, uref):
        # df/dx = da/dx + (db/dx) u + b du/dx ; db/dx = 0 for CAR (sytem_CAR.py:178-188)
        dudx = self.dudx_func(x, xref, uref)
        return self.dadx_func(x).to(dudx.device) + torch.bmm(self.b_func(x).to(dudx.device), dudx)

    def divf_func(self, x, xref, uref):
        return _step(self.controller.controller, x, xref, uref, ("div",))["div"]

    def get_next_x(self, x, xref, uref, dt):
        return x + self.f_func(x, xref, uref).to(x.device) * dt

    def get_next_xref(self, xref, uref, dt):
        return xref + self.fref_func(xref, uref) * dt

    def get_next_rho(self, x, xref, uref, rho, dt):
        divf = self.divf_func(x, xref, uref).to(rho.device)
        return rho + (-divf * rho) * dt

    def get_next_rholog(self, x, xref, uref, rholog, dt):
        divf = self.divf_func(x, xref, uref).to(rholog.device)
        return rholog + torch.log(1 - divf * dt)

    # ---- reference trajectories (CPU torch; RNG order as in the reference) ------------------------------------
    def cut_xref_traj(self, xref_traj, uref_traj):
        """Truncate at the first step where the reference leaves [X_MIN, X_MAX] (:159-193)."""
        n_cut = xref_traj.shape[2]
        for j in range(self.DIM_X):
            for bad in (xref_traj[0, j, :] > self.X_MAX[0, j, 0], xref_traj[0, j, :] < self.X_MIN[0, j, 0]):
                hit = bad.nonzero(as_tuple=True)[0]
                if hit.shape[0] > 0:
                    n_cut = min(n_cut, int(hit[0]))
        return xref_traj[:, :, :n_cut], uref_traj[:, :, :n_cut - 1]

    def sample_uref_traj(self, args, up=None):
        """Piecewise-constant input parametrisation `discr10` / `discr5` (:221-243); other input types of the
        reference (:245-333) are experiment variants outside the hot path."""
        if args.input_type == "discr10":
            n_u = 10
        elif args.input_type == "discr5":
            n_u = 5
        else:
            raise NotImplementedError("input_type %r: only discr10/discr5 are on the hot path" % args.input_type)
        seg = args.N_sim_max // n_u
        if up is None:
            up = 2 * torch.randn((1, self.DIM_U, n_u))
            up = up.clamp(self.UREF_MIN, self.UREF_MAX)
        elif up.dim() == 2:
            up = up.unsqueeze(0)
        uref_traj = torch.repeat_interleave(up, seg, dim=2)
        return uref_traj[:, :, :args.N_sim - 1], up

    def sample_xref0(self):
        return (self.XREF0_MAX - self.XREF0_MIN) * torch.rand(1, self.DIM_X, 1) + self.XREF0_MIN

    def compute_xref_traj(self, xref0, uref_traj, args, short=False):
        """Open-loop Euler integration of the reference (:341-354).

        CUDA inputs run in one kernel launch with an analytic (discrete adjoint) backward; CPU inputs keep the reference's
        torch loop, which the seeded sampling paths (`get_valid_ref`) rely on for bit-identical rejection decisions."""
        n_sim = min(args.N_sim, uref_traj.shape[2] + 1)
        if uref_traj.is_cuda:
            xref_traj = _XrefTraj.apply(xref0.to(uref_traj.device), uref_traj[:, :, :n_sim - 1], float(args.dt_sim))
            return xref_traj[:, :, ::args.factor_pred] if short else xref_traj
        states = [xref0]
        for i in range(n_sim - 1):
            states.append(self.get_next_xref(states[-1], uref_traj[:, :, [i]], args.dt_sim))
        xref_traj = torch.cat(states, dim=2)
        if short:
            return xref_traj[:, :, ::args.factor_pred]
        return xref_traj

    def up2ref_traj(self, xref0, up, args, short=True):
        uref_traj, _ = self.sample_uref_traj(args, up=up)
        xref_traj = self.compute_xref_traj(xref0, uref_traj, args)
        if short:
            return uref_traj[:, :, ::args.factor_pred], xref_traj[:, :, ::args.factor_pred]
        return uref_traj, xref_traj

    def sample_xe(self, param):
        return torch.rand(param, self.DIM_X, 1) * (self.XE_MAX - self.XE_MIN) + self.XE_MIN

    def sample_xe0(self, param):
        return torch.rand(param, self.DIM_X, 1) * (self.XE0_MAX - self.XE0_MIN) + self.XE0_MIN

    def sample_x0(self, xref0, sample_size):
        xe0_max = torch.minimum(self.X_MAX - xref0, self.XE0_MAX)
        xe0_min = torch.maximum(self.X_MIN - xref0, self.XE0_MIN)
        xe0 = torch.rand(sample_size, self.DIM_X, 1) * (xe0_max - xe0_min) + xe0_min
        return xe0 + xref0

    # ---- the hot path ------------------------------------------------------------------------------------------
    def compute_density(self, xe0, xref_traj, uref_traj, dt, rho0=None, cutting=True, compute_density=True,
                        log_density=False, stride=1, impl=None, return_margin=False):
        """Closed-loop rollout of N samples with Liouville (log-)density (:409-463) in one fused kernel launch.

        xe0 (N,5,1), xref_traj (1,5,T+1), uref_traj (1,2,T), rho0 (N,1,1) or None.
        Returns (xe_traj (N,5,Ts), rho_traj (N,1,Ts) or None) on xe0's device; Ts = T+1 unless `stride` > 1 (extension:
        store every stride-th step).  `return_margin` adds the per-sample distance of the raw controller output to the
        actuator clip (audit of clip-mask flips).  `impl="ffma"` evaluates the same rollout with the all-fp32 CUDA-core
        kernel of the DIAGNOSTICS library (the parity suite's fp32 cross-check; not a product backend)."""
        L = _lib.lib()
        dc = self.controller.controller
        dev = dc.device
        out_dev = xe0.device
        N, T = xe0.shape[0], uref_traj.shape[2]
        if xref_traj.shape[0] != 1 or uref_traj.shape[0] != 1:
            raise _lib.DensityPlannerError("only a shared reference (batch dimension 1) is supported")
        if xref_traj.shape[2] != T + 1:
            raise _lib.DensityPlannerError("xref_traj must have uref_traj.shape[2]+1 time points (got %d, %d)"
                                           % (xref_traj.shape[2], T))
        flags = (_lib.DP_LOG_DENSITY if log_density else 0) | (_lib.DP_DENSITY if compute_density else 0) | \
                (_lib.DP_CUT if cutting else 0)
        if impl not in (None, "tc", "ffma"):
            raise _lib.DensityPlannerError("unknown impl %r" % (impl,))
        Ts = T // stride + 1
        status = None
        if N == 0:
            xe_traj = torch.empty((0, 5, Ts), dtype=torch.float32, device=out_dev)
            rho_traj = torch.empty((0, 1, Ts), dtype=torch.float32, device=out_dev) if compute_density else None
            return (xe_traj, rho_traj, torch.empty(0, device=out_dev)) if return_margin else (xe_traj, rho_traj)
        if out_dev.type == "cuda" or impl == "ffma":
            xe0_d = _f32(xe0.reshape(N, 5), dev)
            xref_d = _f32(xref_traj.reshape(5, T + 1), dev)
            uref_d = _f32(uref_traj.reshape(2, T), dev)
            rho0_d = None if rho0 is None else _f32(rho0.reshape(N), dev)
            x_out = torch.empty((Ts, 5, N), dtype=torch.float32, device=dev)
            rho_out = torch.empty((Ts, N), dtype=torch.float32, device=dev) if compute_density else None
            margin = torch.empty(N, dtype=torch.float32, device=dev) if return_margin else None
            st = torch.zeros(2, dtype=torch.int32, device=dev)
            with torch.cuda.device(dev):
                if impl == "ffma":
                    _lib.check_diag(_lib.diag_lib().dp_diag_rollout_ffma(
                        dc.handle, _lib.ptr(xe0_d), _lib.ptr(xref_d), _lib.ptr(uref_d), _lib.ptr(rho0_d), float(dt), N, T,
                        flags, stride, _lib.ptr(x_out), _lib.ptr(rho_out), N, _lib.ptr(margin), _lib.ptr(st),
                        _lib.current_stream(dev)), "dp_diag_rollout_ffma")
                else:
                    _lib.check(L.dp_rollout(dc.handle, _lib.ptr(xe0_d), _lib.ptr(xref_d), _lib.ptr(uref_d), _lib.ptr(rho0_d),
                                            float(dt), N, T, flags, stride, _lib.ptr(x_out), _lib.ptr(rho_out), N,
                                            _lib.ptr(margin), _lib.ptr(st), _lib.current_stream(dev)), "dp_rollout")
            if compute_density and cutting:
                status = st.tolist()
            # results come back on the device of xe0 (another GPU than the controller's, or the CPU for the cross-check)
            if x_out.device != out_dev:
                x_out = x_out.to(out_dev)
                rho_out = rho_out.to(out_dev) if rho_out is not None else None
                margin = margin.to(out_dev) if margin is not None else None
        else:
            xe0_h = xe0.detach().reshape(N, 5).to(torch.float32).contiguous()
            xref_h = xref_traj.detach().reshape(5, T + 1).to(torch.float32).contiguous()
            uref_h = uref_traj.detach().reshape(2, T).to(torch.float32).contiguous()
            rho0_h = None if rho0 is None else rho0.detach().reshape(N).to(torch.float32).contiguous()
            pin = torch.cuda.is_available()
            x_out = torch.empty((Ts, 5, N), dtype=torch.float32, pin_memory=pin)
            rho_out = torch.empty((Ts, N), dtype=torch.float32, pin_memory=pin) if compute_density else None
            margin = torch.empty(N, dtype=torch.float32) if return_margin else None
            st_h = (ctypes.c_int * 2)()
            _lib.check(L.dp_rollout_host(dc.handle, _lib.ptr(xe0_h), _lib.ptr(xref_h), _lib.ptr(uref_h), _lib.ptr(rho0_h),
                                         float(dt), N, T, flags, stride, _lib.ptr(x_out), _lib.ptr(rho_out),
                                         _lib.ptr(margin), st_h), "dp_rollout_host")
            status = [int(st_h[0]), int(st_h[1])]
        if compute_density and cutting and status is not None:
            # the reference reports numeric trouble with prints, never with exceptions (:451-462)
            if status[0]:
                print("clamp rho_traj to 1e30 (log density)" if log_density
                      else "clamp rho_traj between 0 and 1e30 (no log density)")
            if status[1]:
                print("set nan in rho_traj to 1e30")
        xe_traj = x_out.permute(2, 1, 0)
        rho_traj = rho_out.unsqueeze(1).permute(2, 1, 0) if compute_density else None
        if return_margin:
            return xe_traj, rho_traj, margin
        return xe_traj, rho_traj

    def compute_density_batch(self, xe0_list, xref_list, uref_list, dt, cutting=True, compute_density=True,
                              log_density=False, stride=1):
        """`compute_density` for R references in ONE kernel launch (dp_rollout_batch): reference r has its own
        xref (1,5,T_r+1) / uref (1,2,T_r) and initial errors xe0 (n,5,1) (the same n for all).  This is how the raw-data
        generator's 20-sample trajectories (data_generation/compute_rawdata.py:26) fill the GPU.
        Returns a list of (xe_traj (n,5,Ts_r), rho_traj (n,1,Ts_r) or None), each exactly what `compute_density` returns
        for that reference alone."""
        L = _lib.lib()
        dc = self.controller.controller
        dev = dc.device
        R = len(xe0_list)
        if R == 0:
            return []
        out_dev = xe0_list[0].device
        n = xe0_list[0].shape[0]
        Ts_ref = [int(u.shape[2]) for u in uref_list]
        T = max(Ts_ref)
        xref = torch.zeros((R, 5, T + 1), dtype=torch.float32)
        uref = torch.zeros((R, 2, T), dtype=torch.float32)
        for r in range(R):
            if xe0_list[r].shape[0] != n or xref_list[r].shape[2] != Ts_ref[r] + 1:
                raise _lib.DensityPlannerError("compute_density_batch: inconsistent shapes for reference %d" % r)
            xref[r, :, :Ts_ref[r] + 1] = xref_list[r][0].detach().cpu()
            xref[r, :, Ts_ref[r] + 1:] = xref[r, :, Ts_ref[r]:Ts_ref[r] + 1]
            uref[r, :, :Ts_ref[r]] = uref_list[r][0].detach().cpu()
        xe0 = torch.cat([x.detach().reshape(n, 5).to(torch.float32).cpu() for x in xe0_list], dim=0)
        flags = (_lib.DP_LOG_DENSITY if log_density else 0) | (_lib.DP_DENSITY if compute_density else 0) | \
                (_lib.DP_CUT if cutting else 0)
        Ts = T // stride + 1
        N = R * n
        xe0_d, xref_d, uref_d = xe0.to(dev), xref.to(dev), uref.to(dev)
        t_ref = torch.tensor(Ts_ref, dtype=torch.int32, device=dev)
        x_out = torch.empty((Ts, 5, N), dtype=torch.float32, device=dev)
        rho_out = torch.empty((Ts, N), dtype=torch.float32, device=dev) if compute_density else None
        st = torch.zeros(2, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.dp_rollout_batch(dc.handle, _lib.ptr(xe0_d), _lib.ptr(xref_d), _lib.ptr(uref_d), _lib.ptr(t_ref), None,
                                          float(dt), R, n, T, flags, stride, _lib.ptr(x_out), _lib.ptr(rho_out), N, None,
                                          _lib.ptr(st), _lib.current_stream(dev)), "dp_rollout_batch")
        if compute_density and cutting:
            status = st.tolist()
            if status[0]:
                print("clamp rho_traj to 1e30 (log density)" if log_density
                      else "clamp rho_traj between 0 and 1e30 (no log density)")
            if status[1]:
                print("set nan in rho_traj to 1e30")
        x_all = x_out.to(out_dev)
        r_all = rho_out.to(out_dev) if compute_density else None
        res = []
        for r in range(R):
            ts = Ts_ref[r] // stride + 1
            xe = x_all[:ts, :, r * n:(r + 1) * n].permute(2, 1, 0)
            rho = r_all[:ts, r * n:(r + 1) * n].unsqueeze(1).permute(2, 1, 0) if compute_density else None
            res.append((xe, rho))
        return res

    def get_valid_ref(self, args):
        """Rejection-sample a reference that stays admissible for > 99% of the horizon (:465-475)."""
        while True:
            uref_traj, up = self.sample_uref_traj(args)
            xref0 = self.sample_xref0()
            xref_traj = self.compute_xref_traj(xref0, uref_traj, args)
            xref_traj, uref_traj = self.cut_xref_traj(xref_traj, uref_traj)
            if xref_traj.shape[2] > 0.99 * args.N_sim:
                return up, uref_traj, xref_traj

    def get_valid_trajectories(self, sample_size, args, plot=False, log_density=True, compute_density=True):
        """Reference + sampled initial errors + fused rollout, subsampled by args.factor_pred (:477-495)."""
        up, uref_traj, xref_traj = self.get_valid_ref(args)
        xe0 = self.sample_xe0(sample_size)
        xe_traj, rho_traj = self.compute_density(xe0, xref_traj, uref_traj, args.dt_sim, cutting=True,
                                                 log_density=log_density, compute_density=compute_density)
        t = args.dt_sim * torch.arange(0, xe_traj.shape[2])
        if plot:
            raise NotImplementedError("plotting is outside the hot path (plots/plot_functions.py)")
        k = args.factor_pred
        return (xref_traj[:, :, ::k], None if rho_traj is None else rho_traj[:, :, ::k], uref_traj[:, :, ::k], up,
                xe_traj[:, :, ::k], t[::k])


class Car(ControlAffineSystem):
    """Kinematic car with steering bias state: x = (px, py, theta, v, d) (systems/sytem_CAR.py:40-200)."""

    def __init__(self, args=None, device=None, controller_path=None, blob=None):
        self.init_limits(args)
        self.device = device  # None = current CUDA device at first use
        self.controller = self.controller_wrapper(controller_path, blob)
        self.systemname = "CAR"

    def init_limits(self, args):
        dist_max = np.pi / 8
        col = lambda v: torch.tensor(v, dtype=torch.float32).reshape(1, -1, 1)
        self.X_MIN = col([-50., -50., -np.pi + 0.2, 0, -dist_max])
        self.X_MAX = col([50., 50., 3 * np.pi - 0.2, 10, dist_max])
        self.XE_MIN = col([-2, -2, -1, -1, -dist_max])
        self.XE_MAX = col([2, 2, 1, 1, dist_max])
        self.UREF_MIN = col([-3., -3.])
        self.UREF_MAX = col([3., 3.])
        self.U_MIN = col([-3., -3.])
        self.U_MAX = col([3., 3.])
        if args is not None:
            es = args.environment_size
            self.XREF0_MIN = col([es[0], es[2], 0, 1., 0])
            self.XREF0_MAX = col([es[1], es[3], 2 * np.pi, 8., 0])
        else:
            self.XREF0_MIN = col([-10., -30., 0, 1, 0])
            self.XREF0_MAX = col([10., 30., 2 * np.pi, 8, 0])
        self.XE0_MIN = 0.5 * self.XE_MIN
        self.XE0_MAX = 0.5 * self.XE_MAX
        self.X_MIN_MP = col([-9.9, -29.9, -np.pi + 0.2, 0.1, -np.inf])
        self.X_MAX_MP = col([9.9, 9.9, 3 * np.pi - 0.2, 9.9, np.inf])

    def controller_wrapper(self, controller_path=None, blob=None):
        if blob is None:
            blob = load_blob(controller_path)
        if float(self.U_MIN.abs().max()) != 3.0 or float(self.U_MAX.abs().max()) != 3.0:
            raise _lib.DensityPlannerError("the kernels hard-code the +-3 actuator limits of sytem_CAR.py:75-76")
        return Controller(DeviceController(blob, self.device), self.U_MIN, self.U_MAX)

    def load_controller(self):
        return self.controller.controller

    def a_func(self, x):
        a = torch.zeros(x.shape[0], self.DIM_X, 1, dtype=torch.float32, device=x.device)
        a[:, 0, 0] = x[:, 3, 0] * torch.cos(x[:, 2, 0])
        a[:, 1, 0] = x[:, 3, 0] * torch.sin(x[:, 2, 0])
        return a

    def dadx_func(self, x):
        d = torch.zeros(x.shape[0], self.DIM_X, self.DIM_X, dtype=torch.float32, device=x.device)
        d[:, 0, 2] = -x[:, 3, 0] * torch.sin(x[:, 2, 0])
        d[:, 0, 3] = torch.cos(x[:, 2, 0])
        d[:, 1, 2] = x[:, 3, 0] * torch.cos(x[:, 2, 0])
        d[:, 1, 3] = torch.sin(x[:, 2, 0])
        return d

    def b_func(self, x):
        b = torch.zeros(x.shape[0], self.DIM_X, self.DIM_U, dtype=torch.float32, device=x.device)
        b[:, 2, 0] = 1
        b[:, 3, 1] = 1
        return b

    def dbdx_func(self, x):
        return torch.zeros(x.shape[0], self.DIM_X, self.DIM_U, self.DIM_X, dtype=torch.float32, device=x.device)

    def project_angle(self, x_traj):
        with torch.no_grad():
            x_traj[:, 2, :] = (x_traj[:, 2, :] + np.pi) % (2 * np.pi) - np.pi
        return x_traj
