"""Planner-facing layer: the LE branch of `EgoVehicle.predict_density` and the collision-cost calls of
`MotionPlanner` / `MotionPlannerGrad`, with the reference's signatures, on the CUDA library.

Reference surface mirrored here:
  motion_planning/simulation_objects.py  EgoVehicle.predict_density:253-304 (use_nn=False branch + normaliser)
  motion_planning/MotionPlanner.py       get_cost_coll:192-224
  motion_planning/MotionPlannerGrad.py   get_cost_coll_initialize:309-336

The density-NN branch (use_nn=True) is evaluated for all time points in one launch of the fused tcgen05 MLP kernel
(`DensityNN`, csrc/dnn.cu); the goal / bounds cost terms are kernel backed (grid.get_cost_goal / get_cost_bounds).
"""
import torch

from . import _lib, grid as _grid


class _DnnFunction(torch.autograd.Function):
    """y = MLP(x) on the fused tcgen05 kernel (dp_dnn_forward); backward = gradient w.r.t. the rows x (dp_dnn_backward)."""

    @staticmethod
    def forward(ctx, x, nn):
        L = _lib.lib()
        dev = nn.device
        rows = x.shape[0]
        xc = x.detach().to(device=dev, dtype=torch.float32).contiguous()
        y = torch.empty((rows, nn.dims[-1]), dtype=torch.float32, device=dev)
        need = x.requires_grad
        masks = torch.empty((rows, nn.mask_words), dtype=torch.int32, device=dev) if (need and nn.mask_words) else None
        with torch.cuda.device(dev):
            _lib.check(L.dp_dnn_forward(nn.handle, _lib.ptr(xc), rows, _lib.ptr(y), _lib.ptr(masks), _lib.current_stream(dev)),
                       "dp_dnn_forward")
        ctx.nn, ctx.masks, ctx.need = nn, masks, need
        return y

    @staticmethod
    def backward(ctx, gy):
        if not ctx.need:
            return None, None
        nn = ctx.nn
        L = _lib.lib()
        dev = nn.device
        g = gy.to(device=dev, dtype=torch.float32).contiguous()
        gx = torch.empty((g.shape[0], nn.dims[0]), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.dp_dnn_backward(nn.handle, _lib.ptr(g), _lib.ptr(ctx.masks), g.shape[0], _lib.ptr(gx),
                                         _lib.current_stream(dev)), "dp_dnn_backward")
        return gx, None


class DensityNN:
    """The reference's density-predictor MLP (density_training/utils.py:8-40: 31 -> 7 x 150 ReLU -> 6) evaluated for ALL
    prediction time points in one batched pass instead of one call per time point
    (simulation_objects.py:280-287 calls get_nn_prediction 101 times; density_training/utils.py:161-180).

    backend "kernel" (default): the fused tcgen05 kernel csrc/dnn.cu through dp_dnn_forward / dp_dnn_backward (all layers per
    128-row tile on the SM, split-fp16 operands with fp32-level accuracy), differentiable w.r.t. the rows, i.e. w.r.t. `up` and
    `xe0` like the reference's branch.  It needs a B200 and raises without one -- there is no CPU fallback.
    backend "torch": the same batched formulation in plain torch (addmm) -- the cross-check the CPU tests replay against the
    reference golden."""

    def __init__(self, weights=None, device=None, activation="relu", backend="kernel"):
        import ctypes
        import os
        import numpy as np
        if activation not in ("relu", "tanh"):
            raise _lib.DensityPlannerError("DensityNN mirrors the reference's MLP with relu or tanh (got %r)" % (activation,))
        if backend not in ("kernel", "torch"):
            raise _lib.DensityPlannerError("DensityNN backend is 'kernel' or 'torch' (got %r)" % (backend,))
        self.activation = activation
        if weights is None:
            weights = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "density_nn_default.npz")
        z = np.load(weights) if isinstance(weights, str) else weights
        n_layers = len([k for k in z.keys() if k.endswith(".weight")])
        Wn = [np.ascontiguousarray(np.asarray(z["linear_list.%d.weight" % i], dtype="float32")) for i in range(n_layers)]
        bn = [np.ascontiguousarray(np.asarray(z["linear_list.%d.bias" % i], dtype="float32")) for i in range(n_layers)]
        self.dims = [int(Wn[0].shape[1])] + [int(w.shape[0]) for w in Wn]
        self.handle, self.mask_words = None, 0
        self.backend = backend
        if backend == "kernel":
            if not torch.cuda.is_available():
                raise _lib.DensityPlannerError("DensityNN(backend='kernel') needs a B200; there is no CPU fallback "
                                               "(backend='torch' is the CPU cross-check formulation)")
            if device is None or torch.device(device).type != "cuda":
                device = torch.device("cuda", torch.cuda.current_device())
            dev = torch.device(device)
            if dev.index is None:
                dev = torch.device("cuda", torch.cuda.current_device())
            L = _lib.lib()
            wp = (ctypes.c_void_p * n_layers)(*[w.ctypes.data for w in Wn])
            bp = (ctypes.c_void_p * n_layers)(*[b.ctypes.data for b in bn])
            dims = (ctypes.c_int * (n_layers + 1))(*self.dims)
            h = ctypes.c_void_p()
            _lib.check(L.dp_dnn_create(wp, bp, dims, n_layers, 1 if activation == "tanh" else 0, dev.index, ctypes.byref(h)),
                       "dp_dnn_create")
            self.handle = h
            mw = ctypes.c_int()
            _lib.check(L.dp_dnn_mask_words(h, ctypes.byref(mw)), "dp_dnn_mask_words")
            self.mask_words = int(mw.value)
        else:
            dev = torch.device(device) if device is not None else torch.device("cpu")
        self.W = [torch.from_numpy(w).to(dev) for w in Wn]
        self.b = [torch.from_numpy(b).to(dev) for b in bn]
        self.device = dev

    def close(self):
        if getattr(self, "handle", None):
            h, self.handle = self.handle, None
            _lib.lib().dp_dnn_destroy(h)

    def __del__(self, _finalizing=__import__("sys").is_finalizing):
        if _finalizing():
            return
        try:
            self.close()
        except Exception:
            pass

    def mlp(self, x):
        """rows (R, 31) -> (R, 6)."""
        if x.shape[1] != self.dims[0]:
            raise _lib.DensityPlannerError("the density NN expects %d inputs per row, got %d" % (self.dims[0], x.shape[1]))
        if self.backend == "kernel":
            return _DnnFunction.apply(x, self)
        act = torch.relu if self.activation == "relu" else torch.tanh
        for i in range(len(self.W) - 1):
            x = act(torch.addmm(self.b[i], x, self.W[i].t()))
        return torch.addmm(self.b[-1], x, self.W[-1].t())

    def predict(self, xe0, xref0, up, t_vec):
        """xe0 (N,5), xref0 (5,), up (1,2,10) or flat (20,), t_vec (Tn,) -> xe (N,5,Tn), rholog (N,1,Tn).

        Input row layout of data_generation/utils.py:4-37,101-135: [xe0 (5) | xref0 (5) | t | u_params (20)]."""
        dev = self.device
        N, Tn = xe0.shape[0], t_vec.shape[0]
        xe0 = xe0.to(dev, torch.float32)
        head = torch.cat((xe0, xref0.to(dev, torch.float32).reshape(1, -1).expand(N, -1)), dim=1)       # (N, 10)
        upf = up.to(dev, torch.float32).reshape(1, -1).expand(N, -1)                                     # (N, 20)
        t = t_vec.to(dev, torch.float32)
        rows = torch.cat((head.unsqueeze(0).expand(Tn, -1, -1), t.reshape(Tn, 1, 1).expand(-1, N, -1),
                          upf.unsqueeze(0).expand(Tn, -1, -1)), dim=2).reshape(Tn * N, -1)                # (Tn*N, 31)
        if rows.shape[1] != self.dims[0]:
            raise _lib.DensityPlannerError("the density NN expects %d inputs per row, the discr10 row layout has %d"
                                           % (self.dims[0], rows.shape[1]))
        y = self.mlp(rows).reshape(Tn, N, -1)                                                            # (Tn, N, 6)
        return y[:, :, :5].permute(1, 2, 0), y[:, :, 5:6].permute(1, 2, 0)


class EgoPredictor:
    """The part of `EgoVehicle` the LE prediction needs: system, start state, sample set (xe0, rho0)."""

    def __init__(self, system, args, xref0, xe0, rho0, density_nn=None):
        self.system = system
        self.args = args
        self.xref0 = xref0
        self.xe0 = xe0
        self.rho0 = rho0
        self.density_nn = density_nn  # DensityNN for the use_nn=True branch (optional)

    def predict_density(self, up, xref_traj, use_nn=False, xe0=None, rho0=None, compute_density=True):
        """LE prediction of the state and density trajectories (simulation_objects.py:253-304, use_nn=False).

        Returns x_traj (N,5,101) and rho_traj (N,1,1001): as in the reference, the states are subsampled by
        args.factor_pred while the density keeps the full time resolution (:293 vs :296-298)."""
        if xe0 is None:
            xe0 = self.xe0
        if rho0 is None:
            rho0 = self.rho0
            assert rho0.shape[0] == xe0.shape[0]
        args = self.args
        if use_nn:
            if self.density_nn is None:
                raise NotImplementedError("no DensityNN attached to this predictor (pass density_nn=DensityNN(...))")
            n_sim = args.N_sim
            if args.input_type == "discr10" and up.shape[2] != 10:
                n_sim = up.shape[2] * (args.N_sim_max // 10) + 1
                up = torch.cat((up, torch.zeros(up.shape[0], up.shape[1], 10 - up.shape[2], device=up.device)), dim=2)
            elif args.input_type == "discr5" and up.shape[2] != 5:
                n_sim = up.shape[2] * (args.N_sim_max // 5) + 1
                up = torch.cat((up, torch.zeros(up.shape[0], up.shape[1], 5 - up.shape[2], device=up.device)), dim=2)
            t_vec = torch.arange(0, args.dt_sim * n_sim - 0.001, args.dt_sim * args.factor_pred)
            out_dev = xref_traj.device
            xe_nn, rholog_nn = self.density_nn.predict(xe0[:, :, 0], self.xref0[0, :, 0], up, t_vec)
            Tn = xref_traj.shape[2]
            # the reference fills (N, 5, Tn) / (N, 1, Tn) buffers of the reference's length time point by time point (:281-287)
            xe_traj = torch.zeros(xe0.shape[0], 5, Tn, device=xe_nn.device)
            rho_log = torch.zeros(xe0.shape[0], 1, Tn, device=xe_nn.device)
            k = min(Tn, xe_nn.shape[2])
            xe_traj[:, :, :k] = xe_nn[:, :, :k]
            rho_log[:, :, :k] = rholog_nn[:, :, :k]
            rho_traj = None
            if compute_density:
                rho_max, _ = rho_log.max(dim=0)
                rho_un = torch.exp(rho_log - rho_max.unsqueeze(0)) * rho0.to(rho_log.device).reshape(-1, 1, 1)
                rho_traj = (rho_un / rho_un.sum(dim=0).unsqueeze(0)).to(out_dev)
            return (xe_traj + xref_traj.to(xe_traj.device)).to(out_dev), rho_traj
        if args.input_type == "discr10" and up.shape[2] != 10:
            up = torch.cat((up, torch.zeros(up.shape[0], up.shape[1], 10 - up.shape[2])), dim=2)
        elif args.input_type == "discr5" and up.shape[2] != 5:
            up = torch.cat((up, torch.zeros(up.shape[0], up.shape[1], 5 - up.shape[2])), dim=2)
        uref_traj, _ = self.system.sample_uref_traj(args, up=up)
        xref_traj_long = self.system.compute_xref_traj(self.xref0, uref_traj, args)
        xe_traj_long, rho_log_unnorm = self.system.compute_density(xe0, xref_traj_long, uref_traj, args.dt_sim,
                                                                   cutting=False, log_density=True, compute_density=True)
        xe_traj = xe_traj_long[:, :, ::args.factor_pred]
        rho_traj = _grid.normalize_density(rho_log_unnorm, rho0) if compute_density else None
        x_traj = xe_traj + xref_traj.to(xe_traj.device)
        return x_traj, rho_traj


class CollisionCost:
    """Kernel-backed `get_cost_coll` / `get_cost_coll_initialize` for one environment."""

    def __init__(self, env_grid, env_grid_gradientX, env_grid_gradientY, args, device=None):
        self.denv = _grid.DeviceEnv(env_grid, env_grid_gradientX, env_grid_gradientY, args, device)
        self.args = args

    @classmethod
    def from_ego(cls, ego, device=None):
        """From a reference EgoVehicle (its env must have get_gradient() applied, as EgoVehicle.__init__ does)."""
        return cls(ego.env.grid, ego.env.grid_gradientX, ego.env.grid_gradientY, ego.args, device)

    def get_cost_coll(self, x_traj, rho_traj, get_max=False):
        return _grid.get_cost_coll(x_traj, rho_traj, self.denv, get_max=get_max)

    def get_cost_coll_initialize(self, x_traj):
        return _grid.get_cost_coll_initialize(x_traj, self.denv)


def attach(planner, device=None):
    """Bind the kernel-backed calls into a reference planner object in place:
    `planner.get_cost_coll`, `planner.get_cost_coll_initialize` (if present), `planner.get_cost_goal`,
    `planner.get_cost_bounds`, the LE branch of `planner.ego.predict_density` and `ego.system.up2ref_traj` (forward and
    analytic backward in one launch each).  The planner's Python optimisation loop is unchanged."""
    from .systems import Car

    ego = planner.ego
    cc = CollisionCost.from_ego(ego, device)
    planner.get_cost_coll = cc.get_cost_coll
    if hasattr(planner, "get_cost_coll_initialize"):
        # forward + analytic backward on the kernel: the whole initialisation phase (MotionPlannerGrad.py:56-96) runs on it
        planner.get_cost_coll_initialize = cc.get_cost_coll_initialize
    system = Car(ego.args, device=device)
    nn = None
    model = getattr(ego, "model", None)
    mirrored = (getattr(ego.args, "nn_type", "MLP") == "MLP" and getattr(ego.args, "activation", "relu") in ("relu", "tanh")
                and getattr(ego.args, "input_type", "discr10") in ("discr10", "discr5"))
    if model is not None and hasattr(model, "state_dict") and mirrored:
        # the ego's own density-NN weights, evaluated for all time points in one batched pass on the GPU; any other
        # configuration (LSTM, other input rows) keeps the reference's branch (ref_predict below)
        nn = DensityNN(weights={k: v.detach().cpu().numpy() for k, v in model.state_dict().items()},
                       device=system.controller.controller.device, activation=getattr(ego.args, "activation", "relu"))
    pred = EgoPredictor(system, ego.args, ego.xref0, ego.xe0, ego.rho0, density_nn=nn)
    ref_predict = ego.predict_density

    def _predict(up, xref_traj, use_nn=True, xe0=None, rho0=None, compute_density=True):
        if use_nn and nn is None:
            return ref_predict(up, xref_traj, use_nn=True, xe0=xe0, rho0=rho0, compute_density=compute_density)
        return pred.predict_density(up, xref_traj, use_nn=use_nn, xe0=xe0, rho0=rho0, compute_density=compute_density)

    ego.predict_density = _predict

    # goal / bounds terms (MotionPlanner.py:121-190) and the reference-trajectory integrator (ControlAffineSystem.py:370):
    # same signatures, tensors come back on the caller's device with autograd intact
    def _goal(x_traj, rho_traj, evaluate=False, get_max=False):
        if rho_traj is None:
            return goal_ref(x_traj, rho_traj, evaluate=evaluate, get_max=get_max)
        return _grid.get_cost_goal(x_traj, rho_traj, ego.xrefN, ego.args, evaluate=evaluate, get_max=get_max)

    def _bounds(x_traj, rho_traj, evaluate=False, get_max=False):
        return _grid.get_cost_bounds(x_traj, rho_traj, system, evaluate=evaluate, get_max=get_max)

    goal_ref = planner.get_cost_goal
    planner.get_cost_goal = _goal
    planner.get_cost_bounds = _bounds
    dev = system.controller.controller.device

    def _up2ref(xref0, up, args, short=True):
        uref_traj, xref_traj = system.up2ref_traj(xref0.to(dev), up.to(dev), args, short=short)
        return uref_traj.to(up.device), xref_traj.to(up.device)

    ego.system.up2ref_traj = _up2ref
    planner._dp_b200 = (cc, pred)
    return planner
