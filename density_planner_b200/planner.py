"""Planner-facing layer: the LE branch of `EgoVehicle.predict_density` and the collision-cost calls of
`MotionPlanner` / `MotionPlannerGrad`, with the reference's signatures, on the CUDA library.

Reference surface mirrored here:
  motion_planning/simulation_objects.py  EgoVehicle.predict_density:253-304 (use_nn=False branch + normaliser)
  motion_planning/MotionPlanner.py       get_cost_coll:192-224
  motion_planning/MotionPlannerGrad.py   get_cost_coll_initialize:309-336

The density-NN branch (use_nn=True) and the other cost terms stay with the reference.
"""
import torch

from . import _lib, grid as _grid


class EgoPredictor:
    """The part of `EgoVehicle` the LE prediction needs: system, start state, sample set (xe0, rho0)."""

    def __init__(self, system, args, xref0, xe0, rho0):
        self.system = system
        self.args = args
        self.xref0 = xref0
        self.xe0 = xe0
        self.rho0 = rho0

    def predict_density(self, up, xref_traj, use_nn=False, xe0=None, rho0=None, compute_density=True):
        """LE prediction of the state and density trajectories (simulation_objects.py:253-304, use_nn=False).

        Returns x_traj (N,5,101) and rho_traj (N,1,1001): as in the reference, the states are subsampled by
        args.factor_pred while the density keeps the full time resolution (:293 vs :296-298)."""
        if use_nn:
            raise NotImplementedError("the density-NN predictor stays with the reference (SURVEY section 8f item 2)")
        if xe0 is None:
            xe0 = self.xe0
        if rho0 is None:
            rho0 = self.rho0
            assert rho0.shape[0] == xe0.shape[0]
        args = self.args
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
        init_ref = planner.get_cost_coll_initialize

        def _init(x_traj):
            # the initialisation phase differentiates through this cost w.r.t. up; keep the reference there
            return init_ref(x_traj) if x_traj.requires_grad else cc.get_cost_coll_initialize(x_traj)

        planner.get_cost_coll_initialize = _init
    system = Car(ego.args, device=device)
    pred = EgoPredictor(system, ego.args, ego.xref0, ego.xe0, ego.rho0)
    ref_predict = ego.predict_density

    def _predict(up, xref_traj, use_nn=True, xe0=None, rho0=None, compute_density=True):
        if use_nn:
            return ref_predict(up, xref_traj, use_nn=True, xe0=xe0, rho0=rho0, compute_density=compute_density)
        return pred.predict_density(up, xref_traj, use_nn=False, xe0=xe0, rho0=rho0, compute_density=compute_density)

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
