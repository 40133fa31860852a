"""density_planner_b200: B200-native (sm_100a) hot path of MIT-REALM/density_planner.

Monte-Carlo rollout of sampled initial states through the CAR closed loop under the C3M controller with Liouville
(log-)density, occupancy binning and the collision cost -- hand-written CUDA behind a C ABI
(include/density_planner_b200.h), with the reference's Python method surface on top.  No CPU fallback.
"""
from ._lib import DensityPlannerError, build, lib  # noqa: F401
from .hyperparams import default_args  # noqa: F401
from .systems import Car, ControlAffineSystem, Controller  # noqa: F401
from .grid import (pos2gridpos, gridpos2pos, pred2grid, check_collision, get_density_map, normalize_density, bin_occupancy,  # noqa: F401
                   DeviceEnv, get_cost_coll, get_cost_coll_initialize, state_costs, get_cost_goal, get_cost_bounds)
from .planner import EgoPredictor, CollisionCost, DensityNN, attach  # noqa: F401
from .rawdata import compute_data, compute_data_batched  # noqa: F401

__version__ = "0.1.0"
