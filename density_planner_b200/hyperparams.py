"""Default settings namespace: the subset of the reference's hyperparams.py (parse_args, :8-178) the hot path reads,
with the same names and defaults, so callers without the reference checkout can build an `args` object."""
import types

import torch


def default_args(**overrides):
    a = types.SimpleNamespace(
        N_sim=1001, input_type="discr10", dt_sim=0.01, N_sim_max=1001, factor_pred=10, random_seed=0,
        samplesX_rawdata=20, samplesT_rawdata=20, size_rawdata=100, iterations_rawdata=500,
        bin_number=torch.Tensor([20, 20, 10, 10]).long(),
        path_rawdata="data/rawdata/2022-07-12_discr10_train/",
        mp_sample_size=500, mp_sampling="random", mp_gaussians=5,
        environment_size=[-12, 12, -30, 10], grid_wide=0.1,
        weight_coll=1e-1, weight_goal=1e-2, weight_uref=1e-4, weight_bounds=1e1,          # hyperparams.py:136-139
        weight_goal_far=10.0, weight_uref_effort=1e-2, close2goal_thr=3.0,                 # hyperparams.py:107-109
    )
    for k, v in overrides.items():
        setattr(a, k, v)
    es = a.environment_size
    a.grid_size = [int((es[1] - es[0]) / a.grid_wide) + 1, int((es[3] - es[2]) / a.grid_wide) + 1]
    return a
