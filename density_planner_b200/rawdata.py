"""Raw-data generator entry: mirror of data_generation/compute_rawdata.py:9-63 (`compute_data`).

The loop, the torch RNG order (references, initial errors, then the time-index draw at :34) and the result
dictionary schema are the reference's; the rollout inside `get_valid_trajectories` is the fused CUDA kernel.
"""
import pickle
from datetime import datetime

import torch


def compute_data(iteration_number, samples_x, system, args, samples_t=None, save=True, plot=True):
    """Generate `iteration_number[0]` files of `iteration_number[1]` trajectories with `samples_x` samples each.

    The rollout consumes no random numbers and the number of time points follows from the reference alone, so drawing all
    references / initial errors / time indices of a file first (in the reference's order) and rolling them out in ONE launch
    (`compute_data_batched`: references of at most 64 samples share the 128-lane tiles) gives the same result dictionaries as
    the reference's trajectory-by-trajectory loop.  That is what this entry does unless the system overrides
    `get_valid_trajectories` (then the loop below calls it per trajectory, as the reference does)."""
    from .systems import ControlAffineSystem
    if not plot and getattr(type(system), "get_valid_trajectories", None) is ControlAffineSystem.get_valid_trajectories \
            and hasattr(system, "compute_density_batch"):
        return compute_data_batched(iteration_number, samples_x, system, args, samples_t=samples_t, save=save)
    results_all = []
    for _ in range(iteration_number[0]):
        for _ in range(iteration_number[1]):
            xref_traj, rholog_traj, uref_traj, u_params, xe_traj, t = system.get_valid_trajectories(samples_x, args)
            if samples_t == 0:
                indizes = args.N_sim - 1
            elif samples_t is not None:
                indizes = torch.randint(0, t.shape[0], (int(samples_t),))
                indizes = list(set(indizes.tolist()))
            else:
                indizes = torch.arange(0, t.shape[0])
            results_all.append({
                'u_params': u_params,
                'xe0': xe_traj[:, :, 0],
                't': t[indizes],
                'xref0': xref_traj[0, :, 0],
                'xe_traj': xe_traj[:, :, indizes],
                'rholog_traj': rholog_traj[:, :, indizes],
            })
            if plot:
                raise NotImplementedError("plotting is outside the hot path (the reference's plot branch at "
                                          "compute_rawdata.py:47-51 references an undefined name and cannot run either)")
        if save:
            data_name = args.path_rawdata + datetime.now().strftime("%Y-%m-%d-%H-%M-%S") + "_rawData_" + \
                system.systemname + "_%s_Nsim%d_iter%d_xSamples%d_tSamples%d" % (
                    args.input_type, args.N_sim, iteration_number[1], samples_x, samples_t) + ".pickle"
            print("save " + data_name)
            with open(data_name, "wb") as f:
                pickle.dump([results_all, args], f)
            results_all = []
    return results_all


def compute_data_batched(iteration_number, samples_x, system, args, samples_t=None, save=True):
    """`compute_data` with ONE rollout launch per file instead of one per trajectory.

    The torch RNG is consumed in exactly the reference's order -- per trajectory: reference rejection sampling, initial
    errors, time-index draw (compute_rawdata.py:26-36; the rollout itself consumes none and the number of time points
    follows from the reference alone) -- so the result equals `compute_data`'s for the same seed."""
    results_all = []
    k = args.factor_pred
    for _ in range(iteration_number[0]):
        pending = []
        for _ in range(iteration_number[1]):
            up, uref_traj, xref_traj = system.get_valid_ref(args)
            xe0 = system.sample_xe0(samples_x)
            t = (args.dt_sim * torch.arange(0, xref_traj.shape[2]))[::k]
            if samples_t == 0:
                indizes = args.N_sim - 1
            elif samples_t is not None:
                indizes = torch.randint(0, t.shape[0], (int(samples_t),))
                indizes = list(set(indizes.tolist()))
            else:
                indizes = torch.arange(0, t.shape[0])
            pending.append((up, uref_traj, xref_traj, xe0, t, indizes))
        # one launch per file; files whose stored trajectories would exceed ~8 GB are rolled out in several launches
        per_ref = 24 * int(samples_x) * max(int(p[2].shape[2]) for p in pending) if pending else 1
        group = max(1, int(8e9 // max(per_ref, 1)))
        outs = []
        for g0 in range(0, len(pending), group):
            part = pending[g0:g0 + group]
            outs += system.compute_density_batch([p[3] for p in part], [p[2] for p in part], [p[1] for p in part],
                                                 args.dt_sim, cutting=True, log_density=True, compute_density=True)
        for (up, uref_traj, xref_traj, xe0, t, indizes), (xe_traj, rho_traj) in zip(pending, outs):
            xe_s, rho_s = xe_traj[:, :, ::k], rho_traj[:, :, ::k]
            results_all.append({
                'u_params': up,
                'xe0': xe_s[:, :, 0],
                't': t[indizes],
                'xref0': xref_traj[0, :, 0],
                'xe_traj': xe_s[:, :, indizes],
                'rholog_traj': rho_s[:, :, indizes],
            })
        if save:
            data_name = args.path_rawdata + datetime.now().strftime("%Y-%m-%d-%H-%M-%S") + "_rawData_" + \
                system.systemname + "_%s_Nsim%d_iter%d_xSamples%d_tSamples%d" % (
                    args.input_type, args.N_sim, iteration_number[1], samples_x, samples_t) + ".pickle"
            print("save " + data_name)
            with open(data_name, "wb") as f:
                pickle.dump([results_all, args], f)
            results_all = []
    return results_all
