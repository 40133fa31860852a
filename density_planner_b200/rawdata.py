"""Raw-data generator entry: mirror of data_generation/compute_rawdata.py:9-63 (`compute_data`).

The loop, the torch RNG order (references, initial errors, then the time-index draw at :34) and the result
dictionary schema are the reference's; the rollout inside `get_valid_trajectories` is the fused CUDA kernel.
"""
import pickle
from datetime import datetime

import torch


def compute_data(iteration_number, samples_x, system, args, samples_t=None, save=True, plot=True):
    """Generate `iteration_number[0]` files of `iteration_number[1]` trajectories with `samples_x` samples each."""
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
