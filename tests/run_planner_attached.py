#!/usr/bin/env python
"""BASELINE config 4 on the kernels: the reference's own MotionPlannerGrad (unmodified, from the git-ignored copy under
baseline/_ref/) driven through dp.attach on one artificial environment -- initialisation phase (100 trajectories x 100
epochs on get_cost_coll_initialize fwd + bwd, up2ref_traj fwd + adjoint), density phase (100 epochs on the density NN,
goal / bounds / collision cost kernels) and validate_ref (LE rollout of 500 samples x 1000 steps + normaliser + costs).

    python tests/run_planner_attached.py SEED OUT.npz [--plain]      (--plain: the same WITHOUT attach, i.e. the reference on the CPU)

Test infrastructure (it imports oracle/ref_harness.py for the import stubs); not part of the product."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness as rh  # noqa: E402


def main():
    seed, out_path = int(sys.argv[1]), sys.argv[2]
    plain = "--plain" in sys.argv
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(os.path.join(ref_root, "systems")):
        rh.REF_ROOT = ref_root
    assert rh.available(), "no reference tree (run baseline/make_ref.sh in the build container)"
    rh.install()
    args = rh.default_args()
    log = "/tmp/dp_log/"
    os.makedirs(log, exist_ok=True)
    with rh.ref_cwd():
        from motion_planning.example_objects import create_mp_task
        from motion_planning.MotionPlannerGrad import MotionPlannerGrad
        torch.manual_seed(seed)
        np.random.seed(seed)
        ego = create_mp_task(args, seed)
        planner = MotionPlannerGrad(ego, path_log=log)
        if not plain:
            import density_planner_b200 as dp
            dp.attach(planner)
        times = {}
        t0 = time.time()
        up_best, cost_min = planner.find_initial_traj()
        times["init"] = time.time() - t0
        init_costs = planner.initial_traj[1]
        t0 = time.time()
        up, cost = planner.optimize_traj(up_best)
        times["opt"] = time.time() - t0
        opt_costs = planner.improved_traj[-1][1]
        t0 = time.time()
        val = planner.validate_ref(up)
        times["validate"] = time.time() - t0
    cpu = lambda v: v.detach().cpu() if torch.is_tensor(v) else torch.as_tensor(v)
    out = {"up_init_all": cpu(planner.initial_traj[0]), "up_best": cpu(up_best), "cost_min": cpu(cost_min), "up_final": cpu(up)}
    for k, v in init_costs.items():
        out["init_" + k] = torch.stack([cpu(c) for c in v])
    for k, v in opt_costs.items():
        out["opt_" + k] = torch.stack([cpu(c).reshape(()) for c in v])
    for k, v in cost.items():
        out["final_" + k] = cpu(v)
    for k, v in val.items():
        out["val_" + k] = cpu(v)
    for k, v in times.items():
        out["time_" + k] = torch.tensor(v)
    np.savez_compressed(out_path, **{k: v.numpy() for k, v in out.items()})
    print("PLANNER_DONE seed %d %s: init %.1f s, opt %.1f s, validate %.1f s; final cost_sum %.6f, validated %.6f" % (
        seed, "plain" if plain else "attached", times["init"], times["opt"], times["validate"], float(cost["cost_sum"]),
        float(val["cost_sum"])))


if __name__ == "__main__":
    main()
