#!/usr/bin/env python
"""Bring-up probe for the tcgen05 path: runs each stage in its own process under a timeout so that a hang in a new
kernel cannot wedge the GPU box.  Usage: python scripts/tc_probe.py [stage ...]"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def split16(x):
    import numpy as np
    hi = x.astype(np.float16).astype(np.float32)
    lo = (x - hi).astype(np.float16).astype(np.float32)
    return hi, lo


def stage_selftest():
    import numpy as np
    import density_planner_b200 as dp
    from density_planner_b200 import _lib as _dl
    L = _dl.diag_lib()
    rs = np.random.RandomState(0)
    ok = True
    for n, n_valid in ((128, 128), (48, 48), (32, 24)):
        for mode in (0, 1):
            A = rs.uniform(-1, 1, (128, 128)).astype(np.float32)
            W = rs.uniform(-0.25, 0.25, (n_valid, 128)).astype(np.float32)
            D = np.zeros((128, n), np.float32)
            rc = L.dp_tc_selftest(0, A.ctypes.data_as(ctypes.c_void_p), W.ctypes.data_as(ctypes.c_void_p), n_valid, n, mode,
                                  16.0, D.ctypes.data_as(ctypes.c_void_p))
            assert rc == 0, L.dp_last_error()
            Ah, Al = split16(A)
            Wp = np.zeros((n, 128), np.float32)
            Wp[:n_valid] = W * 16
            Wh, Wl = split16(Wp)
            ref = Ah.astype(np.float64) @ (Wh + Wl).astype(np.float64).T
            if mode == 0:
                ref += Al.astype(np.float64) @ Wh.astype(np.float64).T
            exact = A.astype(np.float64) @ Wp.astype(np.float64).T
            err = np.abs(D - ref).max()
            print("selftest n=%d mode=%d: max|D-ref| %.3e   max|D-exact| %.3e   max|ref| %.2f" % (
                n, mode, err, np.abs(D - exact).max(), np.abs(ref).max()))
            if err > 1e-3:
                ok = False
                # diagnostics: which structured alternative does D match?
                alts = {"transposed": ref.T if n == 128 else None,
                        "hi_only": Ah.astype(np.float64) @ Wh.astype(np.float64).T,
                        "zeros": np.zeros_like(ref)}
                for k, v in alts.items():
                    if v is not None:
                        print("   vs %s: %.3e" % (k, np.abs(D - v).max()))
                print("   D[0,:8]  ", D[0, :8])
                print("   ref[0,:8]", ref[0, :8])
                print("   D[:8,0]  ", D[:8, 0])
                print("   ref[:8,0]", ref[:8, 0])
    print("SELFTEST", "OK" if ok else "FAILED")
    return ok


def stage_rollout(N, T, name):
    import numpy as np
    import torch
    import density_planner_b200 as dp
    import dp_oracle as O
    args = dp.default_args()
    car = dp.Car(args)
    torch.manual_seed(3)
    up, uref, xref = car.get_valid_ref(args)
    uref, xref = uref[:, :, :T], xref[:, :, :T + 1]
    xe0 = car.sample_xe0(N)
    xe, rho, margin = car.compute_density(xe0.cuda(), xref, uref, args.dt_sim, log_density=True, impl="tc",
                                          return_margin=True)
    torch.cuda.synchronize()
    o_xe, o_rho, o_m, _ = O.rollout(O.load_weights(), xe0[:, :, 0].numpy(), xref[0].numpy(), uref[0].numpy(), args.dt_sim,
                                    want_margin=True)
    xe, rho, margin = xe.cpu().numpy(), rho[:, 0, :].cpu().numpy(), margin.cpu().numpy()
    safe = np.minimum(margin, o_m) > 2e-5
    print("%s N=%d T=%d: states max err %.3e | log-density max err %.3e (safe %.3e, rel %.2e) | margin err %.2e | unsafe %d" % (
        name, N, T, np.abs(xe - o_xe).max(), np.abs(rho - o_rho).max(), np.abs(rho - o_rho)[safe].max(),
        (np.abs(rho - o_rho)[safe] / (np.abs(o_rho[safe]) + 1e-1)).max(), np.abs(margin - o_m).max(), int((~safe).sum())))
    if np.abs(xe - o_xe).max() > 1e-3:
        print("  xe[0,:,1]", xe[0, :, 1], "oracle", o_xe[0, :, 1])
        print("  rho[0,:4]", rho[0, :4], "oracle", o_rho[0, :4])


STAGES = {
    "selftest": (stage_selftest, (), 90),
    "roll_tiny": (stage_rollout, (128, 2, "tiny"), 90),
    "roll_small": (stage_rollout, (1000, 50, "small"), 120),
    "roll_mid": (stage_rollout, (40000, 100, "mid"), 180),
}

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--run":
        fn, a, _ = STAGES[sys.argv[2]]
        fn(*a)
        sys.exit(0)
    names = sys.argv[1:] or list(STAGES)
    for nm in names:
        _, _, to = STAGES[nm]
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--run", nm], timeout=to, capture_output=True, text=True)
            print(r.stdout, end="")
            if r.returncode != 0:
                print("[%s] exit code %d\n%s" % (nm, r.returncode, r.stderr[-2000:]))
                break
        except subprocess.TimeoutExpired as e:
            print("[%s] TIMEOUT after %d s (kernel hang?)\n%s" % (nm, to, (e.stdout or b"")[-1000:]))
            break
