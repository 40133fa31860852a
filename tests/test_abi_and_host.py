"""CPU suite: the C-ABI library loads and exports every symbol the header declares; host-side logic of the package
(weight packing, reference sampling in the reference's RNG order, loud failure without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, golden


def header_functions(name="density_planner_b200.h"):
    src = open(os.path.join(ROOT, "include", name)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(dp):
    from density_planner_b200 import _lib
    names = header_functions()
    assert len(names) >= 15
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), "missing export %s" % n
    # the binding table covers the header too
    assert set(names) == set(_lib.EXPORTS)
    assert dp.lib().dp_abi_version() == 1


def test_diagnostics_library_is_separate(dp):
    """The fp32 cross-check rollout and the micro-probes live in their own library and header, not in the product ABI."""
    from density_planner_b200 import _lib
    names = [n for n in header_functions("density_planner_b200_diag.h")]
    assert set(names) == set(_lib.DIAG_EXPORTS)
    D = ctypes.CDLL(_lib.DIAG_LIB_PATH)
    for n in names:
        assert hasattr(D, n), "missing diagnostics export %s" % n
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert not hasattr(L, n), "%s must not be exported by the product library" % n
    product = open(os.path.join(ROOT, "include", "density_planner_b200.h")).read()
    assert "DP_IMPL_FFMA" not in product and "dp_tc_selftest" not in product


def test_host_chunk_schedule(dp):
    """dp_rollout_host's chunk schedule (host logic, no device): sizes add up to N, no chunk exceeds two waves (the size of
    the device staging buffers), the first chunk is at most one wave (the copy engine starts early) and so is the last one
    (its copy is the only one nothing overlaps)."""
    L = dp.lib()
    wave = 2 * 128 * 148
    buf = (ctypes.c_int64 * 64)()
    for N in [1, 19, wave - 1, wave, wave + 1, 2 * wave, 3 * wave, 3 * wave + 1, 4 * wave + 5, 1_000_000, 26 * wave, 27 * wave - 3]:
        n = L.dp_host_chunk_schedule(N, wave, buf, 64)
        sizes = list(buf[:n])
        assert n >= 1 and sum(sizes) == N and all(0 < s <= 2 * wave for s in sizes), (N, sizes)
        assert sizes[0] <= wave and sizes[-1] <= wave, (N, sizes)
        assert n <= N // (2 * wave) + 3
    assert L.dp_host_chunk_schedule(0, wave, buf, 64) == 0
    assert L.dp_host_chunk_schedule(10 * wave, wave, buf, 2) < 0  # too small a buffer is reported, not overrun


def test_library_is_sm100a_native():
    """The shipped .so carries sm_100a SASS (no PTX-JIT path, no other arch)."""
    import subprocess
    from density_planner_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out)


def test_weight_packing_matches_oracle(oracle, blob):
    from density_planner_b200.controller import load_blob, random_blob
    b = load_blob()
    assert b.dtype == np.float32 and b.size == 43592
    assert (b == blob).all()
    r = random_blob(3)
    assert r.size == 43592 and np.isfinite(r).all()


def test_reference_sampling_rng_order(dp, args):
    """get_valid_ref + sample_xe0 consume the torch RNG exactly like the reference (golden minted with seed 0)."""
    g = golden("rollout_cfg1.npz")
    car = dp.Car(args)
    torch.manual_seed(0)
    up, uref, xref = car.get_valid_ref(args)
    xe0 = car.sample_xe0(20)
    assert (up.numpy() == g["up"]).all()
    assert (uref[0].numpy() == g["uref"]).all()
    assert (xref[0].numpy() == g["xref"]).all()
    assert (xe0[:, :, 0].numpy() == g["xe0"]).all()


def test_up2ref_and_cut(dp, args):
    car = dp.Car(args)
    up = torch.tensor([[[1.0, -1, 0.5, 0, 0, 2, -2, 0, 1, 0], [0.5, 0, 0, -0.5, 0, 0, 0.3, 0, 0, 0]]])
    xref0 = torch.tensor([0, -28, 1.5, 3, 0.0]).reshape(1, 5, 1)
    uref, xref = car.up2ref_traj(xref0, up, args, short=True)
    assert uref.shape == (1, 2, 100) and xref.shape == (1, 5, 101)
    ur_l, xr_l = car.up2ref_traj(xref0, up, args, short=False)
    assert ur_l.shape == (1, 2, 1000) and xr_l.shape == (1, 5, 1001)
    # Euler step of the bicycle model by hand
    x1 = xref0[0, :, 0] + torch.tensor([3 * np.cos(1.5), 3 * np.sin(1.5), 1.0, 0.5, 0.0]) * 0.01
    assert torch.allclose(xr_l[0, :, 1], x1.float(), atol=1e-6)
    # cut: leave the admissible box in v
    xr_bad = xr_l.clone()
    xr_bad[0, 3, 500:] = 11.0
    xc, uc = car.cut_xref_traj(xr_bad, ur_l)
    assert xc.shape[2] == 500 and uc.shape[2] == 499


def test_no_cpu_fallback(dp, args):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    car = dp.Car(args)
    with pytest.raises(dp.DensityPlannerError):
        car.compute_density(car.sample_xe0(4), torch.zeros(1, 5, 11), torch.zeros(1, 2, 10), 0.01)
    with pytest.raises(dp.DensityPlannerError):
        dp.pos2gridpos(args, pos_x=torch.zeros(3))


def test_product_does_not_import_oracle():
    """The product package must not reach into oracle/ (checked on the sources)."""
    pkg = os.path.join(ROOT, "density_planner_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                s = open(os.path.join(dirpath, f)).read()
                assert "dp_oracle" not in s and "ref_harness" not in s, f


def test_density_nn_batched_matches_reference_golden():
    """the density-NN branch of predict_density, all 101 time points in one batched pass, against the
    reference's 101-call loop (forward, normalised density and the gradient w.r.t. the input parameters).  Here the batched
    formulation (row layout, normaliser, autograd plumbing) in plain torch on the CPU; the product route -- the fused tcgen05
    kernel -- replays the same golden in tests/test_gpu_planner_terms.py."""
    import types
    import torch
    import density_planner_b200 as dp
    from conftest import golden
    g = golden("density_nn.npz")
    args = dp.default_args()
    nn = dp.DensityNN(backend="torch")   # the batched formulation in plain torch: host logic, runs without a GPU
    system = types.SimpleNamespace()
    pred = dp.EgoPredictor(system, args, torch.from_numpy(g["xref0"]).reshape(1, 5, 1), torch.from_numpy(g["xe0"]).unsqueeze(-1),
                           torch.from_numpy(g["rho0"]).reshape(-1, 1, 1), density_nn=nn)
    up = torch.from_numpy(g["up"]).requires_grad_(True)
    xref = torch.from_numpy(g["xref_short"])
    x_traj, rho_traj = pred.predict_density(up, xref, use_nn=True)
    assert x_traj.shape == g["x_traj"].shape and rho_traj.shape == (g["rho_traj"].shape[0], 1, g["rho_traj"].shape[1])
    assert np.abs(x_traj.detach().numpy() - g["x_traj"]).max() < 1e-4
    ref = g["rho_traj"]
    got = rho_traj.detach()[:, 0, :].numpy()
    assert np.abs(got - ref).max() < 1e-4 * ref.max()
    loss = (torch.from_numpy(g["w"]) * x_traj).sum() + 1e3 * (torch.from_numpy(g["wr"]).unsqueeze(1) * rho_traj).sum()
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) < 2e-4 * abs(float(g["loss"]))
    # x_traj = xe_nn + xref and the golden differentiates through the reference integrator as well (xref depends on up):
    # that part comes from the oracle's adjoint with the upstream gradient w summed over the samples
    import dp_oracle as O
    uref_long = np.repeat(g["up"], 100, axis=2)[:, :, :1000]
    xref_long = O.xref_traj(g["xref0"], uref_long, 0.01)
    gx = np.zeros((1, 5, 1001))
    gx[0, :, ::10] = g["w"].sum(axis=0)
    gu, _ = O.xref_traj_grad(xref_long, gx, 0.01)
    gup_int = gu.reshape(1, 2, 10, 100).sum(axis=3)
    total = up.grad.numpy() + gup_int
    assert np.abs(total - g["gup"]).max() < 2e-3 * np.abs(g["gup"]).max()


def test_density_nn_kernel_backend_fails_loudly_without_gpu():
    """The product route of the density-NN branch is the fused CUDA kernel: without a GPU it raises, it does not fall back."""
    import torch
    import density_planner_b200 as dp
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the kernel backend is exercised by tests/test_gpu_planner_terms.py")
    with pytest.raises(dp.DensityPlannerError):
        dp.DensityNN()
    with pytest.raises(dp.DensityPlannerError):
        dp.DensityNN(device="cuda")
