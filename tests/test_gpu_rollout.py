"""GPU parity tests of the fused rollout kernel(s) and the single-step API, through the package (which calls the C ABI).
Checked against (1) the golden vectors of the unmodified reference, (2) the CPU oracle on fresh seeded inputs,
(3) size-independent properties at the full BASELINE size."""
import numpy as np
import pytest
import torch

from conftest import golden, assert_states_close, assert_rho_close, FLIP_MARGIN, RHO_RTOL, RHO_ATOL

pytestmark = pytest.mark.gpu

IMPLS = ["ffma", "tc"]


def _split16(x):
    hi = x.astype(np.float16).astype(np.float32)
    return hi, (x - hi).astype(np.float16).astype(np.float32)


@pytest.mark.parametrize("n,n_valid", [(128, 128), (48, 48), (32, 24)])
@pytest.mark.parametrize("mode", [0, 1])
def test_tc_mma_building_block(dp, n, n_valid, mode):
    """One 128 x n x 128 product through the TMEM-operand tcgen05.mma path against a numpy emulation of the split
    fp16 operands (value path: 3 products, fp32-level; tangent path: fp16 activations x hi+lo weights)."""
    import ctypes
    from density_planner_b200 import _lib
    L = _lib.diag_lib()
    rs = np.random.RandomState(n + mode)
    A = rs.uniform(-1, 1, (128, 128)).astype(np.float32)
    W = rs.uniform(-0.25, 0.25, (n_valid, 128)).astype(np.float32)
    D = np.zeros((128, n), np.float32)
    rc = L.dp_tc_selftest(0, A.ctypes.data_as(ctypes.c_void_p), W.ctypes.data_as(ctypes.c_void_p), n_valid, n, mode, 16.0,
                          D.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    Ah, Al = _split16(A)
    Wp = np.zeros((n, 128), np.float32)
    Wp[:n_valid] = W * 16
    Wh, Wl = _split16(Wp)
    ref = Ah.astype(np.float64) @ (Wh + Wl).astype(np.float64).T
    if mode == 0:
        ref += Al.astype(np.float64) @ Wh.astype(np.float64).T
    assert np.abs(D - ref).max() < 2e-6 * np.abs(ref).max() + 1e-5
    if mode == 0:  # three products reproduce the fp32 product to ~22 bits
        exact = A.astype(np.float64) @ Wp.astype(np.float64).T
        assert np.abs(D - exact).max() < 2e-6 * np.abs(exact).max() + 1e-5
    assert (D[:, n_valid:] == 0).all()


def _run(car, g, impl, stride=1, device="cuda", **kw):
    dev = torch.device(device)
    xe0 = torch.from_numpy(g["xe0"]).unsqueeze(-1).to(dev)
    xref = torch.from_numpy(g["xref"]).unsqueeze(0).to(dev)
    uref = torch.from_numpy(g["uref"]).unsqueeze(0).to(dev)
    out = car.compute_density(xe0, xref, uref, float(g["dt"]), stride=stride, impl=impl, return_margin=True, **kw)
    xe, rho, margin = out
    return (xe.cpu().numpy(), None if rho is None else rho[:, 0, :].cpu().numpy(), margin.cpu().numpy())


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("name", ["rollout_cfg1", "rollout_n2000_t300", "rollout_n4000_t100"])
def test_rollout_vs_reference_golden(car, impl, name):
    g = golden(name + ".npz")
    stride = int(g["stride"]) if "stride" in g.files else 1
    xe, rho, margin = _run(car, g, impl, stride=stride, cutting=True, log_density=True)
    assert xe.shape == g["xe_traj"].shape and rho.shape == g["rholog_traj"].shape
    assert_states_close(xe, g["xe_traj"], g["xref"][:, ::stride], name)
    assert_rho_close(rho, g["rholog_traj"], np.minimum(margin, g["clip_margin"]), name)
    assert np.abs(margin - g["clip_margin"]).max() < 1e-3


@pytest.mark.parametrize("env", [{"DP_TC_BOUNDED": "0"}, {"DP_TC_VARIANT": "120"}, {"DP_TC_VARIANT": "20"},
                                 {"DP_TC_V": "0"}, {"DP_TC_V": "1"}, {"DP_TC_V": "2"}, {"DP_TC_V": "3"},
                                 {"DP_TC_V": "3", "DP_TC_BOUNDED": "0"}])
def test_rollout_kernel_variants_vs_reference_golden(env):
    """The variants the host does not pick by default -- the general tanh (what a controller whose layer-2 weights do not
    bound the pre-activations gets), two tangent products with and without K-chunked MMA issue, and every production variant
    V (layer 1 handing r = (1 + h1) / 2 to layer 2 with the per-step bound check -- with the four-argument reciprocal, V = 1,
    these states exceed the bound regularly, so both the sign-free and the general branch run -- and the pair reciprocal) --
    against the same golden vectors; one fresh process each because the knobs are read once per process."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, os.path.join(here, "run_rollout_golden.py"), "rollout_n4000_t100", "tc"],
                       env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.parametrize("impl", IMPLS)
def test_rollout_host_buffers_and_strided_store(car, impl):
    """CPU tensors in / CPU tensors out (dp_rollout_host), and stride>1 equals subsampling of stride 1."""
    g = golden("rollout_n4000_t100.npz")
    xe_s, rho_s, _ = _run(car, g, impl, stride=10, device="cpu", cutting=True, log_density=True)
    xe_f, rho_f, _ = _run(car, g, impl, stride=1, device="cuda", cutting=True, log_density=True)
    assert (xe_s == xe_f[:, :, ::10]).all() and (rho_s == rho_f[:, ::10]).all()
    assert_states_close(xe_s, g["xe_traj"], g["xref"][:, ::10], "host")


def test_rollout_host_chunking_equals_one_device_launch(car, args):
    """dp_rollout_host processes the samples in chunks (one wave, then two waves each) and copies chunk k to the host while
    chunk k+1 runs; with more than three chunks (N > 5 waves) and a ragged tail the result must be bit-identical to one device
    launch over all samples, including the clip margins."""
    torch.manual_seed(11)
    T, N = 12, 5 * 2 * 128 * 148 + 4321
    up, uref, xref = car.get_valid_ref(args)
    uref, xref = uref[:, :, :T], xref[:, :, :T + 1]
    xe0 = car.sample_xe0(N)
    xh, rh, mh = car.compute_density(xe0, xref, uref, args.dt_sim, log_density=True, stride=4, return_margin=True)
    xd, rd, md = car.compute_density(xe0.cuda(), xref.cuda(), uref.cuda(), args.dt_sim, log_density=True, stride=4,
                                     return_margin=True)
    assert not xh.is_cuda and xd.is_cuda
    assert torch.equal(xh, xd.cpu()) and torch.equal(rh, rd.cpu()) and torch.equal(mh, md.cpu())


@pytest.mark.parametrize("impl", IMPLS)
def test_rollout_modes(car, impl):
    g = golden("rollout_modes.npz")
    s = int(g["stride"])
    xref_s = g["xref"][:, ::s]
    rho0 = torch.from_numpy(g["rho0"]).reshape(-1, 1, 1).cuda()
    xe, rho, _ = _run(car, g, impl, stride=s, rho0=rho0, cutting=True, log_density=False)
    assert_states_close(xe, g["xe_lin"], xref_s, "linear")
    # linear density rho = rho0 exp(l): a relative error of rho is an absolute error of the log-density l, whose contract is
    # rtol 1e-3 (+1e-4) on |l| (here |l| reaches 70 within 150 steps)
    l_ref = np.log(g["rho_lin"] / g["rho0"][:, None])
    rel = np.abs(rho / g["rho_lin"] - 1)
    tol = RHO_RTOL * np.abs(l_ref) + RHO_ATOL
    assert (rel <= tol).all(), "linear density: worst rel err / tol = %.3g" % (rel / tol).max()
    xe, rho, _ = _run(car, g, impl, stride=s, cutting=False, log_density=True)
    assert_rho_close(rho, g["rholog"], None, "log, no cutting")
    xe, rho, _ = _run(car, g, impl, stride=s, compute_density=False)
    assert rho is None
    assert_states_close(xe, g["xe_nodens"], xref_s, "no density")


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("N,T", [(1, 7), (31, 33), (33, 1), (1000, 50), (4737, 20)])
def test_rollout_vs_oracle_ragged(car, oracle, blob, args, impl, N, T):
    """Ragged sizes (tail tiles, single sample, single step) against the CPU oracle on seeded inputs."""
    torch.manual_seed(100 + N)
    up, uref, xref = car.get_valid_ref(args)
    uref, xref = uref[:, :, :T], xref[:, :, :T + 1]
    xe0 = car.sample_xe0(N)
    xe, rho, margin = car.compute_density(xe0.cuda(), xref, uref, args.dt_sim, log_density=True, impl=impl,
                                          return_margin=True)
    o_xe, o_rho, o_m, _ = oracle.rollout(blob, xe0[:, :, 0].numpy(), xref[0].numpy(), uref[0].numpy(), args.dt_sim,
                                         want_margin=True)
    assert_states_close(xe.cpu().numpy(), o_xe, xref[0].numpy(), "N=%d T=%d" % (N, T))
    assert_rho_close(rho[:, 0, :].cpu().numpy(), o_rho, np.minimum(margin.cpu().numpy(), o_m), "N=%d T=%d" % (N, T))


def test_rollout_empty(car, args):
    xe, rho = car.compute_density(torch.zeros(0, 5, 1).cuda(), torch.zeros(1, 5, 11), torch.zeros(1, 2, 10), 0.01,
                                  log_density=True)
    assert xe.shape == (0, 5, 11) and rho.shape == (0, 1, 11)


@pytest.mark.parametrize("impl", IMPLS)
def test_cutting_clamps_and_reports(car, args, impl, capsys):
    """Densities above 1e30 are clamped with the reference's message (ControlAffineSystem.py:451-462)."""
    torch.manual_seed(7)
    up, uref, xref = car.get_valid_ref(args)
    uref, xref = uref[:, :, :5], xref[:, :, :6]
    xe0 = car.sample_xe0(64).cuda()
    rho0 = torch.full((64, 1, 1), 3e30).cuda()
    xe, rho = car.compute_density(xe0, xref, uref, args.dt_sim, rho0=rho0, cutting=True, log_density=True, impl=impl)
    assert float(rho.max()) <= float(np.float32(1e30))
    assert "clamp rho_traj to 1e30 (log density)" in capsys.readouterr().out
    xe, rho = car.compute_density(xe0, xref, uref, args.dt_sim, rho0=rho0, cutting=False, log_density=True, impl=impl)
    assert float(rho.max()) > 2e30


@pytest.mark.parametrize("impl", IMPLS)
def test_full_size_properties(car, oracle, blob, args, impl):
    """BASELINE config 2 size (1M samples x 100 steps): (a) permutation equivariance / tile independence: the same
    sample gives the same trajectory wherever it sits in the batch; (b) a strided spot check against the oracle;
    (c) density of a duplicated sample is identical (determinism)."""
    N, T = 1_000_000, 100
    torch.manual_seed(11)
    up, uref, xref = car.get_valid_ref(args)
    uref, xref = uref[:, :, :T], xref[:, :, :T + 1]
    base = car.sample_xe0(4096)
    idx = torch.randint(0, 4096, (N,))
    xe0 = base[idx].cuda()
    xe, rho = car.compute_density(xe0, xref, uref, args.dt_sim, log_density=True, stride=10, impl=impl)
    # (a)+(c): all copies of base sample k are bit-identical
    xe_last = xe[:, :, -1]
    rho_last = rho[:, 0, -1]
    first = torch.full((4096,), -1, dtype=torch.long)
    order = torch.arange(N - 1, -1, -1)
    first[idx[order]] = order
    ref_x = xe_last[first.cuda()][idx.cuda()]
    ref_r = rho_last[first.cuda()][idx.cuda()]
    assert bool((ref_x == xe_last).all()) and bool((ref_r == rho_last).all())
    # (b) oracle spot check on the 4096 base samples
    o_xe, o_rho, o_m, _ = oracle.rollout(blob, base[:, :, 0].numpy(), xref[0].numpy(), uref[0].numpy(), args.dt_sim,
                                         stride=10, want_margin=True)
    got_x = xe[first.cuda()].cpu().numpy()
    got_r = rho[first.cuda()][:, 0, :].cpu().numpy()
    assert_states_close(got_x, o_xe, xref[0].numpy()[:, ::10], "1M spot check")
    assert_rho_close(got_r, o_rho, o_m, "1M spot check")


def test_step_api_vs_reference_golden(car):
    g = golden("step_api.npz")
    x = torch.from_numpy(g["x"]).unsqueeze(-1)
    xref = torch.from_numpy(g["xref"]).reshape(1, 5, 1)
    uref = torch.from_numpy(g["uref"]).reshape(1, 2, 1)
    dt = float(g["dt"])
    u = car.u_func(x, xref, uref)
    assert u.shape == (4096, 2, 1) and u.device.type == "cpu"
    assert np.allclose(u[:, :, 0].numpy(), g["u"], rtol=1e-5, atol=2e-5)
    f = car.f_func(x.cuda(), xref, uref)
    assert f.is_cuda and np.allclose(f[:, :, 0].cpu().numpy(), g["f"], rtol=1e-5, atol=2e-5)
    safe = (np.abs(np.abs(g["u_raw"]) - 3) > 1e-4).all(axis=1)
    dudx = car.dudx_func(x, xref, uref).numpy()
    assert dudx.shape == (4096, 2, 5)
    assert np.abs(dudx[safe] - g["dudx"][safe]).max() < 1e-4 * np.abs(g["dudx"]).max()
    dfdx = car.dfdx_func(x, xref, uref).numpy()
    assert np.abs(dfdx[safe] - g["dfdx"][safe]).max() < 1e-4 * np.abs(g["dfdx"]).max()
    div = car.divf_func(x, xref, uref).numpy()
    assert np.abs(div[safe] - g["div"][safe]).max() < 1e-4 * np.abs(g["div"]).max()
    xn = car.get_next_x(x, xref, uref, dt)
    assert np.allclose(xn[:, :, 0].numpy(), g["x_next"], rtol=1e-6, atol=2e-6)
    rho = torch.from_numpy(g["rho"])
    rn = car.get_next_rho(x, xref, uref, rho, dt).numpy()
    rln = car.get_next_rholog(x, xref, uref, rho, dt).numpy()
    assert np.allclose(rn[safe], g["rho_next"][safe], rtol=1e-4, atol=1e-5)
    assert np.allclose(rln[safe], g["rholog_next"][safe], rtol=1e-4, atol=1e-5)


def _check_rawdata_against_golden(car, args, res, g):
    """Elementwise contract tolerances (states rtol 1e-4, log-density rtol 1e-3 + clip-flip audit) for the raw-data entries.
    The clip margins come from replaying the seeded RNG order of compute_data (reference, initial errors, time-index draw:
    compute_rawdata.py:26-36) and rolling the same samples out once more with return_margin=True."""
    torch.manual_seed(5)
    k = args.factor_pred
    for i, r in enumerate(res):
        up, uref, xref = car.get_valid_ref(args)
        xe0 = car.sample_xe0(20)
        torch.randint(0, xref[:, :, ::k].shape[2], (10,))          # the time-index draw consumes the RNG
        # (the stored 'xe0' is the rollout's slice 0, i.e. (xe0 + xref0) - xref0: equal to the drawn xe0 up to one rounding)
        assert np.abs(xe0[:, :, 0].numpy() - g["r%d_xe0" % i]).max() < 1e-5
        xe, rho, margin = car.compute_density(xe0.cuda(), xref, uref, args.dt_sim, log_density=True, return_margin=True)
        idx = np.rint(r["t"].numpy() / (args.dt_sim * k)).astype(int)
        assert bool((xe.cpu()[:, :, ::k][:, :, idx] == r["xe_traj"]).all())    # compute_data returned this very rollout
        assert_states_close(r["xe_traj"].numpy(), g["r%d_xe_traj" % i], xref[0].numpy()[:, ::k][:, idx], "compute_data[%d]" % i)
        assert_rho_close(r["rholog_traj"][:, 0, :].numpy(), g["r%d_rholog_traj" % i][:, 0, :], margin.cpu().numpy(),
                         "compute_data[%d]" % i)


def test_compute_data_matches_reference_golden(dp, car, args):
    """data_generation/compute_rawdata.py:9 with seed 5: same references, same samples, same time indices.  `compute_data`
    rolls a file out in one packed launch; a system that overrides get_valid_trajectories gets the reference's
    trajectory-by-trajectory loop (one launch each) -- both reproduce the golden."""
    g = golden("compute_data.npz")

    class LoopCar(type(car)):  # forces the per-trajectory loop of compute_data
        def get_valid_trajectories(self, *a, **k):
            return super().get_valid_trajectories(*a, **k)

    loop_car = LoopCar(args)
    torch.manual_seed(5)
    res_loop = dp.compute_data([1, 2], 20, loop_car, args, samples_t=10, save=False, plot=False)
    torch.manual_seed(5)
    res = dp.compute_data([1, 2], 20, car, args, samples_t=10, save=False, plot=False)
    assert len(res_loop) == len(res) == 2
    for a, b in zip(res_loop, res):
        for k in a:
            assert bool((torch.as_tensor(a[k]) == torch.as_tensor(b[k])).all()), k   # packed launch == single launches, bit for bit
    assert len(res) == 2
    for i, r in enumerate(res):
        assert (r["u_params"].numpy() == g["r%d_u_params" % i]).all()
        assert (r["xref0"].numpy() == g["r%d_xref0" % i]).all()
        assert (r["t"].numpy() == g["r%d_t" % i]).all()
        assert (r["xe0"].numpy() == g["r%d_xe0" % i]).all()
    _check_rawdata_against_golden(car, args, res, g)


def test_batched_references_equal_single_launches(dp, car, args):
    """dp_rollout_batch: R references with their own horizons in one launch == R separate compute_density calls, bit for
    bit (every tile belongs to one reference), for sample counts below, at and above one tile."""
    torch.manual_seed(21)
    refs = []
    for r in range(5):
        up, uref, xref = car.get_valid_ref(args)
        T = 37 + 11 * r
        refs.append((uref[:, :, :T], xref[:, :, :T + 1]))
    for n in (20, 128, 150):
        xe0s = [car.sample_xe0(n) for _ in refs]
        outs = car.compute_density_batch(xe0s, [x for _, x in refs], [u for u, _ in refs], args.dt_sim, log_density=True)
        for (uref, xref), xe0, (xe_b, rho_b) in zip(refs, xe0s, outs):
            xe_s, rho_s = car.compute_density(xe0.cuda(), xref, uref, args.dt_sim, log_density=True, impl="tc")
            assert xe_b.shape == xe_s.shape and rho_b.shape == rho_s.shape
            assert bool((xe_b == xe_s.cpu()).all()) and bool((rho_b == rho_s.cpu()).all())


def test_compute_data_batched_matches_reference_golden(dp, car, args):
    """The batched raw-data generator (one launch per file) reproduces the reference's seeded output like compute_data."""
    g = golden("compute_data.npz")
    torch.manual_seed(5)
    res = dp.compute_data_batched([1, 2], 20, car, args, samples_t=10, save=False)
    assert len(res) == 2
    for i, r in enumerate(res):
        assert (r["u_params"].numpy() == g["r%d_u_params" % i]).all()
        assert (r["xref0"].numpy() == g["r%d_xref0" % i]).all()
        assert (r["t"].numpy() == g["r%d_t" % i]).all()
        assert (r["xe0"].numpy() == g["r%d_xe0" % i]).all()
    _check_rawdata_against_golden(car, args, res, g)
