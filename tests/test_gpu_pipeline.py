"""GPU parity tests of the device-resident pipeline (dp_pipeline_* / dp_rollout_fused, density_planner_b200/pipeline.py):
samples -> rollout -> density weights -> occupancy grids + collision cost, against the separately pinned kernels, the
oracle, and its own sharding invariance (the multi-GPU contract: integer results identical for any split of the samples)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import golden, COST_RTOL, ROOT

pytestmark = pytest.mark.gpu

T, STRIDE = 100, 10


def _setup(dp, args, car, N, seed=0, env_seed=2):
    import dp_oracle
    e = golden("env_seed%d.npz" % env_seed)
    gx, gy = dp_oracle.env_gradient(e["grid"])
    denv = dp.DeviceEnv(torch.from_numpy(e["grid"]), torch.from_numpy(gx), torch.from_numpy(gy), args)
    torch.manual_seed(seed)
    # a reference that drives through the obstacles of this environment: start below the street, heading up
    while True:
        up, uref, xref = car.get_valid_ref(args)
        if uref.shape[2] >= T:
            break
    uref, xref = uref[:, :, :T].contiguous(), xref[:, :, :T + 1].contiguous()
    xe0 = car.sample_xe0(N)
    rho0 = torch.rand(N, 1, 1) + 0.5
    return e, (gx, gy), denv, xe0, rho0, xref, uref


def test_pipeline_matches_separate_kernels_and_oracle(dp, oracle, args, car):
    from density_planner_b200.pipeline import OccupancyPipeline
    N = 60_003
    e, (egx, egy), denv, xe0, rho0, xref, uref = _setup(dp, args, car, N)
    pipe = OccupancyPipeline(car, denv, args, T, STRIDE, N, rho0_bound=2.0)
    res = pipe.run(xe0.cuda(), rho0.cuda(), xref, uref, args.dt_sim)
    Ts = T // STRIDE + 1
    assert res.mass.shape == (Ts, 242, 402) and res.count.shape == (Ts, 242, 402)
    # (1) the stored slices are the rollout kernel's: log-density bit-equal to compute_density's, positions = xe + xref
    x, y, l = pipe.slices(N)
    xe, lt = car.compute_density(xe0.cuda(), xref, uref, args.dt_sim, cutting=False, log_density=True, stride=STRIDE)
    assert bool((l == lt[:, 0, :].t()).all())
    xa = xe[:, :2, :] + xref[:, :2, ::STRIDE].cuda()
    assert float((x.t() - xa[:, 0, :]).abs().max()) < 2e-6 and float((y.t() - xa[:, 1, :]).abs().max()) < 2e-6
    assert bool((pipe.lmax == l.max(dim=1).values).all())
    # (2) counts: bit-exact against the oracle's cells of the same positions
    xn, yn, ln = x.cpu().numpy(), y.cpu().numpy(), l.cpu().numpy()
    r0 = rho0.reshape(-1).numpy()
    k = res.mass_scale_log2
    for t in range(Ts):
        gx, gy = oracle.pos2gridpos(xn[t], yn[t])
        gx, gy = np.clip(gx, 0, 241), np.clip(gy, 0, 401)
        cnt = np.zeros((242, 402), np.int32)
        np.add.at(cnt, (gx, gy), 1)
        assert (res.count[t].cpu().numpy() == cnt).all()
        # mass: float64 accumulation of the same weights (numpy's expf may differ from CUDA's in the last bit)
        w = np.exp((ln[t] - ln[t].max()).astype(np.float32)).astype(np.float32) * r0
        m = np.zeros((242, 402))
        np.add.at(m, (gx, gy), w.astype(np.float64))
        got = res.mass[t].cpu().numpy() / 2.0 ** k
        assert np.abs(got - m).max() <= 1e-6 * m.max() + 2.0 ** (-k) * cnt.max()
        assert abs(float(res.weight_sum[t]) - w.astype(np.float64).sum()) <= 1e-6 * w.sum()
        assert int(res.mass[t].sum()) * 2.0 ** (-k) == float(res.weight_sum[t])
    # (3) collision cost and occupancy dot product of the NORMALISED density against the pinned cost kernel / oracle
    rho = dp.normalize_density(lt, rho0.cuda())                       # (N,1,Ts), the reference's normaliser
    x5 = torch.zeros(N, 5, Ts, device="cuda")
    x5[:, 0, :], x5[:, 1, :] = x.t(), y.t()
    c_ref = dp.get_cost_coll(x5, rho, denv)
    assert abs(float(res.cost) - float(c_ref)) <= COST_RTOL * abs(float(c_ref))
    oc = oracle.cost_coll(xn.T.copy(), yn.T.copy(), rho[:, 0, :].cpu().numpy(), e["grid"], egx, egy)
    oc = oc[0] if isinstance(oc, tuple) else oc
    assert abs(float(res.cost) - oc) <= COST_RTOL * abs(oc)
    assert float(c_ref) > 0, "the test reference must cross obstacles"
    grid = torch.from_numpy(e["grid"]).cuda().double()
    dot = (res.mass[:, :241, :401].double() * grid[:, :, :Ts].permute(2, 0, 1)).sum(dim=(1, 2)) / res.mass.double().sum(dim=(1, 2))
    assert float((res.occ_dot - dot).abs().max()) <= 1e-12 + 1e-10 * float(dot.abs().max())
    # pred2grid's convention from the same grids equals the standalone pred2grid of the slice
    t = Ts - 1
    xs = torch.zeros(N, 5, 1)
    xs[:, 0, 0], xs[:, 1, 0] = x[t].cpu(), y[t].cpu()
    g_ref = dp.pred2grid(xs, rho[:, :, t:t + 1].cpu(), args)[:, :, 0].double()
    assert float((res.pred2grid(t).cpu() - g_ref).abs().max()) <= 2e-6 * float(g_ref.max())
    pipe.close()


def test_pipeline_sharding_invariance_and_host_route(dp, args, car):
    """The multi-GPU contract on one GPU: shards binned against the GLOBAL max sum to the unsharded integer buffer bit for
    bit (mass, the int32 counts packed two per word, the fixed-point cost); the host route returns the same buffer."""
    import ctypes
    from density_planner_b200 import _lib
    from density_planner_b200.pipeline import OccupancyPipeline
    N = 150_001
    _, _, denv, xe0, rho0, xref, uref = _setup(dp, args, car, N, seed=3)
    L = dp.lib()
    pipe = OccupancyPipeline(car, denv, args, T, STRIDE, N, rho0_bound=2.0)
    full = pipe.run(xe0.cuda(), rho0.cuda(), xref, uref, args.dt_sim)
    full_buf = full.ibuf.clone()
    gmax = pipe.lmax.clone()
    host = pipe.run(xe0.pin_memory(), rho0.pin_memory(), xref, uref, args.dt_sim, to_host=True)
    assert bool((host.ibuf == full_buf.cpu()).all())
    assert bool((host.cost_slices == full.cost_slices.cpu()).all())
    dev = pipe.device
    sp = _lib.current_stream(dev)
    total = torch.zeros_like(full_buf)
    cuts = [0, 40_000, 40_001, 97_531, N]
    lmaxes = []
    xr, ur = xref.reshape(5, T + 1).cuda().contiguous(), uref.reshape(2, T).cuda().contiguous()
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        xs = xe0[lo:hi].reshape(-1, 5).cuda().contiguous()
        rs = rho0[lo:hi].reshape(-1).cuda().contiguous()
        lm = torch.empty(pipe.Ts, dtype=torch.float32, device=dev)
        _lib.check(L.dp_pipeline_rollout(pipe.handle, _lib.ptr(xs), _lib.ptr(rs), _lib.ptr(xr), _lib.ptr(ur), 0,
                                         float(args.dt_sim), hi - lo, _lib.ptr(lm), sp), "dp_pipeline_rollout")
        lmaxes.append(lm)
        buf = torch.empty(pipe.words, dtype=torch.int64, device=dev)
        _lib.check(L.dp_pipeline_bin(pipe.handle, _lib.ptr(gmax), _lib.ptr(buf), sp), "dp_pipeline_bin")
        total += buf                      # what the SUM allreduce does
    assert bool((torch.stack(lmaxes).max(dim=0).values == gmax).all())   # what the MAX allreduce does
    assert bool((total == full_buf).all())
    out3 = torch.empty((3, pipe.Ts), dtype=torch.float64, device=dev)
    _lib.check(L.dp_pipeline_finalize(pipe.handle, _lib.ptr(total), _lib.ptr(out3), sp), "dp_pipeline_finalize")
    assert bool((out3[1] == full.cost_slices).all()) and bool((out3[0] == full.weight_sum).all())
    # empty shard and capacity check
    empty = pipe.run(xe0[:0].cuda(), rho0[:0].cuda(), xref, uref, args.dt_sim)
    assert int(empty.ibuf.abs().sum()) == 0 and bool(torch.isinf(pipe.lmax).all())
    with pytest.raises(dp.DensityPlannerError):
        pipe.run(torch.zeros(N + 1, 5, 1).cuda(), None, xref, uref, args.dt_sim)
    pipe.close()


def test_pipeline_binning_knobs_agree(dp):
    """The binning kernel's A/B knobs (block size, warp aggregation) produce the same integer buffer: one fresh process per
    setting (the knobs are read once), each prints a fingerprint of the buffer."""
    here = os.path.dirname(os.path.abspath(__file__))
    prints = []
    for env in ({}, {"DP_BIN_AGG": "1"}, {"DP_BIN_AGG": "0", "DP_BIN_THREADS": "512"}, {"DP_BIN_AGG": "1", "DP_BIN_THREADS": "512"}):
        r = subprocess.run([sys.executable, os.path.join(here, "run_pipeline_fingerprint.py")], env=dict(os.environ, **env),
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        prints.append(r.stdout.strip().splitlines()[-1])
    assert len(set(prints)) == 1, prints


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (run with gpurun --gpus 2)")
def test_pipeline_two_ranks_nccl_equals_one_rank():
    """Hardware multi-GPU parity: two ranks over NCCL, each with half of the samples, return the integer buffer a single
    rank computes for all of them, bit for bit."""
    here = os.path.dirname(os.path.abspath(__file__))
    port = 29700 + (os.getpid() % 1000)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), os.path.join(here, "run_pipeline_ranks.py")],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    assert "PIPELINE_RANKS_OK" in r.stdout
