#!/usr/bin/env python
"""Benchmark of the density_planner hot path on B200 (contract: see the task statement / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this framework's arm
    python bench.py --impl reference [--gpus N] ...                 # CPU arm: the reference's own compute_density on the host
                                                                    # cores (baseline/_ref; the oracle port if it is absent)

One "step" = one pass of the hot path over one batch of synthetic input on every rank, through the device-resident
pipeline (density_planner_b200/pipeline.py, dp_pipeline_*):
    fused rollout of SAMPLES x STEPS closed-loop sample-steps with Liouville log-density (BASELINE config 2:
    1M samples x 100 steps); the (x, y) slices and the log-density of every 10th step stay on the device, the per-slice
    max of the log-density is reduced inside the rollout  [-> NCCL allreduce MAX of Ts floats when N > 1]
    ->  ONE pass over the stored slices: density weights, int64/int32 occupancy grids and the collision cost in int64 fixed
    point  [-> ONE NCCL allreduce SUM of that integer buffer when N > 1]  ->  per-slice normalisation.
`value` = closed-loop sample-steps/s over all ranks with inputs resident in HBM (CUDA events, max over ranks).
`e2e`   = the same metric through the public API `OccupancyPipeline.run` with HOST tensors: pinned xe0 / rho0 in (H2D),
          grids + per-slice costs out (D2H), both inside the timed region.  `e2e_trajectory` is the other public route,
          `Car.compute_density` with host tensors, which returns the full (N,5,T+1)+(N,1,T+1) trajectory (2.4 GB per step).
`config5` = BASELINE config 5: 100M samples x 100 steps in total (100M / N per rank), host samples in, allreduced grids +
          cost out; the integer results are bit-identical at 1/2/4/8 GPUs (`checks`).
Multi-GPU is weak scaling for `value` / `e2e` (every rank rolls out its own SAMPLES samples), strong for `config5`; there is
no collective in the rollout.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SAMPLES = 1_000_000
STEPS = 100
STRIDE = 10
FLOP_PER_SAMPLE_STEP = 248_288  # SURVEY section 8d / DESIGN.md: algorithmic FLOPs (124,144 MAC) per sample-step
# executed tensor-pipe work: value rows (43,008 MAC) run 3 split-fp16 products, tangent rows (80,896 MAC) one; the head sums
# (240 MAC) run on CUDA cores
EXECUTED_MMA_FLOP_PER_SAMPLE_STEP = 2 * (3 * 43_008 + 80_896)
CONFIG5_SAMPLES = 100_000_000
BLOCK = 1_000_000  # config 5 samples are generated in seeded blocks of 1M so that the global set is the same at every N
METRIC = "closed-loop sample-steps/sec with Liouville density"
UNIT = "sample-steps/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def synthetic_env(args, seed, device):
    """Synthetic occupancy environment of the reference's shape (241 x 401 x 101, time innermost): street walls plus
    random static / moving boxes with graded certainty, and centred-difference gradients (border value 1)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    G0, G1, Tg = args.grid_size[0], args.grid_size[1], 101
    grid = torch.zeros(G0, G1, Tg)
    grid[:35, :, :] = 1.0
    grid[-35:, :, :] = 1.0
    for _ in range(12):
        w, h = [int(v) for v in torch.randint(5, 40, (2,), generator=g)]
        x0 = int(torch.randint(45, G0 - 45 - w, (1,), generator=g))
        y0 = int(torch.randint(60, G1 - 60 - h, (1,), generator=g))
        cert = float(torch.randint(3, 10, (1,), generator=g)) / 10
        vx, vy = [int(v) for v in torch.randint(-2, 2, (2,), generator=g)]
        moving = bool(torch.rand(1, generator=g) < 0.6)
        for t in range(Tg):
            xs = x0 + (vx * t if moving else 0)
            ys = y0 + (vy * t if moving else 0)
            for ring in range(4):
                a0, a1 = max(xs - ring, 0), min(xs + w + ring, G0)
                b0, b1 = max(ys - ring, 0), min(ys + h + ring, G1)
                if a1 > a0 and b1 > b0:
                    grid[a0:a1, b0:b1, t] += cert / 4
    grid = grid.clamp(0, 1).to(device)
    pad = torch.nn.functional.pad
    gx = (pad(grid, (0, 0, 0, 0, 1, 0), value=1.0)[:-1] - pad(grid, (0, 0, 0, 0, 0, 1), value=1.0)[1:]) / 2
    gy = (pad(grid, (0, 0, 1, 0), value=1.0)[:, :-1] - pad(grid, (0, 0, 0, 1), value=1.0)[:, 1:]) / 2
    return grid, gx, gy


def cpu_oracle_rate(target_seconds=12.0):
    """Times the CPU restatement (oracle/dp_oracle.c, OpenMP over samples) on a bounded sample of the workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import dp_oracle as O
    import torch
    import density_planner_b200 as dp
    blob = O.load_weights()
    a = dp.default_args()
    car = dp.Car(a)
    torch.manual_seed(0)
    up, uref, xref = car.get_valid_ref(a)
    uref, xref = uref[0, :, :STEPS].numpy(), xref[0, :, :STEPS + 1].numpy()
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    cores = O.set_threads(cores)  # all host threads (torchrun exports OMP_NUM_THREADS=1 by default)
    n_probe = max(cores * 8, 64)
    xe0 = car.sample_xe0(n_probe)[:, :, 0].numpy()
    t0 = time.perf_counter()
    O.rollout(blob, xe0, xref, uref, a.dt_sim, stride=STRIDE)
    rate = n_probe * STEPS / (time.perf_counter() - t0)
    n, dt = n_probe, 0.0
    for _ in range(3):  # the first probe runs cold; re-size until the sample is a bounded ~target_seconds of CPU work
        n = int(min(max(rate * target_seconds / STEPS, n_probe), SAMPLES))
        xe0 = car.sample_xe0(n)[:, :, 0].numpy()
        t0 = time.perf_counter()
        O.rollout(blob, xe0, xref, uref, a.dt_sim, stride=STRIDE)
        dt = time.perf_counter() - t0
        rate = n * STEPS / dt
        if dt > 0.6 * target_seconds or n == SAMPLES:
            break
    return {"value": n * STEPS / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d samples x %d steps of the same rollout (C restatement of the reference's compute_density, "
                      "analytic divergence, OpenMP over samples), %.1f s" % (n, STEPS, dt)}, n, dt


def cpu_reference_rate(target_seconds=12.0):
    """Times the reference's OWN compute_density (systems/ControlAffineSystem.py:409-463: per step three forward passes and
    two autograd passes over the whole sample batch, PyTorch on the host cores) on a bounded sample of the workload.  The
    unmodified reference tree travels to the GPU box as the git-ignored copy baseline/_ref/ (baseline/make_ref.sh); returns
    None when it is absent."""
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_root, "systems")):
        return None
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import ref_harness as rh
    rh.REF_ROOT = ref_root
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    torch.set_num_threads(cores)  # torchrun exports OMP_NUM_THREADS=1 by default
    car, a = rh.make_car()
    torch.manual_seed(0)
    with rh.ref_cwd():
        while True:  # get_valid_trajectories' rejection sampling of the reference (:465-475)
            uref, _ = car.sample_uref_traj(a)
            xref = car.compute_xref_traj(car.sample_xref0(), uref, a)
            xref, uref = car.cut_xref_traj(xref, uref)
            if xref.shape[2] > 0.99 * a.N_sim:
                break
        n = 20000  # the reference's per-step cost is flat below ~1e4 samples (Python / autograd overhead), linear above
        xe0 = car.sample_xe0(n)
        t0 = time.perf_counter()
        car.compute_density(xe0, xref[:, :, :5], uref[:, :, :4], a.dt_sim, cutting=True, log_density=True)
        per_step = (time.perf_counter() - t0) / 4
        if per_step * STEPS < 0.7 * target_seconds:  # a fast host: more samples, so that the sample stays ~target_seconds
            n = int(min(n * target_seconds / (per_step * STEPS), 200000))
            xe0 = car.sample_xe0(n)
            per_step *= n / 20000.0
        steps = int(min(max(target_seconds / per_step, 4), STEPS))
        t0 = time.perf_counter()
        car.compute_density(xe0, xref[:, :, :steps + 1], uref[:, :, :steps], a.dt_sim, cutting=True, log_density=True)
        dt = time.perf_counter() - t0
    return {"value": n * steps / dt, "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": "%d samples x %d steps of the same rollout through the UNMODIFIED reference's Car.compute_density "
                      "(PyTorch CPU, autograd divergence, %d threads), %.1f s" % (n, steps, cores, dt)}, n * steps, dt


def run_reference_arm(opts):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_all = time.perf_counter()
    base = None
    times = []
    budget = 6.0 if opts.steps > 1 else 12.0
    try:  # the reference itself when its tree is on the box (and importable there), else the C port of its algorithm
        use_ref = cpu_reference_rate(target_seconds=2.0) is not None
    except Exception as exc:  # noqa: BLE001 -- a broken copy must not take the CPU arm down
        print("bench.py: the reference under baseline/_ref could not be run (%r); timing the C port instead" % (exc,), file=sys.stderr)
        use_ref = False
    for i in range(opts.warmup + opts.steps):
        if use_ref:
            base, units, dt = cpu_reference_rate(target_seconds=budget)
        else:
            base, n, dt = cpu_oracle_rate(target_seconds=budget)
            units = n * STEPS
        if i >= opts.warmup:
            times.append((units, dt))
        if time.perf_counter() - t_all > 240:
            break
    if not times:
        times.append((units, dt))
    tot_u = sum(u for u, _ in times)
    tot_t = sum(dt for _, dt in times)
    value = tot_u / tot_t
    base["value"] = value
    if use_ref:  # the C restatement beside it (analytic divergence, OpenMP): what the reference could do on the same cores
        base["port"] = cpu_oracle_rate(target_seconds=6.0)[0]
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": opts.gpus,
            "steps": len(times), "warmup": opts.warmup, "ms_per_step": 1e3 * tot_t / max(len(times), 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "CAR rollout + Liouville density, %d samples x %d steps per GPU (config 2); CPU arm "
                                   "runs a bounded sample of it on the host cores" % (SAMPLES, STEPS)},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def config5_samples(car, rank, world, total):
    """The rank's shard of the GLOBAL config-5 sample set: blocks of BLOCK samples, block b drawn with seed 5000 + b, so the
    union over ranks is the same set at every world size (which is what makes the integer results comparable across N)."""
    import torch
    nblocks = (total + BLOCK - 1) // BLOCK
    from density_planner_b200.dist import shard_range
    b_lo, b_hi = shard_range(nblocks, rank, world)
    n = sum(min(BLOCK, total - b * BLOCK) for b in range(b_lo, b_hi))
    xe0 = torch.empty((n, 5, 1), dtype=torch.float32, pin_memory=True)
    rho0 = torch.empty((n,), dtype=torch.float32, pin_memory=True)
    off = 0
    for b in range(b_lo, b_hi):
        k = min(BLOCK, total - b * BLOCK)
        torch.manual_seed(5000 + b)
        xe0[off:off + k] = car.sample_xe0(k)
        rho0[off:off + k] = torch.rand(k) + 0.5
        off += k
    return xe0, rho0


def grid_checks(res):
    """Integer fingerprints of a pipeline result (identical for any sharding of the same global sample set)."""
    import torch
    m = res.mass.reshape(res.mass.shape[0], -1)
    c = res.count.reshape(res.count.shape[0], -1).to(torch.int64)
    pos = (torch.arange(m.shape[1], device=m.device, dtype=torch.int64) % 1009) + 1
    return {"grid_mass_total": int(m.sum().item()), "grid_count_total": int(c.sum().item()),
            "grid_mass_checksum": int(((m % 1000003) * pos).sum().item()),
            "grid_count_checksum": int((c * pos).sum().item()),
            "cost_fixed_total": int(res.cost_fixed.sum().item()),
            "cost_total": float(res.cost_slices.sum().item()), "occupancy_dot_total": float(res.occ_dot.sum().item())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--samples", type=int, default=SAMPLES)
    ap.add_argument("--config5-samples", type=int, default=CONFIG5_SAMPLES, help="total samples of the config-5 block (0: skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-trajectory-e2e", action="store_true")
    opts = ap.parse_args()
    if opts.impl == "reference":
        return run_reference_arm(opts)

    # stdout carries exactly ONE JSON line: anything libraries print there (NCCL's version banner, ...) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import density_planner_b200 as dp
    from density_planner_b200 import _lib, dist as D, grid as G
    from density_planner_b200.pipeline import OccupancyPipeline
    import ctypes

    rank, world, local_rank = D.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200; there is no CPU fallback (use --impl reference for the CPU arm)")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    numa_node = D.bind_to_gpu_numa_node(local_rank) if (world > 1 and not os.environ.get("DP_NO_NUMA_BIND")) else None
    L = dp.lib()
    args = dp.default_args()
    car = dp.Car(args, device=dev)
    N, T, Ts = opts.samples, STEPS, STEPS // STRIDE + 1

    # ---- synthetic inputs (seeded; rank r uses seed 1000 + r for its samples) ---------------------------------------
    torch.manual_seed(0)
    up, uref, xref = car.get_valid_ref(args)
    uref, xref = uref[:, :, :T].contiguous(), xref[:, :, :T + 1].contiguous()
    torch.manual_seed(1000 + rank)
    xe0_host = car.sample_xe0(N).pin_memory()
    rho0_host = (torch.rand(N) + 0.5).pin_memory()   # unnormalised sample weights in [0.5, 1.5): independent of N and world
    xe0 = xe0_host.to(dev)
    rho0 = rho0_host.to(dev)
    xref_d, uref_d = xref.to(dev), uref.to(dev)
    grid, gx, gy = synthetic_env(args, 0, dev)
    denv = dp.DeviceEnv(grid, gx, gy, args, device=dev)
    del grid, gx, gy
    cap5 = 0
    if opts.config5_samples > 0:
        nb = (opts.config5_samples + BLOCK - 1) // BLOCK
        b_lo, b_hi = D.shard_range(nb, rank, world)
        cap5 = (b_hi - b_lo) * BLOCK
    pipe = OccupancyPipeline(car, denv, args, T, STRIDE, max(N, cap5, 1), rho0_bound=2.0)
    stream = torch.cuda.current_stream(dev)
    sp = ctypes.c_void_p(stream.cuda_stream)
    ev_k0 = [torch.cuda.Event(enable_timing=True) for _ in range(opts.steps)]
    ev_k1 = [torch.cuda.Event(enable_timing=True) for _ in range(opts.steps)]
    launches = {"n": 0}
    lmax = torch.empty(Ts, dtype=torch.float32, device=dev)
    ibuf = torch.empty(pipe.words, dtype=torch.int64, device=dev)
    out3 = torch.empty((3, Ts), dtype=torch.float64, device=dev)
    xe0_f, xref_f, uref_f = xe0.reshape(N, 5), xref_d.reshape(5, T + 1), uref_d.reshape(2, T)

    def one_step(i_timed=None):
        # the body of OccupancyPipeline.run for device inputs, with events around phase 1 (the rollout kernel; the four
        # device-to-device input copies and two tiny kernels inside that call are < 0.1 % of it)
        if i_timed is not None:
            ev_k0[i_timed].record(stream)
        _lib.check(L.dp_pipeline_rollout(pipe.handle, _lib.ptr(xe0_f), _lib.ptr(rho0), _lib.ptr(xref_f), _lib.ptr(uref_f), 0,
                                         float(args.dt_sim), N, _lib.ptr(lmax), sp), "dp_pipeline_rollout")
        if i_timed is not None:
            ev_k1[i_timed].record(stream)
        if world > 1:
            torch.distributed.all_reduce(lmax, op=torch.distributed.ReduceOp.MAX)
        _lib.check(L.dp_pipeline_bin(pipe.handle, _lib.ptr(lmax), _lib.ptr(ibuf), sp), "dp_pipeline_bin")
        if world > 1:
            torch.distributed.all_reduce(ibuf, op=torch.distributed.ReduceOp.SUM)
        _lib.check(L.dp_pipeline_finalize(pipe.handle, _lib.ptr(ibuf), _lib.ptr(out3), sp), "dp_pipeline_finalize")
        # kernels of this library per step: rollout, lmax decode, fused binning + cost, finalize (+ 2 memsets, 4 D2D copies,
        # + the NCCL kernels when N > 1)
        launches["n"] += 4

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(opts.warmup, 3)):
        one_step()
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches["n"] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(opts.steps):
        one_step(i)
    e1.record(stream)
    barrier()
    ms_total = D.max_over_ranks(e0.elapsed_time(e1), dev)
    clk = clocks.stop() if rank == 0 else None
    ms_step = ms_total / opts.steps
    value = world * N * T / (ms_step * 1e-3)
    k_ms = sum(a.elapsed_time(b) for a, b in zip(ev_k0, ev_k1)) / opts.steps
    k_ms = D.max_over_ranks(k_ms, dev)
    from density_planner_b200.pipeline import OccupancyResult
    checks = grid_checks(OccupancyResult(ibuf, out3, Ts, pipe.cx, pipe.cy, pipe.spec.mass_scale_log2, pipe.cost_scale_log2))
    checks["samples_total"] = world * N

    # ---- end to end through the public API with host tensors: pinned samples in, grids + costs out ------------------------
    res = pipe.run(xe0_host, rho0_host, xref, uref, args.dt_sim, to_host=True)  # warm the host path (pinned result buffers)
    barrier()
    e2e_steps = max(1, opts.steps)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = pipe.run(xe0_host, rho0_host, xref, uref, args.dt_sim, to_host=True)
        _ = float(res.cost_slices[-1]) + float(res.mass[0, 0, 0])  # the caller reads the result
    t_e2e = D.max_over_ranks((time.perf_counter() - t0) / e2e_steps, dev)
    e2e = {"value": world * N * T / t_e2e, "unit": UNIT, "samples_per_gpu": N,
           "h2d_bytes_per_step": int(4 * (N * 5 + N + 5 * (T + 1) + 2 * T)),
           "d2h_bytes_per_step": int(8 * pipe.words + 8 * 3 * Ts),
           "api": "OccupancyPipeline.run(xe0, rho0, xref_traj, uref_traj, dt, to_host=True) with pinned CPU tensors "
                  "(dp_rollout_fused at N=1; dp_pipeline_rollout / _bin / _finalize around two allreduces at N>1): H2D of the "
                  "samples, rollout, fused binning + cost, D2H of the int64 grids / int32 counts / per-slice costs",
           "ms_per_step": 1e3 * t_e2e, "numa_node_rank0": numa_node}

    # ---- the other public route: Car.compute_density with host tensors (full trajectory back to the host) -----------------
    e2e_traj = None
    if not opts.no_trajectory_e2e:
        Ne = min(N, SAMPLES)
        xe0_cpu = xe0_host[:Ne]
        _w = car.compute_density(xe0_cpu, xref, uref, args.dt_sim, log_density=True, cutting=True)
        del _w
        barrier()
        reps = 2
        t_sum = 0.0
        for _ in range(reps):
            t0 = time.perf_counter()
            xe_traj, rho_traj = car.compute_density(xe0_cpu, xref, uref, args.dt_sim, log_density=True, cutting=True)
            _ = float(rho_traj[0, 0, -1])
            t_sum += time.perf_counter() - t0
            del xe_traj, rho_traj
        t_tr = D.max_over_ranks(t_sum / reps, dev)
        e2e_traj = {"value": world * Ne * T / t_tr, "unit": UNIT, "samples_per_gpu": Ne, "ms_per_step": 1e3 * t_tr,
                    "h2d_bytes_per_step": int(4 * (Ne * 5 + 5 * (T + 1) + 2 * T)), "d2h_bytes_per_step": int(4 * (T + 1) * 6 * Ne),
                    "api": "Car.compute_density(xe0, xref_traj, uref_traj, dt, log_density=True) with CPU tensors "
                           "(dp_rollout_host: H2D + kernel + D2H of the full (N,5,T+1)+(N,1,T+1) trajectory; copy bound)"}

    # ---- BASELINE config 5: 100M samples x 100 steps in total, host samples in, allreduced grids + cost out -----------------
    cfg5 = None
    if opts.config5_samples > 0:
        x5, r5 = config5_samples(car, rank, world, opts.config5_samples)
        barrier()
        t0 = time.perf_counter()
        res5 = pipe.run(x5, r5, xref, uref, args.dt_sim, to_host=True)
        _ = float(res5.cost_slices[-1])
        t5 = D.max_over_ranks(time.perf_counter() - t0, dev)
        cfg5 = {"workload": "BASELINE config 5: %d samples x %d steps in total, %d samples on rank 0; pinned host samples in, "
                            "allreduced int64 grids + fixed-point cost out" % (opts.config5_samples, T, x5.shape[0]),
                "value": opts.config5_samples * T / t5, "unit": UNIT, "seconds": t5, "scaling": "strong",
                "h2d_bytes_rank0": int(4 * 6 * x5.shape[0]), "d2h_bytes_rank0": int(8 * pipe.words + 8 * 3 * Ts),
                "collectives_per_pass": 2 if world > 1 else 0,
                "checks": dict(grid_checks(res5), samples_total=opts.config5_samples,
                               note="integer fingerprints: identical at every --gpus N (same global sample set)")}
        del x5, r5, res5

    # ---- the fused binning + cost pass alone on the stored slices of the step (K3 inside the pipeline) --------------------
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    one_step()
    reps = 20
    for r in range(reps + 3):
        if r == 3:
            c0.record(stream)
        _lib.check(L.dp_pipeline_bin(pipe.handle, _lib.ptr(lmax), _lib.ptr(ibuf), sp), "dp_pipeline_bin")
    c1.record(stream)
    torch.cuda.synchronize(dev)
    bin_ms = c0.elapsed_time(c1) / reps
    bin_evals = N * Ts / (bin_ms * 1e-3)
    pipe.close()
    del pipe, ibuf

    # ---- BASELINE config 3: collision cost (forward, forward+backward) and fused binning on 1M weighted samples x 101
    #      time slices of the environment, [t][n] layout, inputs resident in HBM (1.2 GB streamed per pass > L2) -------
    cfg3 = None
    if rank == 0:
        spec = _lib.BinSpec()
        spec.mode = 0
        spec.geom = G._geom(args)
        spec.mass_scale_log2 = G.mass_scale_log2(1 << 27, 1.0)
        cells = (args.grid_size[0] + 1) * (args.grid_size[1] + 1)
        N3, T3 = SAMPLES, 101
        g3 = torch.Generator(device=dev).manual_seed(1)
        tt = torch.linspace(0, 1, T3, device=dev)
        x3 = (0.5 * torch.sin(3 * tt))[:, None] + 0.6 * torch.randn(T3, N3, device=dev, generator=g3)
        y3 = (-28 + 36 * tt)[:, None] + 0.6 * torch.randn(T3, N3, device=dev, generator=g3)
        r3 = torch.rand(T3, N3, device=dev, generator=g3)
        r3 = (r3 / r3.sum(dim=1, keepdim=True)).contiguous()
        c3 = torch.empty(T3, dtype=torch.float64, device=dev)
        d3 = torch.empty(T3, dtype=torch.float64, device=dev)
        gx3, gy3, gr3 = torch.empty_like(x3), torch.empty_like(x3), torch.empty_like(x3)
        m3 = torch.zeros((T3, cells), dtype=torch.int64, device=dev)
        n3 = torch.zeros((T3, cells), dtype=torch.int32, device=dev)

        def c_fwd():
            _lib.check(L.dp_cost_coll(denv.handle, _lib.ptr(x3), _lib.ptr(y3), 1, N3, _lib.ptr(r3), 1, N3, N3, T3, 0, _lib.ptr(c3),
                                      None, None, None, None, 0, 0, None, sp), "dp_cost_coll")

        def c_bwd():
            _lib.check(L.dp_cost_coll(denv.handle, _lib.ptr(x3), _lib.ptr(y3), 1, N3, _lib.ptr(r3), 1, N3, N3, T3, 0, _lib.ptr(c3),
                                      None, _lib.ptr(gx3), _lib.ptr(gy3), _lib.ptr(gr3), 1, N3, None, sp), "dp_cost_coll")

        def c_bin():
            _lib.check(L.dp_bin2d_occ(ctypes.byref(spec), denv.handle, _lib.ptr(x3), _lib.ptr(y3), 1, N3, _lib.ptr(r3), 1, N3, N3,
                                      T3, _lib.ptr(m3), _lib.ptr(n3), _lib.ptr(d3), sp), "dp_bin2d_occ")

        cfg3 = {"workload": "1M weighted samples x 101 time slices (BASELINE config 3), synthetic environment 241x401x101",
                "hbm_peak_gbs": measured_peaks()[0]["hbm_gbs"]}
        for name, fn, nbytes in (("cost_fwd", c_fwd, 12), ("cost_fwd_bwd", c_bwd, 24), ("binning_fused_occupancy_dot", c_bin, 12)):
            for _ in range(3):
                fn()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            for _ in range(10):
                fn()
            a1.record(stream)
            torch.cuda.synchronize(dev)
            ms = a0.elapsed_time(a1) / 10
            ev = N3 * T3 / (ms * 1e-3)
            cfg3[name] = {"ms": ms, "evals_per_s": ev, "algorithmic_bytes_per_eval": nbytes, "achieved_gbs": ev * nbytes / 1e9,
                          "frac_of_hbm_peak": ev * nbytes / 1e9 / cfg3["hbm_peak_gbs"]}
        del x3, y3, r3, gx3, gy3, gr3, m3, n3

    # ---- BASELINE config 1 shape: 20 samples x 1000 steps per reference (data_generation/compute_rawdata.py): one
    #      reference per launch (the reference's call pattern) and 296 references in one launch (dp_rollout_batch) --------
    cfg1 = None
    if rank == 0:
        flags = _lib.DP_LOG_DENSITY | _lib.DP_DENSITY
        n1, T1c, R1 = 20, 1000, 6 * 2 * 148
        torch.manual_seed(7)
        refs = []
        while len(refs) < 8:
            _, u1, x1 = car.get_valid_ref(args)
            if u1.shape[2] >= T1c:
                refs.append((u1[0, :, :T1c], x1[0, :, :T1c + 1]))
        ur1 = torch.stack([refs[r % 8][0] for r in range(R1)]).contiguous().to(dev)
        xr1 = torch.stack([refs[r % 8][1] for r in range(R1)]).contiguous().to(dev)
        xe1 = car.sample_xe0(R1 * n1).reshape(R1 * n1, 5).to(dev)
        xo1 = torch.empty((T1c + 1, 5, R1 * n1), dtype=torch.float32, device=dev)
        lo1 = torch.empty((T1c + 1, R1 * n1), dtype=torch.float32, device=dev)
        h = car.controller.controller.handle

        def one_ref():
            _lib.check(L.dp_rollout(h, _lib.ptr(xe1), _lib.ptr(xr1), _lib.ptr(ur1), None, float(args.dt_sim), n1, T1c, flags, 1,
                                    _lib.ptr(xo1), _lib.ptr(lo1), R1 * n1, None, None, sp), "dp_rollout")

        def all_refs():
            _lib.check(L.dp_rollout_batch(h, _lib.ptr(xe1), _lib.ptr(xr1), _lib.ptr(ur1), None, None, float(args.dt_sim), R1, n1,
                                          T1c, flags, 1, _lib.ptr(xo1), _lib.ptr(lo1), R1 * n1, None, None, sp), "dp_rollout_batch")

        cfg1 = {"workload": "%d samples x %d steps per reference, every step stored (BASELINE config 1 shape)" % (n1, T1c)}
        for name, fn, work in (("one_reference_per_launch", one_ref, n1 * T1c), ("batched", all_refs, R1 * n1 * T1c)):
            fn()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            for _ in range(3):
                fn()
            a1.record(stream)
            torch.cuda.synchronize(dev)
            ms = a0.elapsed_time(a1) / 3
            cfg1[name] = {"ms": ms, "sample_steps_per_s": work / (ms * 1e-3)}
        cfg1["batched"]["references_per_launch"] = R1
        del xo1, lo1

    if rank != 0:
        return
    peaks, peaks_src = measured_peaks()
    tf_peak = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))
    achieved_tf = N * T * FLOP_PER_SAMPLE_STEP / (k_ms * 1e-3) / 1e12
    executed_tf = N * T * EXECUTED_MMA_FLOP_PER_SAMPLE_STEP / (k_ms * 1e-3) / 1e12
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("rollout_dram_bytes_per_launch")
    roofline = {"bound": "tensor", "achieved": achieved_tf, "peak": tf_peak, "unit": "TFLOP/s",
                "frac": achieved_tf / tf_peak, "traffic": traffic,
                "kernel": "rollout_tc_kernel (tcgen05)", "kernel_ms": k_ms,
                "peak_source": "%s bf16_tflops_sustained (dense tensor pipe; the rollout is a K=128 dense contraction)" % peaks_src,
                "algorithmic_flop_per_sample_step": FLOP_PER_SAMPLE_STEP,
                "executed_mma_flop_per_sample_step": EXECUTED_MMA_FLOP_PER_SAMPLE_STEP,
                "executed_tflops": executed_tf, "frac_executed": executed_tf / tf_peak,
                "note": "value rows need 3 split-fp16 products for fp32-level accuracy (clip-mask parity), tangent rows 1: the tensor "
                        "pipe executes 1.69x the algorithmic FLOPs, so frac <= 0.59 x pipe utilisation; frac_executed is the "
                        "utilisation of the pipe itself",
                "hbm_bytes_per_launch_algorithmic": int(4 * Ts * 3 * N + 4 * 6 * N),
                "binning_cost_kernel": {"bound": "hbm", "evals_per_s": bin_evals, "achieved": 12 * bin_evals / 1e9,
                                        "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": 12 * bin_evals / 1e9 / peaks["hbm_gbs"],
                                        "bytes_per_eval": 12, "ms": bin_ms,
                                        "what": "dp_pipeline_bin on the step's %d stored slices: memset of the 12.9 MB integer "
                                                "buffer + ONE pass (weights, grids, counts, cost)" % Ts}}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": opts.steps, "warmup": max(opts.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "CAR rollout + Liouville density, %d samples x %d steps per GPU (%s), device-resident pipeline: "
                                   "density weights, occupancy grids and collision cost on %d stored slices" % (
                                       N, T, "BASELINE config 2" if N == SAMPLES else "--samples", Ts),
                       "samples_per_gpu": N, "horizon_steps": T, "stored_slices": Ts,
                       "controller": "trained C3M controller_CAR_3layers (43,592 fp32 weights)",
                       "l2_policy": "inputs+outputs per step (%.0f MB of slices + 13 MB of grids written) exceed the 126 MB L2" % (
                           4e-6 * Ts * 3 * N),
                       "parallelism": "samples sharded, %d rank(s); 2 collectives per step (MAX of %d floats, SUM of one int64 "
                                      "buffer with grids, counts and cost)" % (world, Ts)},
            "e2e": e2e, "e2e_trajectory": e2e_traj, "config5": cfg5, "gpu_launches": launches["n"], "clocks": clk, "roofline": roofline,
            "collision_cost_evals_per_s": cfg3["cost_fwd"]["evals_per_s"] if cfg3 else None,
            "collision_cost_config3": cfg3, "rawdata_config1": cfg1, "checks": checks}
    if not opts.no_cpu_baseline and world == 1:
        try:
            ref = cpu_reference_rate()
        except Exception as exc:  # noqa: BLE001
            print("bench.py: the reference under baseline/_ref could not be run (%r)" % (exc,), file=sys.stderr)
            ref = None
        port, _, _ = cpu_oracle_rate()
        if ref is not None:  # the reference's own PyTorch path; the C port beside it
            base = ref[0]
            base["port"] = port
        else:
            base = port
        line["cpu_baseline"] = base
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    main()
