"""Device-resident pipeline: sampled initial states -> rollout with Liouville log-density -> density weights ->
time-sliced occupancy grids + collision cost, with the trajectories never leaving the GPU (dp_pipeline_* / dp_rollout_fused).

Reference chain it replaces (one call instead of four passes over host tensors):
  motion_planning/simulation_objects.py:288-299  EgoVehicle.predict_density, LE branch + normaliser
  motion_planning/utils.py:255-280               pred2grid, per time slice
  motion_planning/MotionPlanner.py:192-224       get_cost_coll
  motion_planning/utils.py:231-252               check_collision (occupancy dot product)

Multi-GPU: every rank rolls out its own shard of the samples; two collectives make the result global -- a MAX allreduce
of the per-slice maximum of the log-density (Ts floats) and ONE SUM allreduce of the int64 result buffer (mass grids,
counts, fixed-point cost).  All accumulators are integers, so the summed buffer is bit-identical for 1/2/4/8 ranks given
the same global sample set (tests/test_dist_gloo.py, tests/test_gpu_pipeline.py, bench.py's config-5 check).
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib, grid as _grid

# fixed-point scales are chosen from DECLARED global bounds so that every shard and every world size uses the same ones
DEFAULT_TOTAL_SAMPLES_BOUND = 1 << 27   # 134M >= the 100M samples of BASELINE config 5
COST_SCALE_LOG2 = 20                    # w <= 2^1, p <= 1, |des - pos|^2 < 2^14 over this grid, 2^27 samples: < 2^62


class OccupancyResult:
    """Result of one pipeline pass.  mass / count: (Ts, G0+1, G1+1) integer grids (cells on the upper clamp row / column
    G are samples beyond the grid, as in pred2grid); weight_sum, cost_slices, occ_dot: (Ts,) float64."""

    def __init__(self, ibuf, out3, Ts, cx, cy, mass_scale_log2, cost_scale_log2):
        gc = Ts * cx * cy
        self.ibuf = ibuf
        self.mass = ibuf[:gc].view(Ts, cx, cy)
        self.count = ibuf[gc:gc + (gc + 1) // 2].view(torch.int32)[:gc].view(Ts, cx, cy)
        self.cost_fixed = ibuf[gc + (gc + 1) // 2:]
        self.weight_sum, self.cost_slices, self.occ_dot = out3[0], out3[1], out3[2]
        self.mass_scale_log2, self.cost_scale_log2 = mass_scale_log2, cost_scale_log2

    @property
    def cost(self):
        """get_cost_coll of the normalised density: sum over the slices (MotionPlanner.py:192-224)."""
        return self.cost_slices.sum()

    def density(self, t):
        """Sum-binned, normalised density of slice t on the environment grid (G0, G1), float64."""
        m = self.mass[t, :-1, :-1].to(torch.float64)
        return m / self.mass[t].sum().to(torch.float64)

    def pred2grid(self, t):
        """pred2grid's convention for slice t (mean density per occupied cell, normalised to 1; utils.py:268-279)."""
        m = self.mass[t].to(torch.float64)
        c = self.count[t]
        mean = torch.where(c > 0, m / c.clamp(min=1).to(torch.float64), torch.zeros_like(m))
        return (mean / mean.sum())[:-1, :-1]


class OccupancyPipeline:
    """Owns a dp_pipeline: device workspace for up to `capacity` samples per call on this GPU."""

    def __init__(self, system, denv, args, T, stride, capacity, total_samples_bound=DEFAULT_TOTAL_SAMPLES_BOUND,
                 rho0_bound=1.0, cost_scale_log2=COST_SCALE_LOG2):
        L = _lib.lib()
        self.system, self.denv, self.args = system, denv, args
        self.dc = system.controller.controller
        self.device = self.dc.device
        if denv.device != self.device:
            raise _lib.DensityPlannerError("controller on %s, environment on %s" % (self.device, denv.device))
        self.T, self.stride, self.Ts = int(T), int(stride), int(T) // int(stride) + 1
        self.capacity = int(capacity)
        spec = _lib.BinSpec()
        spec.mode = 0
        spec.geom = _grid._geom(args)
        # weights are exp(l - max l) * rho0 <= rho0_bound; the scale depends on declared bounds only (never on the shard)
        spec.mass_scale_log2 = _grid.mass_scale_log2(total_samples_bound, rho0_bound)
        self.spec = spec
        self.cost_scale_log2 = int(cost_scale_log2)
        self.cx, self.cy = spec.geom.gx + 1, spec.geom.gy + 1
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(L.dp_pipeline_create(self.dc.handle, denv.handle, ctypes.byref(spec), self.T, self.stride, self.capacity,
                                            self.cost_scale_log2, ctypes.byref(h)), "dp_pipeline_create")
        self.handle = h
        w = ctypes.c_int64()
        _lib.check(L.dp_pipeline_words(h, ctypes.byref(w)), "dp_pipeline_words")
        self.words = int(w.value)
        self._host = None  # pinned result buffers of the host entry

    def close(self):
        if getattr(self, "handle", None):
            _lib.lib().dp_pipeline_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _prep(self, xe0, rho0, xref_traj, uref_traj):
        N = xe0.shape[0]
        if N > self.capacity:
            raise _lib.DensityPlannerError("N=%d exceeds the pipeline capacity %d" % (N, self.capacity))
        if uref_traj.shape[2] != self.T or xref_traj.shape[2] != self.T + 1:
            raise _lib.DensityPlannerError("the pipeline was built for T=%d (got uref %d, xref %d time points)"
                                           % (self.T, uref_traj.shape[2], xref_traj.shape[2]))
        on_host = not xe0.is_cuda
        dev = torch.device("cpu") if on_host else self.device
        f = lambda t, shape: t.detach().reshape(shape).to(device=dev, dtype=torch.float32).contiguous()
        return (N, on_host, f(xe0, (N, 5)), None if rho0 is None else f(rho0, (N,)), f(xref_traj, (5, self.T + 1)),
                f(uref_traj, (2, self.T)))

    def run(self, xe0, rho0, xref_traj, uref_traj, dt, to_host=False, local=False):
        """xe0 (N,5,1), rho0 (N,1,1) or None, xref_traj (1,5,T+1), uref_traj (1,2,T): CPU tensors (copied in inside the call;
        pin them for speed) or tensors on the pipeline's GPU.  With torch.distributed initialised every rank passes ITS shard
        and gets the global result (`local=True` skips the collectives: the result of this rank's samples alone).
        Returns an OccupancyResult on the GPU (or on the host with to_host=True)."""
        L = _lib.lib()
        N, on_host, xe0_c, rho0_c, xref_c, uref_c = self._prep(xe0, rho0, xref_traj, uref_traj)
        world = dist.get_world_size() if (dist.is_initialized() and not local) else 1
        dev = self.device
        if to_host and self._host is None:
            self._host = (torch.empty(self.words, dtype=torch.int64, pin_memory=True),
                          torch.empty((3, self.Ts), dtype=torch.float64, pin_memory=True))
        if world == 1 and on_host and to_host:
            # one C call: H2D, three phases, D2H of the integer buffer and the per-slice results
            ibuf, out3 = self._host
            _lib.check(L.dp_rollout_fused(self.handle, _lib.ptr(xe0_c), _lib.ptr(rho0_c), _lib.ptr(xref_c), _lib.ptr(uref_c),
                                          float(dt), N, _lib.ptr(ibuf), _lib.ptr(out3)), "dp_rollout_fused")
            return OccupancyResult(ibuf, out3, self.Ts, self.cx, self.cy, self.spec.mass_scale_log2, self.cost_scale_log2)
        with torch.cuda.device(dev):
            sp = _lib.current_stream(dev)
            lmax = torch.empty(self.Ts, dtype=torch.float32, device=dev)
            _lib.check(L.dp_pipeline_rollout(self.handle, _lib.ptr(xe0_c), _lib.ptr(rho0_c), _lib.ptr(xref_c), _lib.ptr(uref_c),
                                             1 if on_host else 0, float(dt), N, _lib.ptr(lmax), sp), "dp_pipeline_rollout")
            if world > 1:
                dist.all_reduce(lmax, op=dist.ReduceOp.MAX)      # collective 1 of 2: Ts floats
            ibuf = torch.empty(self.words, dtype=torch.int64, device=dev)
            _lib.check(L.dp_pipeline_bin(self.handle, _lib.ptr(lmax), _lib.ptr(ibuf), sp), "dp_pipeline_bin")
            if world > 1:
                dist.all_reduce(ibuf, op=dist.ReduceOp.SUM)      # collective 2 of 2: grids, counts and cost in one buffer
            out3 = torch.empty((3, self.Ts), dtype=torch.float64, device=dev)
            _lib.check(L.dp_pipeline_finalize(self.handle, _lib.ptr(ibuf), _lib.ptr(out3), sp), "dp_pipeline_finalize")
            if to_host:  # pinned result buffers, reused from call to call (the caller consumes the result before the next run)
                self._host[0].copy_(ibuf, non_blocking=True)
                self._host[1].copy_(out3, non_blocking=True)
                torch.cuda.current_stream(dev).synchronize()
                ibuf, out3 = self._host
            elif on_host:
                torch.cuda.current_stream(dev).synchronize()  # the asynchronous H2D copies read the caller's host tensors
        self.lmax = lmax
        return OccupancyResult(ibuf, out3, self.Ts, self.cx, self.cy, self.spec.mass_scale_log2, self.cost_scale_log2)

    def slices(self, N):
        """Copies of the stored slices of the last run: x, y (Ts, N) absolute positions and l (Ts, N) log-density."""
        xy = torch.empty((self.Ts, 2, N), dtype=torch.float32, device=self.device)
        l = torch.empty((self.Ts, N), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().dp_pipeline_copy_slices(self.handle, _lib.ptr(xy), _lib.ptr(l), _lib.current_stream(self.device)),
                       "dp_pipeline_copy_slices")
        return xy[:, 0, :], xy[:, 1, :], l
