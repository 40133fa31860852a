#!/usr/bin/env python
"""CUDA-event timing of the collision-cost kernel (forward, forward+backward) and of the occupancy binning at the
BASELINE config-3 size: python scripts/time_cost.py [N] [Ts]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import density_planner_b200 as dp
from density_planner_b200 import _lib, grid as G
from bench import synthetic_env
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
Ts = int(sys.argv[2]) if len(sys.argv) > 2 else 101
args = dp.default_args()
dev = torch.device("cuda", 0)
L = dp.lib()
grid, gx, gy = synthetic_env(args, 0, dev)
denv = dp.DeviceEnv(grid, gx, gy, args, device=dev)
g = torch.Generator(device="cuda").manual_seed(1)
tt = torch.linspace(0, 1, Ts, device=dev)
x = (0.5 * torch.sin(3 * tt))[:, None] + 0.6 * torch.randn(Ts, N, device=dev, generator=g)
y = (-28 + 36 * tt)[:, None] + 0.6 * torch.randn(Ts, N, device=dev, generator=g)
rho = torch.rand(Ts, N, device=dev, generator=g)
rho = rho / rho.sum(dim=1, keepdim=True)
cost = torch.empty(Ts, dtype=torch.float64, device=dev)
gxo, gyo, gro = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
def fwd():
    _lib.check(L.dp_cost_coll(denv.handle, _lib.ptr(x), _lib.ptr(y), 1, N, _lib.ptr(rho), 1, N, N, Ts, 0, _lib.ptr(cost), None,
                              None, None, None, 0, 0, None, sp), "cost")
def fwdbwd():
    _lib.check(L.dp_cost_coll(denv.handle, _lib.ptr(x), _lib.ptr(y), 1, N, _lib.ptr(rho), 1, N, N, Ts, 0, _lib.ptr(cost), None,
                              _lib.ptr(gxo), _lib.ptr(gyo), _lib.ptr(gro), 1, N, None, sp), "cost")
spec = _lib.BinSpec(); spec.mode = 0; spec.geom = G._geom(args); spec.mass_scale_log2 = G.mass_scale_log2(N, 1.0)
cells = (args.grid_size[0] + 1) * (args.grid_size[1] + 1)
mass = torch.zeros((Ts, cells), dtype=torch.int64, device=dev); count = torch.zeros((Ts, cells), dtype=torch.int32, device=dev)
def bin2d():
    _lib.check(L.dp_bin2d(ctypes.byref(spec), _lib.ptr(x), _lib.ptr(y), 1, N, _lib.ptr(rho), 1, N, N, Ts, _lib.ptr(mass),
                          _lib.ptr(count), sp), "bin2d")
dot = torch.empty(Ts, dtype=torch.float64, device=dev)
def bin2d_occ():
    _lib.check(L.dp_bin2d_occ(ctypes.byref(spec), denv.handle, _lib.ptr(x), _lib.ptr(y), 1, N, _lib.ptr(rho), 1, N, N, Ts,
                              _lib.ptr(mass), _lib.ptr(count), _lib.ptr(dot), sp), "bin2d_occ")
peak = 6554.9
print("knobs: DP_BIN_THREADS=%s DP_BIN_AGG=%s" % (os.environ.get("DP_BIN_THREADS", "default"), os.environ.get("DP_BIN_AGG", "default")))
for name, fn, bytes_per in (("cost fwd", fwd, 12), ("cost fwd+bwd", fwdbwd, 24), ("bin2d", bin2d, 12), ("bin2d_occ", bin2d_occ, 12)):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); reps = 10
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    ev = N * Ts / ms * 1e3
    print("%-14s N=%d Ts=%d: %.3f ms  %.3e evals/s  %.0f GB/s algorithmic (%.2f of %.0f)" % (name, N, Ts, ms, ev, ev * bytes_per / 1e9, ev * bytes_per / 1e9 / peak, peak))
