#!/usr/bin/env python
"""CUDA-event timing of the density-NN branch at the planner's size (101 time points x 500 samples = 50,500 rows):
fused tcgen05 kernel (forward, forward + backward) against the same batched pass in plain torch (cuBLAS fp32)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import density_planner_b200 as dp
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 50500
x = torch.randn(rows, 31, device="cuda")
for backend in ("kernel", "torch"):
    nn = dp.DensityNN(device="cuda", backend=backend)
    def fwd():
        with torch.no_grad():
            return nn.mlp(x)
    def fwdbwd():
        xx = x.clone().requires_grad_(True)
        nn.mlp(xx).sum().backward()
    for name, fn in (("forward", fwd), ("forward+backward", fwdbwd)):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        flop = 2 * rows * (31 * 150 + 6 * 150 * 150 + 150 * 6) * (2 if "backward" in name else 1)
        print("%-7s %-17s rows=%d: %.3f ms  %.1f algorithmic TFLOP/s" % (backend, name, rows, ms, flop / ms / 1e9))
