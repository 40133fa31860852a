#!/bin/bash
# GPU call 17 of round 2: binning kernels at 640 threads per block (51 registers) against 768 (40).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for NT in 768 640 768 640; do
  echo "== DP_BIN_THREADS=$NT" | tee -a $O/r2c17_time.log
  DP_BIN_THREADS=$NT timeout 300 python scripts/time_cost.py 2>&1 | grep -E "^bin2d" | tee -a $O/r2c17_time.log
  DP_BIN_THREADS=$NT timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-trajectory-e2e --config5-samples 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('fused pass', d['roofline']['binning_cost_kernel']['ms'], d['roofline']['binning_cost_kernel']['frac'])" | tee -a $O/r2c17_time.log
done
echo done
