#!/bin/bash
# GPU call 11 of round 2: cost / binning kernels with the XU-free cell computation and the L2 prefetch knob: parity + A/B.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_grid.py tests/test_gpu_pipeline.py tests/test_gpu_planner_terms.py -m gpu -q -x > $O/r2c11_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r2c11_tests.log
tail -3 $O/r2c11_tests.log
for K in "0 0" "1 1" "2 2" "4 4" "0 0" "2 1" "3 3"; do
  set -- $K
  echo "== DP_COST_PREFETCH=$1 DP_BIN_PREFETCH=$2" | tee -a $O/r2c11_prefetch.log
  DP_COST_PREFETCH=$1 DP_BIN_PREFETCH=$2 timeout 300 python scripts/time_cost.py 2>&1 | grep -E "^cost|^bin2d" | tee -a $O/r2c11_prefetch.log
done
echo done
