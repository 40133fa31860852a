#!/bin/bash
# GPU call 14 of round 2: per-cell desired positions in the env table (cost + fused pass): parity, timing; binning waves rounded down.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_grid.py tests/test_gpu_pipeline.py tests/test_gpu_planner_terms.py tests/test_gpu_planner_loop.py -m gpu -q -x > $O/r2c14_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r2c14_tests.log
tail -3 $O/r2c14_tests.log
for K in 8 6 10 12 8; do
  echo "== DP_BIN_BLOCKS_PER_SM=$K" | tee -a $O/r2c14_time.log
  DP_BIN_BLOCKS_PER_SM=$K timeout 300 python scripts/time_cost.py 2>&1 | grep -E "^cost|^bin2d" | tee -a $O/r2c14_time.log
done
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-trajectory-e2e --config5-samples 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', d['value'], 'fused pass', d['roofline']['binning_cost_kernel']['ms'], d['roofline']['binning_cost_kernel']['frac'])
print(json.dumps(d['collision_cost_config3']))
print(json.dumps(d['checks']))
" | tee -a $O/r2c14_time.log
echo done
