#!/bin/bash
# GPU call 7 of round 2: in-step fused binning pass (11 slices): block-count knobs A/B.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for K in "6144 8" "24576 8" "49152 8" "24576 2" "49152 2" "98304 4" "12288 4"; do
  set -- $K
  DP_BIN_MIN_SAMPLES=$1 DP_BIN_BLOCKS_PER_SM=$2 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-trajectory-e2e --config5-samples 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
b=d['roofline']['binning_cost_kernel']; c=d['collision_cost_config3']['binning_fused_occupancy_dot']
print('min_samples=$1 blocks_per_sm=$2: in-step fused pass %.4f ms (%.2f of HBM) | config-3 binning+occ %.4f ms (%.2f)' % (b['ms'], b['frac'], c['ms'], c['frac_of_hbm_peak']))
" | tee -a $O/r2c7_binstep.log
done
echo done
