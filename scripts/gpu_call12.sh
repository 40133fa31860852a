#!/bin/bash
# GPU call 12 of round 2: forward cost kernel: resident blocks per SM / waves A/B (prefetch distance 1 is the default now).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for K in "6 2" "4 2" "8 2" "6 4" "4 4" "6 1" "4 1" "6 2"; do
  set -- $K
  echo "== DP_COST_FWD_MINB=$1 DP_COST_WAVES=$2" | tee -a $O/r2c12_costfwd.log
  DP_COST_FWD_MINB=$1 DP_COST_WAVES=$2 timeout 300 python scripts/time_cost.py 2>&1 | grep -E "^cost" | tee -a $O/r2c12_costfwd.log
done
echo done
