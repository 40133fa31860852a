#!/bin/bash
# GPU call 13 of round 2: cost kernels: more waves of blocks (finer load balance) A/B.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for K in "4 4" "4 6" "4 8" "4 12" "4 16" "6 8" "4 8"; do
  set -- $K
  echo "== DP_COST_FWD_MINB=$1 DP_COST_WAVES=$2" | tee -a $O/r2c13_costfwd.log
  DP_COST_FWD_MINB=$1 DP_COST_WAVES=$2 timeout 300 python scripts/time_cost.py 2>&1 | grep -E "^cost" | tee -a $O/r2c13_costfwd.log
done
echo done
