#!/bin/bash
# GPU call 18 of round 2: forward cost kernel with five resident blocks per SM (51 registers) against four.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for K in "4 4" "5 4" "5 3" "5 5" "4 4"; do
  set -- $K
  echo "== DP_COST_FWD_MINB=$1 DP_COST_WAVES=$2" | tee -a $O/r2c18_time.log
  DP_COST_FWD_MINB=$1 DP_COST_WAVES=$2 timeout 300 python scripts/time_cost.py 2>&1 | grep -E "^cost fwd " | tee -a $O/r2c18_time.log
done
echo done
