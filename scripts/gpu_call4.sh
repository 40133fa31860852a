#!/bin/bash
# GPU call 4 of round 2: the reference planner through dp.attach (config 4), binning A/B with the flush-time occupancy dot,
# reference CPU arm with the real reference.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for S in 0 1; do
  timeout 1500 python tests/run_planner_attached.py $S $O/r2c4_attached_seed$S.npz > $O/r2c4_planner_seed$S.log 2>&1; echo "planner seed $S rc=$?"
  tail -3 $O/r2c4_planner_seed$S.log
done
timeout 1500 python -m pytest tests/test_gpu_planner_real.py tests/test_gpu_grid.py tests/test_gpu_pipeline.py -m gpu -q -s > $O/r2c4_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r2c4_tests.log
for K in "768 8" "768 16" "1024 8" "768 32"; do
  set -- $K
  echo "== DP_BIN_THREADS=$1 DP_BIN_BLOCKS_PER_SM=$2" | tee -a $O/r2c4_k3.log
  DP_BIN_THREADS=$1 DP_BIN_BLOCKS_PER_SM=$2 timeout 300 python scripts/time_cost.py 2>&1 | grep -E "^bin2d" | tee -a $O/r2c4_k3.log
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2c4_refarm.json 2> $O/r2c4_refarm.err; echo "refarm rc=$?"
echo done
