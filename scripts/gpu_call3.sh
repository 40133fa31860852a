#!/bin/bash
# GPU call 3 of round 2: lean binning kernel A/B (threads x blocks per SM), parity of the grid / pipeline suites, ncu of the lean kernel.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_grid.py tests/test_gpu_pipeline.py tests/test_gpu_rollout.py -m gpu -q -x > $O/r2c3_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r2c3_tests.log
for K in "768 4" "768 2" "1024 4" "1024 2" "512 4" "768 8"; do
  set -- $K
  echo "== DP_BIN_THREADS=$1 DP_BIN_BLOCKS_PER_SM=$2" | tee -a $O/r2c3_k3.log
  DP_BIN_THREADS=$1 DP_BIN_BLOCKS_PER_SM=$2 timeout 300 python scripts/time_cost.py 2>&1 | grep -E "^bin2d|^pipeline" | tee -a $O/r2c3_k3.log
done
echo "== legacy general kernel (DP_BIN_AGG=1 forces it with aggregation)" >> $O/r2c3_k3.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bin2d_lean -s 8 -c 1 -o $O/r2c3_bin2d_lean python scripts/time_cost.py > $O/r2c3_ncu_bin.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bin2d_lean -s 20 -c 1 -o $O/r2c3_bin2d_lean_occ python scripts/time_cost.py > $O/r2c3_ncu_bin_occ.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > $O/r2c3_bench1.json 2> $O/r2c3_bench1.err; echo "bench rc=$?"
echo done
