#!/bin/bash
# GPU call 1 of round 2: parity suite, smoke, bench, K1 variant A/B, K3 knob A/B, ncu of the binning kernel.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/r2_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2_tests1.log 2>&1; echo "tests rc=$?" | tee -a $O/r2_tests1.log
timeout 300 python __graft_entry__.py smoke > $O/r2_smoke1.log 2>&1; echo "smoke rc=$?" | tee -a $O/r2_smoke1.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/r2_bench1.json 2> $O/r2_bench1.err; echo "bench rc=$?"
for V in 0 3 1 2 0 3; do
  DP_TC_V=$V timeout 300 python scripts/time_rollout.py tc 1000000 100 10 8 2>&1 | tail -1 | sed "s/^/V=$V /" | tee -a $O/r2_k1_variants.log
done
for V in 0 3 1; do
  echo "== DP_TC_V=$V" >> $O/r2_precision.log
  DP_TC_V=$V timeout 600 python scripts/precision_report.py tc >> $O/r2_precision.log 2>&1
done
for K in "1024 2" "1024 0" "1024 1" "512 2" "512 0"; do
  set -- $K
  DP_BIN_THREADS=$1 DP_BIN_AGG=$2 timeout 300 python scripts/time_cost.py 2>&1 | tail -5 | tee -a $O/r2_k3_knobs.log
done
timeout 300 python scripts/time_cost.py > $O/r2_plain_cost.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bin2d_kernel -s 8 -c 1 -o $O/r2_bin2d python scripts/time_cost.py > $O/r2_ncu_bin.log 2>&1
echo done
