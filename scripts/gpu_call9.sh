#!/bin/bash
# GPU call 9 of round 2: evidence pass -- parity suite, bench, launch list, ncu --set full of the rollout / fused binning /
# density-NN kernels (each after the plain command exited 0).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/r2c9_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r2c9_tests.log
tail -4 $O/r2c9_tests.log
timeout 300 python __graft_entry__.py smoke > $O/r2c9_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r2c9_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/r2c9_bench.json 2> $O/r2c9_bench.err; echo "bench rc=$?"
BARGS="--steps 2 --warmup 1 --no-cpu-baseline --no-trajectory-e2e --config5-samples 0"
timeout 600 python bench.py $BARGS > $O/r2c9_bench_short.json 2>/dev/null && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2c9_launches.csv python bench.py $BARGS > $O/r2c9_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_tc_kernel -s 2 -c 1 -o $O/r2c9_rollout python bench.py $BARGS > $O/r2c9_ncu_rollout.log 2>&1; echo "ncu rollout rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bin2d_lean -s 1 -c 1 -o $O/r2c9_binfused python bench.py $BARGS > $O/r2c9_ncu_binfused.log 2>&1; echo "ncu bin rc=$?"
timeout 300 python scripts/time_dnn.py > $O/r2c9_time_dnn.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dnn_kernel -s 6 -c 2 -o $O/r2c9_dnn python scripts/time_dnn.py > $O/r2c9_ncu_dnn.log 2>&1; echo "ncu dnn rc=$?"
echo done
