#!/bin/bash
# GPU call 15 of round 2: evidence pass of the final build -- parity suite, smoke, bench (20 steps), reference arm, launch
# list, ncu --set full of the cost kernels and the rollout kernel (each after the plain command exited 0).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/r2c15_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r2c15_tests.log
tail -4 $O/r2c15_tests.log
timeout 300 python __graft_entry__.py smoke > $O/r2c15_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r2c15_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/r2c15_bench.json 2> $O/r2c15_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2c15_refarm.json 2> $O/r2c15_refarm.err; echo "refarm rc=$?"
BARGS="--steps 2 --warmup 1 --no-cpu-baseline --no-trajectory-e2e --config5-samples 0"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2c15_launches.csv python bench.py $BARGS > $O/r2c15_ncu_launches.log 2>&1
timeout 300 python scripts/time_cost.py > $O/r2c15_time_cost.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cost_vec4 -s 8 -c 1 -o $O/r2c15_costfwd python scripts/time_cost.py > $O/r2c15_ncu_costfwd.log 2>&1; echo "ncu cost fwd rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cost_vec4 -s 20 -c 1 -o $O/r2c15_costbwd python scripts/time_cost.py > $O/r2c15_ncu_costbwd.log 2>&1; echo "ncu cost bwd rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bin2d_lean -s 20 -c 1 -o $O/r2c15_binocc python scripts/time_cost.py > $O/r2c15_ncu_binocc.log 2>&1; echo "ncu bin occ rc=$?"
cat $O/r2c15_time_cost.log
echo done
