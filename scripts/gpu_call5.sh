#!/bin/bash
# GPU call 5 of round 2: the fused density-NN kernel (dp_dnn_*) against fp64 and the reference golden, planner suites on it.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_planner_terms.py -m gpu -q -x -k "density_nn" > $O/r2c5_dnn.log 2>&1; echo "dnn rc=$?" | tee -a $O/r2c5_dnn.log
tail -30 $O/r2c5_dnn.log
timeout 1500 python -m pytest tests -m gpu -q > $O/r2c5_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r2c5_tests.log
timeout 300 python scripts/time_dnn.py > $O/r2c5_time_dnn.log 2>&1; cat $O/r2c5_time_dnn.log
echo done
