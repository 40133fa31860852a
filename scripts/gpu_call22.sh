#!/bin/bash
# GPU call 22 of round 2: tanh activation in the fused density-NN kernel; planner suites.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_planner_terms.py tests/test_gpu_planner_loop.py tests/test_gpu_planner_real.py -m gpu -q -x > $O/r2c22_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r2c22_tests.log
tail -15 $O/r2c22_tests.log
echo done
