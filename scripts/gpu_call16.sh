#!/bin/bash
# GPU call 16 of round 2: 32-bit loop indices / constant clamp bounds in the streaming kernels: parity + timing.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_grid.py tests/test_gpu_pipeline.py -m gpu -q -x > $O/r2c16_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r2c16_tests.log
tail -2 $O/r2c16_tests.log
for i in 1 2; do timeout 300 python scripts/time_cost.py 2>&1 | grep -E "^cost|^bin2d" | tee -a $O/r2c16_time.log; done
echo done
