#!/bin/bash
# GPU call 21 of round 2: K1 compiler-flag / split variants against the default, same box, interleaved.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for rep in 1 2; do
for V in "" ptxO2 sp1; do
  DP_LIB_VARIANT=$V timeout 300 python scripts/time_rollout.py tc 1000000 100 10 8 2>&1 | tail -1 | sed "s/^/variant=${V:-default} /" | tee -a $O/r2c21_variants.log
done
done
echo done
