#!/bin/bash
# GPU call 10 of round 2: K1 A/B -- chained tanh groups (DP_TIE) against the default, same box, interleaved.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for rep in 1 2; do
for V in "" tie1 tie2 tie3 tie3p; do
  DP_LIB_VARIANT=$V timeout 300 python scripts/time_rollout.py tc 1000000 100 10 8 2>&1 | tail -1 | sed "s/^/variant=${V:-default} /" | tee -a $O/r2c10_tie.log
done
done
echo done
