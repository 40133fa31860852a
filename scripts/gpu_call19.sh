#!/bin/bash
# GPU call 19 of round 2: compute_data on the packed launch by default; full parity suite; smoke.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/r2c19_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r2c19_tests.log
tail -5 $O/r2c19_tests.log
timeout 300 python __graft_entry__.py smoke > $O/r2c19_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r2c19_smoke.log
echo done
