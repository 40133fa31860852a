#!/bin/bash
# GPU call 20 of round 2: repeatability -- the parity suite twice more with -x (as the driver runs it), the bench three times.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for i in 1 2; do timeout 1500 python -m pytest tests -x -q -m gpu > $O/r2c20_tests$i.log 2>&1; echo "tests run $i rc=$?"; tail -1 $O/r2c20_tests$i.log; done
for i in 1 2 3; do
  timeout 900 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
c=d['collision_cost_config3']
print('run $i: value %.4e (%.3f ms) e2e %.4e frac %.4f clock %s | cost fwd %.3f ms bwd %.3f ms bin %.3f ms | fused %.4f ms | cfg1 %.3e'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'],d['clocks']['sm_mhz'],c['cost_fwd']['ms'],c['cost_fwd_bwd']['ms'],c['binning_fused_occupancy_dot']['ms'],d['roofline']['binning_cost_kernel']['ms'],d['rawdata_config1']['batched']['sample_steps_per_s']))" | tee -a $O/r2c20_bench.log
done
echo done
