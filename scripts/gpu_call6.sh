#!/bin/bash
# GPU call 6 of round 2: packed raw-data launches (several 20-sample references per tile): parity + timing; full suite; bench.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_rollout.py -m gpu -q -x > $O/r2c6_rollout.log 2>&1; echo "rollout tests rc=$?" | tee -a $O/r2c6_rollout.log
tail -5 $O/r2c6_rollout.log
timeout 1500 python -m pytest tests -m gpu -q > $O/r2c6_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r2c6_tests.log
tail -8 $O/r2c6_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r2c6_bench_pack.json 2> $O/r2c6_bench_pack.err; echo "bench rc=$?"
DP_TC_PACK=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r2c6_bench_nopack.json 2> $O/r2c6_bench_nopack.err; echo "bench nopack rc=$?"
python - <<'PY'
import json
for f in ("r2c6_bench_pack.json", "r2c6_bench_nopack.json"):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, d["value"], d["roofline"]["frac"], d["rawdata_config1"])
    except Exception as e:
        print(f, "ERR", e)
PY
echo done
