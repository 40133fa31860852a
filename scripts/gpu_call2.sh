#!/bin/bash
# GPU call 2 of round 2 (2 GPUs): full parity suite incl. the 2-rank NCCL pipeline parity, bench at N=1 and N=2.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/r2c2_gpu.txt
timeout 1200 python -m pytest tests -m gpu -q > $O/r2c2_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r2c2_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2c2_bench1.json 2> $O/r2c2_bench1.err; echo "bench1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2c2_bench2.json 2> $O/r2c2_bench2.err; echo "bench2 rc=$?"
echo done
