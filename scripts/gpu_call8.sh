#!/bin/bash
# GPU call 8 of round 2 (8 GPUs): the pipeline bench at N=8 (weak scaling value / e2e, config 5 strong scaling, integer fingerprints).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > $O/r2c8_gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 > $O/r2c8_bench8.json 2> $O/r2c8_bench8.err; echo "bench8 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 10 --warmup 3 > $O/r2c8_bench4.json 2> $O/r2c8_bench4.err; echo "bench4 rc=$?"
tail -3 $O/r2c8_bench8.err
echo done
