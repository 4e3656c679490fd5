#!/bin/bash
# 2-GPU pipeline session: pp2 with micro-batches of 2 and of 8 sequences.
mkdir -p gpurun_out; N=2
run() { timeout -s KILL $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
F='^W0\|OMP_NUM\|^\*\*\*'
echo "== pp2 micro-batch 2"; run 150 29531 bench.py --gpus 2 --steps 4 --warmup 3 --no-e2e --parallelism pp2 --batch 2 2>&1 | grep -v "$F" | tail -1 | tee gpurun_out/bench_pp2_mb2.log | cut -c1-900
echo "== pp2 micro-batch 8"; run 200 29532 bench.py --gpus 2 --steps 4 --warmup 3 --parallelism pp2 --batch 8 2>&1 | grep -v "$F" | tail -3 | tee gpurun_out/bench_pp2_mb8.log | cut -c1-1500
