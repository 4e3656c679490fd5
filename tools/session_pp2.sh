#!/bin/bash
# 2-GPU pipeline session: pp2 with micro-batches of 2 sequences (8 layers, then the full model) + GEMM kernel tests.
mkdir -p gpurun_out; N=2
run() { timeout -s KILL $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
F='^W0\|OMP_NUM\|^\*\*\*'
echo "== pp2 micro-batch 2, 8 layers"; run 100 29531 bench.py --gpus 2 --steps 2 --warmup 3 --no-e2e --parallelism pp2 --batch 2 --layers 8 > gpurun_out/bench_pp2_mb2_l8.log 2>&1; grep -v "$F" gpurun_out/bench_pp2_mb2_l8.log | grep "epl\]\|Error\|metric" | tail -12 | cut -c1-400
echo "== pp2 micro-batch 2"; run 150 29533 bench.py --gpus 2 --steps 4 --warmup 3 --no-e2e --parallelism pp2 --batch 2 > gpurun_out/bench_pp2_mb2.log 2>&1; grep -v "$F" gpurun_out/bench_pp2_mb2.log | grep "epl\]\|Error\|metric" | tail -12 | cut -c1-400
echo "== gemm tests"; CUDA_VISIBLE_DEVICES=0 timeout 200 python -m pytest tests/test_kernels_gpu.py -x -q -k "gemm or linear or fp8" 2>&1 | tail -3
