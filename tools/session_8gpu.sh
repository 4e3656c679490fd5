#!/bin/bash
# 8-GPU evidence run (charged 8x: every item has a tight timeout).  Usage: gpurun --gpus 8 --timeout 300 -- 'bash tools/session_8gpu.sh'
mkdir -p gpurun_out; N=8
run() { timeout -s KILL $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
F='^W0\|OMP_NUM\|^\*\*\*'
echo "== dp8 K1 after backward"; EPL_FUSED_OVERLAP=0 run 80 29514 bench.py --gpus 8 --steps 6 --warmup 3 --no-e2e 2>&1 | grep -v "$F" | tail -1 | tee gpurun_out/bench_dp8_r2_ov0.log | cut -c1-500
echo "== dp8 default (K1 overlapped with backward from 8 ranks)"; run 80 29513 bench.py --gpus 8 --steps 6 --warmup 3 --no-e2e 2>&1 | grep -v "$F" | tail -1 | tee gpurun_out/bench_dp8_r2_default.log | cut -c1-2300
echo "== pp2 x dp4, micro-batches of 8"; run 90 29517 bench.py --gpus 8 --steps 3 --warmup 3 --no-e2e --parallelism pp2 --batch 8 2>&1 | grep -v "$F" | tail -1 | tee gpurun_out/bench_pp2dp4_r2.log | cut -c1-700
echo "== multi-GPU checks at 8 ranks"; EPL_CUDA_GRAPH=1 run 120 29519 tools/mgpu_check.py all > gpurun_out/mgpu_check_w8.log 2>&1; grep "CHECK\|Error" gpurun_out/mgpu_check_w8.log | sort | uniq -c | head -30
echo "== dp8 library arm"; run 80 29515 bench.py --gpus 8 --steps 6 --warmup 3 --no-e2e --impl baseline 2>&1 | grep -v "$F" | tail -1 | tee gpurun_out/bench_dp8_r2_lib.log | cut -c1-500
