#!/bin/bash
# 8-GPU evidence run (charged 8x: every item has a tight timeout).  Usage: gpurun --gpus 8 --timeout 420 -- 'bash tools/session_8gpu.sh'
mkdir -p gpurun_out; N=8
run() { timeout -s KILL $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
F='^W0\|OMP_NUM\|^\*\*\*'
echo "== dp8 default (CUDA graph, K1 overlapped with backward)"; run 110 29513 bench.py --gpus 8 --steps 8 --warmup 4 2>&1 | grep -v "$F" | tail -1 | tee gpurun_out/bench_dp8_r2_default.log | cut -c1-2300
echo "== dp8 K1 after backward"; EPL_FUSED_OVERLAP=0 run 100 29514 bench.py --gpus 8 --steps 8 --warmup 4 --no-e2e 2>&1 | grep -v "$F" | tail -1 | tee gpurun_out/bench_dp8_r2_ov0.log | cut -c1-500
echo "== dp8 library arm"; run 100 29515 bench.py --gpus 8 --steps 8 --warmup 4 --no-e2e --impl baseline 2>&1 | grep -v "$F" | tail -1 | tee gpurun_out/bench_dp8_r2_lib.log | cut -c1-500
echo "== pp2 x dp4"; run 110 29517 bench.py --gpus 8 --steps 4 --warmup 3 --no-e2e --parallelism pp2 2>&1 | grep -v "$F" | tail -1 | tee gpurun_out/bench_pp2dp4_r2.log | cut -c1-600
echo "== bert-large split(8)"; run 100 29518 bench.py --gpus 8 --steps 6 --warmup 4 --no-e2e --workload bert --parallelism tp8 2>&1 | grep -v "$F" | tail -1 | tee gpurun_out/bench_bert_tp8.log | cut -c1-600
