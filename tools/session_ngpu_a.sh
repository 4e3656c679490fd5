#!/bin/bash
# 8-GPU session A: the headline DP8 number (CUDA graph + K1 overlapped), its A/B arms, K1 micro-benchmark, pp2 x dp4 with a timeline.
mkdir -p gpurun_out; N=${N:-8}
run() { timeout -s KILL $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
F='^W0\|OMP_NUM\|^\*\*\*'
nproc; cat /proc/loadavg
echo "== dp$N default (graph + K1 overlap)"; run 150 29513 bench.py --gpus $N --steps 10 --warmup 4 2>&1 | grep -v "$F" | tail -1 | tee gpurun_out/bench_dp${N}_r2_default.log | cut -c1-2200
echo "== dp$N library arm"; run 150 29515 bench.py --gpus $N --steps 10 --warmup 4 --no-e2e --impl baseline 2>&1 | grep -v "$F" | tail -1 | tee gpurun_out/bench_dp${N}_r2_lib.log | cut -c1-500
echo "== dp$N overlap off"; EPL_FUSED_OVERLAP=0 run 150 29514 bench.py --gpus $N --steps 10 --warmup 4 --no-e2e 2>&1 | grep -v "$F" | tail -1 | tee gpurun_out/bench_dp${N}_r2_ov0.log | cut -c1-500
echo "== dp$N eager (no graph)"; run 150 29516 bench.py --gpus $N --steps 10 --warmup 4 --no-e2e --no-graph 2>&1 | grep -v "$F" | tail -1 | tee gpurun_out/bench_dp${N}_r2_eager.log | cut -c1-500
echo "== k1 exact + bench W=$N"; run 100 29511 tools/mgpu_check.py k1 k1bench > gpurun_out/mgpu_k1_w${N}.log 2>&1; grep "k1 \|PASSED" gpurun_out/mgpu_k1_w${N}.log | cut -c1-230 | tail -26
echo "== pp2 x dp4 (timeline of rank 0)"; run 200 29517 bench.py --gpus $N --steps 4 --warmup 3 --no-e2e --parallelism pp2 --profile gpurun_out/step_profile_pp2dpx.txt > gpurun_out/bench_pp2dpx_r2.log 2>&1; grep -v "$F" gpurun_out/bench_pp2dpx_r2.log | tail -2 | cut -c1-600; head -30 gpurun_out/step_profile_pp2dpx.txt | cut -c1-170
