#!/bin/bash
# Multi-GPU check-out on N = all visible GPUs: communicator / symmetric-memory / fused-kernel parity, pipeline, DP benchmark.
# Usage:  gpurun --gpus 2 --timeout 900 -- 'bash tools/session_multigpu.sh'      (N = 2, 4 or 8; charged N x)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
run() { timeout -s KILL $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $3 "${@:4}"; }
echo "== mgpu_check"; run 300 $N 29511 tools/mgpu_check.py native symm fused tp tptrain moe > gpurun_out/mgpu_check.log 2>&1; grep -v "^W0\|OMP_NUM\|frame #\|^\*\*\*" gpurun_out/mgpu_check.log | tail -30 | cut -c1-300
echo "== pipeline 2 stages x $((N/2)) replicas"; run 250 $N 29515 bench.py --gpus $N --steps 4 --warmup 3 --parallelism pp2 --batch 4 --micro-batches 4 --no-e2e 2>&1 | grep -v "^W0\|OMP_NUM" | tail -12 | tee gpurun_out/bench_pp2.log | cut -c1-700
echo "== bench dp$N"; run 200 $N 29512 bench.py --gpus $N --steps 6 --warmup 3 2>&1 | grep -v "^W0\|OMP_NUM" | tail -1 | tee gpurun_out/bench_dp$N.log | cut -c1-1700
echo "== bench dp$N, reference algorithm on library calls"; run 200 $N 29516 bench.py --impl baseline --gpus $N --steps 6 --warmup 3 --no-e2e 2>&1 | grep -v "^W0\|OMP_NUM" | tail -1 | tee gpurun_out/bench_dp${N}_baseline.log | cut -c1-500
