#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
echo "== mgpu_check fused"; timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/mgpu_check.py fused > gpurun_out/mgpu_check_fused.log 2>&1; grep -v "^W0\|OMP_NUM\|^\*\*\*" gpurun_out/mgpu_check_fused.log | tail -12 | cut -c1-400
echo "== pp2 debug (small)"; EPL_PIPE_DEBUG=1 EPL_HANG_DUMP=60 timeout -s KILL 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 3 --warmup 3 --parallelism pp2 --model small --batch 2 --seq 256 --micro-batches 4 --no-e2e > gpurun_out/pp2_debug.log 2>&1; grep -v "^W0\|OMP_NUM\|^\*\*\*\|site-packages" gpurun_out/pp2_debug.log | tail -6 | cut -c1-600
echo "== bench pp2 xl"; EPL_HANG_DUMP=150 timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 4 --warmup 3 --parallelism pp2 --batch 2 --no-e2e 2>&1 | grep -v "^W0\|OMP_NUM" | tail -2 | tee gpurun_out/bench_pp2.log | cut -c1-700
echo "== bench dp2"; timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 6 --warmup 3 2>&1 | grep -v "^W0\|OMP_NUM" | tail -1 | tee gpurun_out/bench_dp2_v5.log | cut -c1-1600
