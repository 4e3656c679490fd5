#!/bin/bash
# 2-GPU session: native comm, symmetric memory, fused DP kernel parity, then dp2 bench fused vs baseline
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
echo "== mgpu check"; timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py 2>&1 | grep -v "^W0\|OMP_NUM" | tail -30 | tee gpurun_out/mgpu_check.log
echo "== bench dp fused"; timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 8 --warmup 3 2>&1 | grep -v "^W0\|OMP_NUM" | tail -3 | tee gpurun_out/bench_dp${N}_fused.log
echo "== bench dp baseline"; timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 8 --warmup 3 --impl baseline 2>&1 | grep -v "^W0\|OMP_NUM" | tail -3 | tee gpurun_out/bench_dp${N}_baseline.log
