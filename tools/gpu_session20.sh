#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
echo "== bench dp$N"; timeout -s KILL 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 6 --warmup 3 2>&1 | grep -v "^W0\|OMP_NUM" | tail -1 | tee gpurun_out/bench_dp${N}_v5.log | cut -c1-1700
echo "== mgpu_check fused ($N ranks)"; timeout -s KILL 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/mgpu_check.py fused > gpurun_out/mgpu_check_fused_$N.log 2>&1; grep "fused vs\|PASSED\|Error" gpurun_out/mgpu_check_fused_$N.log | head -5 | cut -c1-300
echo "== bench pp2 x dp$((N/2))"; EPL_HANG_DUMP=100 timeout -s KILL 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --steps 4 --warmup 3 --parallelism pp2 --batch 4 --micro-batches 4 --no-e2e 2>&1 | grep -v "^W0\|OMP_NUM" | tail -1 | tee gpurun_out/bench_pp2_dp$((N/2)).log | cut -c1-900
echo "== bench baseline dp$N (NCCL all-reduce + unfused Adam)"; timeout -s KILL 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29516 bench.py --impl baseline --gpus $N --steps 6 --warmup 3 --no-e2e 2>&1 | grep -v "^W0\|OMP_NUM" | tail -1 | tee gpurun_out/bench_dp${N}_baseline.log | cut -c1-500
