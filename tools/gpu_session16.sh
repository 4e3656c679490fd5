#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
echo "== pp2 debug (small)"; EPL_HANG_DUMP=50 timeout -s KILL 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 3 --warmup 3 --parallelism pp2 --model small --batch 2 --seq 256 --micro-batches 4 --no-e2e > gpurun_out/pp2_debug.log 2>&1; grep -v "^W0\|OMP_NUM\|^\*\*\*\|site-packages" gpurun_out/pp2_debug.log | tail -8 | cut -c1-400
echo "== bench pp2 xl"; EPL_HANG_DUMP=150 timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 4 --warmup 3 --parallelism pp2 --batch 2 --no-e2e 2>&1 | grep -v "^W0\|OMP_NUM" | tail -3 | tee gpurun_out/bench_pp2.log | cut -c1-700
echo "== multi-gpu tests"; timeout -s KILL 300 python -m pytest tests/test_multigpu.py -q -x --timeout 250 --tb=short 2>&1 | tail -6 | cut -c1-300 | tee gpurun_out/pytest_mgpu.log
echo "== bench dp"; timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 8 --warmup 3 2>&1 | grep -v "^W0\|OMP_NUM" | tail -1 | tee gpurun_out/bench_dp${N}_v4.log | cut -c1-1500
