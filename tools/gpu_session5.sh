#!/bin/bash
# 2-GPU: pipeline debug (short timeouts), TP fused kernels, dp2 bench with overlap
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
echo "== pp2 debug (small)"; EPL_PIPE_DEBUG=1 EPL_HANG_DUMP=60 timeout -s KILL 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 2 --warmup 3 --parallelism pp2 --model small --batch 2 --seq 256 --micro-batches 4 --no-e2e > gpurun_out/pp2_debug.log 2>&1; grep -v "^W0\|OMP_NUM\|^\*\*\*" gpurun_out/pp2_debug.log | tail -60 | cut -c1-400
for what in native tp; do
echo "== mgpu check $what"; timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py $what > gpurun_out/mgpu_check_$what.log 2>&1; grep -v "^W0\|OMP_NUM\|frame #\|^  File\|^    \|^\*\*\*" gpurun_out/mgpu_check_$what.log | head -40 | cut -c1-300
done
echo "== bench dp fused+overlap"; timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 8 --warmup 3 2>&1 | grep -v "^W0\|OMP_NUM" | tail -1 | tee gpurun_out/bench_dp${N}_fused_v2.log | cut -c1-900
