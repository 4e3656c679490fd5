#!/bin/bash
# 2-GPU session: native comm, symmetric memory, fused DP kernel parity, then dp2 bench fused vs baseline
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
for what in native symm fused; do
echo "== mgpu check $what"; timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py $what > gpurun_out/mgpu_check_$what.log 2>&1; grep -v "^W0\|OMP_NUM\|frame #\|^  File\|^    " gpurun_out/mgpu_check_$what.log | head -40
done
echo "== pytest new"; timeout -s KILL 300 python -m pytest tests/test_kernels_gpu.py -q -k "fork or layernorm or linear_and_mlp or gpt2" --timeout 200 2>&1 | tail -5
echo "== bench 1gpu"; CUDA_VISIBLE_DEVICES=0 timeout -s KILL 600 python bench.py --steps 6 --warmup 3 --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_1gpu_v2.log
echo "== bench pp2"; timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 4 --warmup 3 --parallelism pp2 --batch 2 --no-e2e 2>&1 | grep -v "^W0\|OMP_NUM" | tail -3 | tee gpurun_out/bench_pp2.log
echo "== bench dp fused"; timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 8 --warmup 3 2>&1 | grep -v "^W0\|OMP_NUM" | tail -3 | tee gpurun_out/bench_dp${N}_fused.log
echo "== bench dp baseline"; timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 8 --warmup 3 --impl baseline 2>&1 | grep -v "^W0\|OMP_NUM" | tail -3 | tee gpurun_out/bench_dp${N}_baseline.log
