#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
echo "== attention + gemm tests"; EPL_ATTENTION=epl CUDA_VISIBLE_DEVICES=0 timeout -s KILL 200 python -m pytest tests/test_kernels_gpu.py -q -x -k "flash_attention or gemm_layouts or linear_and_mlp or gpt2" --timeout 60 2>&1 | tail -6 | tee gpurun_out/pytest_s8.log
echo "== attn bench"; CUDA_VISIBLE_DEVICES=0 timeout -s KILL 120 python tools/attn_bench.py quick 2>&1 | tail -2 | tee gpurun_out/attn_bench_v3.log
echo "== pp2 debug (small)"; EPL_HANG_DUMP=50 timeout -s KILL 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 3 --warmup 3 --parallelism pp2 --model small --batch 2 --seq 256 --micro-batches 4 --no-e2e > gpurun_out/pp2_debug.log 2>&1; grep -v "^W0\|OMP_NUM\|^\*\*\*\|site-packages" gpurun_out/pp2_debug.log | tail -12 | cut -c1-500
echo "== mgpu check tp"; timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py tp > gpurun_out/mgpu_check_tp.log 2>&1; grep -v "^W0\|OMP_NUM\|frame #\|^  File\|^    \|^\*\*\*" gpurun_out/mgpu_check_tp.log | head -30 | cut -c1-300
echo "== bench pp2 xl"; EPL_HANG_DUMP=150 timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 4 --warmup 3 --parallelism pp2 --batch 2 --no-e2e 2>&1 | grep -v "^W0\|OMP_NUM" | tail -1 | tee gpurun_out/bench_pp2.log | cut -c1-700
echo "== bench dp2"; timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 8 --warmup 3 --no-e2e 2>&1 | grep -v "^W0\|OMP_NUM" | tail -1 | tee gpurun_out/bench_dp${N}_v3.log | cut -c1-700
