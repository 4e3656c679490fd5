#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== sustained gemm"; timeout -s KILL 200 python tools/gemm_sustained.py 2>&1 | tail -14 | tee gpurun_out/gemm_sustained.log
echo "== bench 1gpu + kernel timeline"; timeout -s KILL 300 python bench.py --steps 6 --warmup 3 --no-e2e --profile gpurun_out/step_profile_1gpu.txt 2>&1 | tail -1 | tee gpurun_out/bench_1gpu_v5.log | cut -c1-300
tail -52 gpurun_out/step_profile_1gpu.txt
echo "== bench 1gpu with cuBLAS linears (A/B)"; EPL_LINEAR=torch timeout -s KILL 300 python bench.py --steps 6 --warmup 3 --no-e2e --profile gpurun_out/step_profile_1gpu_cublas.txt 2>&1 | tail -1 | tee gpurun_out/bench_1gpu_v5_cublas.log | cut -c1-300
head -12 gpurun_out/step_profile_1gpu_cublas.txt | cut -c1-200
