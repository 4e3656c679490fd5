#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== gemm tests"; timeout -s KILL 300 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 60 --tb=short -k "gemm or linear or mlp or block" 2>&1 | tail -12 | cut -c1-300 | tee gpurun_out/pytest_s17.log
echo "== gemm bench"; timeout -s KILL 400 python tools/gemm_bench.py 8192 2>&1 | tail -96 | tee gpurun_out/gemm_bench_v3.log | grep -v "bn=256 " 
echo "== bench 1gpu + kernel timeline"; timeout -s KILL 300 python bench.py --steps 6 --warmup 3 --no-e2e --profile gpurun_out/step_profile_1gpu_v8.txt 2>&1 | tail -1 | tee gpurun_out/bench_1gpu_v8.log | cut -c1-300
head -4 gpurun_out/step_profile_1gpu_v8.txt | cut -c1-180; grep -A4 "GEMM kernel durations" gpurun_out/step_profile_1gpu_v8.txt; tail -4 gpurun_out/step_profile_1gpu_v8.txt
