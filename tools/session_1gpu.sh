#!/bin/bash
# One-GPU check-out: build, kernel numerics, GEMM / attention micro-benchmarks, the 1-GPU benchmark with a kernel timeline.
# Usage (from the repo root):  gpurun --timeout 900 -- 'bash tools/session_1gpu.sh'
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== gpu tests"; timeout -s KILL 400 python -m pytest tests -q -x -m gpu --timeout 120 --tb=short 2>&1 | tail -8 | cut -c1-300 | tee gpurun_out/pytest_gpu.log
echo "== gemm bench"; timeout -s KILL 300 python tools/gemm_bench.py 8192 2>&1 | tail -48 | tee gpurun_out/gemm_bench.log
echo "== attention bench"; timeout -s KILL 120 python tools/attn_bench.py 2>&1 | tail -3 | tee gpurun_out/attn_bench.log
echo "== bench 1 gpu"; timeout -s KILL 300 python bench.py --steps 6 --warmup 3 --profile gpurun_out/step_profile_1gpu.txt 2>&1 | tail -1 | tee gpurun_out/bench_1gpu.log | cut -c1-600
head -24 gpurun_out/step_profile_1gpu.txt | cut -c1-160
