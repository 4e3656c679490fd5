#!/bin/bash
# First GPU visit: numerics tests, GEMM throughput, 1-GPU bench, launch list + ncu capture of the GEMM.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== pytest gpu"; timeout -s KILL 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== gemm bench"; timeout -s KILL 600 python tools/gemm_bench.py 8192 2>&1 | tail -80 | tee gpurun_out/gemm_bench.log
echo "== bench"; timeout -s KILL 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -5 | tee gpurun_out/bench1.log
