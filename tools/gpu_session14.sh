#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== kernel tests"; timeout -s KILL 300 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 60 --tb=short 2>&1 | tail -12 | cut -c1-300 | tee gpurun_out/pytest_s14.log
echo "== bench 1gpu + kernel timeline"; timeout -s KILL 300 python bench.py --steps 6 --warmup 3 --no-e2e --profile gpurun_out/step_profile_1gpu_v7.txt 2>&1 | tail -1 | tee gpurun_out/bench_1gpu_v7.log | cut -c1-300
head -22 gpurun_out/step_profile_1gpu_v7.txt | cut -c1-180
