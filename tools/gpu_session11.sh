#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== bench 1gpu + kernel timeline"; timeout -s KILL 400 python bench.py --steps 6 --warmup 3 --no-e2e --profile gpurun_out/step_profile_1gpu.txt 2>&1 | tail -1 | tee gpurun_out/bench_1gpu_v5.log | cut -c1-300
head -45 gpurun_out/step_profile_1gpu.txt
