#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== attn trace"; timeout -s KILL 120 python tools/attn_trace.py 2>&1 | tail -22 | tee gpurun_out/attn_trace_v4.log
echo "== gemm stress"; timeout -s KILL 300 python tools/gemm_stress.py 200 2>&1 | tail -12 | tee gpurun_out/gemm_stress.log
echo "== kernel tests x3 (flakiness)"; for i in 1 2 3; do timeout -s KILL 200 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 60 --tb=short 2>&1 | tail -4 | cut -c1-400; done | tee gpurun_out/pytest_s10.log
