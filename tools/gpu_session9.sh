#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== failing test (full)"; timeout -s KILL 200 python -m pytest tests/test_kernels_gpu.py -q -x -k "linear_and_mlp" --timeout 60 --tb=short 2>&1 | grep -v "^$" | tail -30 | cut -c1-300 | tee gpurun_out/pytest_s9_mlp.log
echo "== attention tests"; EPL_ATTENTION=epl timeout -s KILL 200 python -m pytest tests/test_kernels_gpu.py -q -x -k "flash_attention or gpt2" --timeout 60 --tb=short 2>&1 | tail -15 | cut -c1-300 | tee gpurun_out/pytest_s9.log
echo "== attn bench"; timeout -s KILL 120 python tools/attn_bench.py 2>&1 | tail -3 | tee gpurun_out/attn_bench_v4.log
echo "== ncu attention"; timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel\|attn_bwd_kernel -c 2 -o gpurun_out/ncu_attn_v4 -f python tools/attn_bench.py quick > gpurun_out/ncu_attn_v4.log 2>&1; tail -2 gpurun_out/ncu_attn_v4.log
echo "== all gpu tests"; timeout -s KILL 400 python -m pytest tests -q -x -m gpu --timeout 120 --tb=short 2>&1 | tail -8 | cut -c1-300 | tee gpurun_out/pytest_s9_all.log
echo "== bench 1gpu own attention"; EPL_ATTENTION=epl timeout -s KILL 300 python bench.py --steps 6 --warmup 3 --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_1gpu_v4_eplattn.log | cut -c1-420
echo "== bench 1gpu sdpa"; EPL_ATTENTION=sdpa timeout -s KILL 300 python bench.py --steps 6 --warmup 3 --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_1gpu_v4_sdpa.log | cut -c1-420
