#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== new tests"; EPL_ATTENTION=epl timeout -s KILL 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "two_cta or tile_widths or flash_attention or vocab_parallel or layernorm" --timeout 60 2>&1 | tail -12 | tee gpurun_out/pytest_s7.log
echo "== attn bench"; timeout -s KILL 200 python tools/attn_bench.py 2>&1 | tail -4 | tee gpurun_out/attn_bench_v2.log
echo "== gemm bench"; timeout -s KILL 300 python tools/gemm_bench.py 8192 2>&1 | tail -50 | tee gpurun_out/gemm_bench_v2.log
echo "== bench 1gpu own attention"; EPL_ATTENTION=epl timeout -s KILL 300 python bench.py --steps 6 --warmup 3 --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_1gpu_v3_eplattn.log | cut -c1-420
echo "== bench 1gpu sdpa"; EPL_ATTENTION=sdpa timeout -s KILL 300 python bench.py --steps 6 --warmup 3 --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_1gpu_v3_sdpa.log | cut -c1-420
