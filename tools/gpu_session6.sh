#!/bin/bash
# 1-GPU profiling session: attention timing + ncu --set full on the three top kernels
mkdir -p gpurun_out
export EPL_ATTENTION=epl
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== attn bench"; timeout -s KILL 200 python tools/attn_bench.py 2>&1 | tail -5 | tee gpurun_out/attn_bench.log
echo "== ncu attention"; timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel\|attn_bwd_kernel -s 4 -c 2 -o gpurun_out/prof_attn -f python tools/attn_bench.py quick > gpurun_out/ncu_attn.log 2>&1; tail -2 gpurun_out/ncu_attn.log
echo "== ncu gemm"; timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 20 -c 3 -o gpurun_out/prof_gemm -f python tools/gemm_bench.py 8192 > gpurun_out/ncu_gemm.log 2>&1; tail -2 gpurun_out/ncu_gemm.log
ls -la gpurun_out/*.ncu-rep
