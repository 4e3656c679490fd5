#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== norm bench"; timeout -s KILL 120 python tools/norm_bench.py 2>&1 | tail -4 | tee gpurun_out/norm_bench.log
echo "== ncu norm"; timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:norm_fwd_kernel\|norm_bwd_dx_kernel -s 6 -c 2 -o gpurun_out/ncu_norm -f python tools/norm_bench.py > gpurun_out/ncu_norm.log 2>&1; tail -2 gpurun_out/ncu_norm.log
