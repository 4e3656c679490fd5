#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
python - <<'PY'
import ctypes
from easyparallellibrary_b200.ops import _lib
lib=_lib.require(); lib.epl_gemm4_max_clusters.restype=ctypes.c_int
print("co-resident 4-CTA clusters:", lib.epl_gemm4_max_clusters())
PY
echo "== gemm tests"; timeout -s KILL 300 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 60 --tb=short -k "two_cta" 2>&1 | tail -8 | cut -c1-300 | tee gpurun_out/pytest_s18.log
echo "== gemm stress 4cta"; timeout -s KILL 200 python tools/gemm_stress.py 80 1024 2>&1 | tail -6 | tee gpurun_out/gemm_stress4.log
echo "== gemm bench"; timeout -s KILL 400 python tools/gemm_bench.py 8192 2>&1 | tail -64 | tee gpurun_out/gemm_bench_v4.log
