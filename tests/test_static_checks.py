"""Static checks that stand in for the GPU on a CPU-only box: code that only runs on a B200 (fused data parallel, CUDA-graph
capture, stage graphs, NVLS, the benchmark driver) must at least refer to names and native functions that exist."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_global_name_and_native_symbol_resolves():
  r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "check_globals.py")], capture_output=True, text=True, cwd=ROOT)
  assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
  assert "0 unresolved names" in r.stdout


def test_bench_and_entry_point_compile():
  import py_compile
  for f in ("bench.py", "__graft_entry__.py", "tools/mgpu_check.py", "tools/gemm_bench.py", "tools/sass_mnemonics.py"):
    py_compile.compile(os.path.join(ROOT, f), doraise=True)
