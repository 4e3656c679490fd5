"""GEMM throughput of the tcgen05 kernel vs torch.matmul (cuBLAS) on GPT-2-XL training shapes."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from easyparallellibrary_b200.ops import linear as L


def timeit(fn, iters=20, warm=5):
  flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
  for _ in range(warm):
    fn()
  torch.cuda.synchronize()
  ts = []
  for _ in range(iters):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
  ts.sort()
  return ts[len(ts) // 2]


def main():
  T = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
  d = 1600
  shapes = [("qkv", T, 3 * d, d), ("proj", T, d, d), ("fc1", T, 4 * d, d), ("fc2", T, d, 4 * d), ("lm_head", T, 50304, d)]
  rows = []
  for name, M, N, K in shapes:
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    dy = torch.randn(M, N, device="cuda").bfloat16()
    fl = 2.0 * M * N * K
    for layout, mine, ref in (
        ("nt fwd", lambda: L.gemm(x, w), lambda: x @ w.t()),
        ("nn dX", lambda: L.gemm(dy, w, b_mn_major=True), lambda: dy @ w),
        ("tn dW", lambda: L.gemm(dy, x, a_mn_major=True, b_mn_major=True), lambda: dy.t() @ x)):
      for bn in (0, 512):
        if bn == 160 and layout != "nt fwd":
          continue
        L._FORCE_BN = bn
        try:
          t = timeit(mine)
        except Exception as e:
          t = float("nan")
        rows.append((name, layout, bn, t, fl / t / 1e9 if t == t else 0))
      L._FORCE_BN = 0
      t = timeit(ref)
      rows.append((name, layout, "cublas", t, fl / t / 1e9))
  # fp8 (e4m3) forward GEMMs: kernel only (operands pre-quantised) and including both per-tensor quantisation passes
  from easyparallellibrary_b200.ops import fp8, _lib
  lib = _lib.require()
  for name, M, N, K in shapes:
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    xq, sx = fp8.quantize_e4m3(x)
    wq, sw = fp8.quantize_e4m3(w)
    out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    fl = 2.0 * M * N * K

    def kern():
      rc = lib.epl_gemm_fp8(xq.data_ptr(), wq.data_ptr(), out.data_ptr(), M, N, K, K, K, N, None, None, None, 0, _lib.dtype_code(out.dtype),
                            1.0, sx.data_ptr(), sw.data_ptr(), 0, _lib.stream())
      assert rc == 0
    t = timeit(kern)
    rows.append((name, "nt fwd", "fp8", t, fl / t / 1e9))
    t = timeit(lambda: fp8.gemm_fp8(x, w, b_q=(wq, sw)))
    rows.append((name, "nt fwd", "fp8+qx", t, fl / t / 1e9))
  for r in rows:
    print("%-8s %-7s bn=%-6s %8.3f ms %8.1f TFLOP/s" % r)
  json.dump(rows, open("gpurun_out/gemm_bench.json", "w"))


if __name__ == "__main__":
  os.makedirs("gpurun_out", exist_ok=True)
  main()
