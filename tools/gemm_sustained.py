"""Sustained (power-capped) GEMM throughput: each variant runs back to back for ~2.5 s, timed over the last 1.5 s,
with nvidia-smi clock / power samples.  The burst numbers of tools/gemm_bench.py are taken at ~1.9 GHz; a training step
runs at the ~1000 W cap where the SM clock drops to ~1.35 GHz, so efficiency per joule decides the in-model time."""
import os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from easyparallellibrary_b200.ops import linear as L


class Smi(threading.Thread):
  def __init__(self):
    super().__init__(daemon=True); self.rows = []; self.halt = threading.Event()
  def run(self):
    while not self.halt.is_set():
      try:
        o = subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits"],
                           capture_output=True, text=True, timeout=5).stdout.strip().split(",")
        self.rows.append((float(o[0]), float(o[1])))
      except Exception:
        pass
      self.halt.wait(0.1)
  def stop(self):
    self.halt.set(); self.join(timeout=3)
    r = self.rows[len(self.rows) // 2:]
    return (sorted(x[0] for x in r)[len(r) // 2], sorted(x[1] for x in r)[len(r) // 2]) if r else (0, 0)


def sustained(fn, flops, secs=2.5):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  smi = Smi(); smi.start()
  t_end = time.time() + secs
  n = 0
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  started = False
  while time.time() < t_end:
    if not started and time.time() > t_end - 1.5:
      a.record(); started = True; n = 0
    for _ in range(20):
      fn()
    n += 20
    torch.cuda.synchronize()
  b.record(); torch.cuda.synchronize()
  ms = a.elapsed_time(b)
  clk, pw = smi.stop()
  return flops * n / ms / 1e9, clk, pw


def main():
  T, d = 8192, 1600
  for name, M, N, K in (("fc1 fwd", T, 4 * d, d), ("fc2 fwd", T, d, 4 * d), ("qkv fwd", T, 3 * d, d)):
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    fl = 2.0 * M * N * K
    for label, bn, fn in (("epl 2cta", 0, lambda: L.gemm(x, w)), ("epl 1cta", 256, lambda: L.gemm(x, w)), ("cublas", 0, lambda: x @ w.t())):
      L._FORCE_BN = bn
      tf, clk, pw = sustained(fn, fl)
      L._FORCE_BN = 0
      print("%-8s %-9s sustained %7.1f TFLOP/s  sm %4.0f MHz  %4.0f W  (%.2f TFLOP/J)" % (name, label, tf, clk, pw, tf / max(pw, 1)), flush=True)
    time.sleep(1.0)
  # dW with fp32 accumulate into a main-grad buffer (the in-model weight-gradient GEMM)
  M, N, K = 4 * d, d, T
  dy = torch.randn(T, 4 * d, device="cuda").bfloat16(); x = torch.randn(T, d, device="cuda").bfloat16()
  acc = torch.zeros(4 * d, d, device="cuda", dtype=torch.float32)
  fl = 2.0 * M * N * K
  for label, fn in (("epl fp32 accumulate", lambda: L.gemm(dy, x, a_mn_major=True, b_mn_major=True, out=acc, accumulate=True)),
                    ("epl bf16 out", lambda: L.gemm(dy, x, a_mn_major=True, b_mn_major=True)),
                    ("cublas bf16 out", lambda: dy.t() @ x)):
    tf, clk, pw = sustained(fn, fl)
    print("fc1 dW   %-20s sustained %7.1f TFLOP/s  sm %4.0f MHz  %4.0f W" % (label, tf, clk, pw), flush=True)


if __name__ == "__main__":
  main()
