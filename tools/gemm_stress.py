"""Randomised stress of the tcgen05 GEMM (default dispatch -> 2-CTA kernel for M,N >= 256) against torch; reports any
launch error or mismatch with its shape.  Looks for intermittent failures, so every shape runs several times."""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from easyparallellibrary_b200.ops import linear as L

random.seed(0); torch.manual_seed(0)
bad = 0
n = int(sys.argv[1]) if len(sys.argv) > 1 else 150
L._FORCE_BN = int(sys.argv[2]) if len(sys.argv) > 2 else 0      # e.g. 1024 = 4-CTA multicast kernel
for it in range(n):
  M = random.choice([256, 384, 512, 1000, 1024, 2048, 8192])
  N = random.choice([256, 768, 1024, 1600, 4800])
  K = random.choice([64, 256, 384, 1024, 1600, 6400])
  a_mn, b_mn = random.random() < 0.3, random.random() < 0.3
  a = (torch.randn((K, M) if a_mn else (M, K), device="cuda") * 0.1).bfloat16()
  b = (torch.randn((K, N) if b_mn else (N, K), device="cuda") * 0.1).bfloat16()
  ref = (a.float().t() if a_mn else a.float()) @ (b.float() if b_mn else b.float().t())
  for rep in range(3):
    try:
      d = L.gemm(a, b, a_mn_major=a_mn, b_mn_major=b_mn)
      torch.cuda.synchronize()
    except Exception as e:  # noqa
      print("ERROR it=%d rep=%d M=%d N=%d K=%d a_mn=%d b_mn=%d: %s" % (it, rep, M, N, K, a_mn, b_mn, e), flush=True)
      bad += 1
      break
    err = (d.float() - ref).abs().max().item()
    if err > 0.05 * max(ref.abs().max().item(), 1.0):
      print("MISMATCH it=%d rep=%d M=%d N=%d K=%d a_mn=%d b_mn=%d err=%.4f" % (it, rep, M, N, K, a_mn, b_mn, err), flush=True)
      bad += 1
print("gemm stress: %d shapes x3, %d failures" % (n, bad))
