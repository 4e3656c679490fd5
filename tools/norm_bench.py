"""LayerNorm forward / backward timing on the GPT-2-XL activation shape, vs torch and vs the HBM roofline."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from easyparallellibrary_b200.ops.layernorm import layer_norm


def timeit(fn, iters=20, warm=5, flush=True):
  fl = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
  for _ in range(warm):
    fn()
  torch.cuda.synchronize()
  ts = []
  for _ in range(iters):
    if flush:
      fl.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
  return sorted(ts)[len(ts) // 2] * 1e3


rows, D = 8192, 1600
x = torch.randn(rows, D, device="cuda").bfloat16().requires_grad_()
g = torch.randn(D, device="cuda").bfloat16().requires_grad_()
b = torch.randn(D, device="cuda").bfloat16().requires_grad_()
dy = torch.randn(rows, D, device="cuda").bfloat16()
mb = rows * D * 2 / 1e6
for flush in (True, False):
  t_f = timeit(lambda: layer_norm(x, g, b, 1e-5), flush=flush)
  y = layer_norm(x, g, b, 1e-5)
  def bw():
    y.backward(dy, retain_graph=True); x.grad = None; g.grad = None; b.grad = None
  t_b = timeit(bw, flush=flush)
  t_tf = timeit(lambda: torch.nn.functional.layer_norm(x, (D,), g, b, 1e-5), flush=flush)
  yt = torch.nn.functional.layer_norm(x, (D,), g, b, 1e-5)
  def bwt():
    yt.backward(dy, retain_graph=True); x.grad = None; g.grad = None; b.grad = None
  t_tb = timeit(bwt, flush=flush)
  print("L2 %s | epl fwd %.1f us (%.2f TB/s) bwd %.1f us | torch fwd %.1f us bwd %.1f us | roofline fwd %.1f us bwd(dx) %.1f us @6.4TB/s" % (
      "flushed" if flush else "warm   ", t_f, 2 * mb / t_f, t_b, t_tf, t_tb, 2 * mb / 6.4, 3 * mb / 6.4), flush=True)
