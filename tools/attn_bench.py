"""Flash-attention kernel timing vs cuDNN SDPA (through torch) — CUDA events, L2 flushed between iterations."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from easyparallellibrary_b200.ops.attention_kernel import flash_attention_packed


def timeit(fn, iters=10, warm=3):
  flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
  for _ in range(warm):
    fn()
  torch.cuda.synchronize()
  ts = []
  for _ in range(iters):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
  return sorted(ts)[len(ts) // 2]


def main():
  quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
  shapes = [(8, 1024, 25), (8, 128, 25), (2, 4096, 16)] if not quick else [(8, 1024, 25)]
  for B, S, H in shapes:
    D = 64
    qkv = (torch.randn(B, S, 3, H, D, device="cuda") * 0.5).bfloat16().requires_grad_()
    dout = torch.randn(B, S, H * D, device="cuda").bfloat16()
    out = flash_attention_packed(qkv, True)
    fl = 4.0 * B * H * S * S * D / 2
    t_f = timeit(lambda: flash_attention_packed(qkv, True))
    def fb():
      o = flash_attention_packed(qkv, True); o.backward(dout); qkv.grad = None
    t_fb = timeit(fb)
    q, k, v = qkv.detach().permute(2, 0, 3, 1, 4).unbind(0)
    q, k, v = q.contiguous().requires_grad_(), k.contiguous().requires_grad_(), v.contiguous().requires_grad_()
    t_sf = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True))
    def sfb():
      o = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True); o.backward(dout.view(B, S, H, D).transpose(1, 2))
      q.grad = k.grad = v.grad = None
    t_sfb = timeit(sfb)
    print("B=%d S=%d H=%d | epl fwd %.3f ms (%.0f TF) fwd+bwd %.3f ms (%.0f TF) | sdpa fwd %.3f ms fwd+bwd %.3f ms" % (
        B, S, H, t_f, fl / t_f / 1e9, t_fb, 3.5 * fl / t_fb / 1e9, t_sf, t_sfb), flush=True)


if __name__ == "__main__":
  main()
