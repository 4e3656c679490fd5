"""The library arm of the benchmark (``bench.py --impl baseline``): the best that stock libraries do for the same
model / precision / step, with NONE of this repo's kernels or engine on the path.

* GPT-2: ``nn.Linear`` (cuBLASLt with the bias epilogue), ``F.gelu`` , ``F.scaled_dot_product_attention`` (cuDNN / flash
  SDPA, causal), ``F.layer_norm``, ``F.cross_entropy`` on bf16 logits;
* bf16 parameters and gradients, fp32 master weights + moments updated by the fused multi-tensor AdamW
  (``torch.optim.AdamW(fused=True)`` on the fp32 masters, ``torch._foreach_copy_`` for the two casts) — the same
  "bf16 compute, fp32 master" contract as the product;
* data parallel: ``DistributedDataParallel`` (bucketed NCCL all-reduce of the bf16 gradients overlapped with backward,
  ``gradient_as_bucket_view``, static graph).

This is what the reference's algorithm (bucketed all-reduce over a communicator pool, ``epl/parallel/graph_editor.py:670-725``
+ ``communication_pool.py:84-105``, stock framework kernels for all math) amounts to on today's library stack.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn


class Block(nn.Module):
  def __init__(self, d: int, n_head: int, n_layer: int):
    super().__init__()
    self.n_head = n_head
    self.ln_1 = nn.LayerNorm(d)
    self.qkv = nn.Linear(d, 3 * d)
    self.proj = nn.Linear(d, d)
    self.ln_2 = nn.LayerNorm(d)
    self.fc = nn.Linear(d, 4 * d)
    self.out = nn.Linear(4 * d, d)
    for lin, std in ((self.qkv, 0.02), (self.fc, 0.02), (self.proj, 0.02 / math.sqrt(2 * n_layer)), (self.out, 0.02 / math.sqrt(2 * n_layer))):
      nn.init.normal_(lin.weight, std=std)
      nn.init.zeros_(lin.bias)

  def forward(self, x):
    B, S, d = x.shape
    q, k, v = self.qkv(self.ln_1(x)).view(B, S, 3, self.n_head, d // self.n_head).permute(2, 0, 3, 1, 4)
    a = F.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(B, S, d)
    x = x + self.proj(a)
    # (torch._addmm_activation, the cuBLASLt bias+GELU epilogue, has no autograd formula in torch 2.11 and does not return
    #  the pre-activation its backward would need: bias epilogue + an elementwise GELU kernel is the stock training path)
    h = F.gelu(self.fc(self.ln_2(x)), approximate="tanh")
    return x + self.out(h)


class TorchGPT2(nn.Module):
  def __init__(self, vocab: int, n_pos: int, d: int, n_layer: int, n_head: int):
    super().__init__()
    self.wte = nn.Embedding(vocab, d)
    self.wpe = nn.Embedding(n_pos, d)
    self.h = nn.ModuleList(Block(d, n_head, n_layer) for _ in range(n_layer))
    self.ln_f = nn.LayerNorm(d)
    nn.init.normal_(self.wte.weight, std=0.02)
    nn.init.normal_(self.wpe.weight, std=0.02)

  def forward(self, idx, labels):
    x = self.wte(idx) + self.wpe(torch.arange(idx.shape[1], device=idx.device))
    for b in self.h:
      x = b(x)
    logits = F.linear(self.ln_f(x), self.wte.weight)                         # tied head
    return F.cross_entropy(logits.view(-1, logits.shape[-1]), labels.reshape(-1))


class MasterAdamW(object):
  """bf16 model parameters, fp32 master copies + fused multi-tensor AdamW on the masters."""

  def __init__(self, params, lr: float, weight_decay: float):
    self.model_params = [p for p in params if p.requires_grad]
    self.masters = [p.detach().float().clone().requires_grad_(True) for p in self.model_params]
    decay = [m for m, p in zip(self.masters, self.model_params) if p.dim() > 1]
    no_decay = [m for m, p in zip(self.masters, self.model_params) if p.dim() <= 1]
    self.opt = torch.optim.AdamW([{"params": decay, "weight_decay": weight_decay}, {"params": no_decay, "weight_decay": 0.0}],
                                 lr=lr, betas=(0.9, 0.999), eps=1e-8, fused=True)
    for m in self.masters:
      m.grad = torch.zeros_like(m)

  @torch.no_grad()
  def step(self):
    torch._foreach_copy_([m.grad for m in self.masters], [p.grad for p in self.model_params])     # bf16 grads -> fp32
    self.opt.step()
    torch._foreach_copy_(self.model_params, self.masters)                                            # fp32 masters -> bf16 weights


class LibraryTrainer(object):
  """``step(tokens, labels) -> loss`` with the same signature role as ``epl.Trainer.step`` in bench.py."""

  def __init__(self, model: nn.Module, device, world: int, lr: float = 1e-4, weight_decay: float = 0.01):
    self.raw = model.to(device=device, dtype=torch.bfloat16)
    self.world = world
    if world > 1:
      from torch.nn.parallel import DistributedDataParallel as DDP
      self.model = DDP(self.raw, device_ids=[device.index], gradient_as_bucket_view=True, static_graph=True, bucket_cap_mb=200)
    else:
      self.model = self.raw
    self.opt = MasterAdamW(self.raw.parameters(), lr, weight_decay)

  def step(self, *batch):
    self.model.zero_grad(set_to_none=(self.world == 1))       # DDP: gradients are views of the communication buckets, zero in place
    loss = self.model(*batch)
    loss.backward()
    self.opt.step()
    return loss.detach()
