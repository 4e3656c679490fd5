"""A small MoE transformer (reference ``examples/moe``: T5-small-like, 8 experts, top-2, capacity 1.25)."""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import nn

from easyparallellibrary_b200.models.gpt2 import SelfAttention, GPT2Config
from easyparallellibrary_b200.ops.cross_entropy import softmax_cross_entropy
from easyparallellibrary_b200.ops.layernorm import LayerNorm
from easyparallellibrary_b200.ops.moe import MoEFFN


@dataclass
class MoEConfig:
  vocab_size: int = 32128
  d_model: int = 512
  d_ff: int = 2048
  n_layer: int = 6
  n_head: int = 8
  num_experts: int = 8
  capacity_factor: float = 1.25
  gating: str = "top2"
  aux_weight: float = 0.01
  n_positions: int = 512
  moe_every: int = 2              # every second FFN is an MoE layer


class MoEBlock(nn.Module):
  def __init__(self, cfg: MoEConfig, use_moe: bool, split):
    super().__init__()
    gcfg = GPT2Config(n_embd=cfg.d_model, n_head=cfg.n_head, n_layer=cfg.n_layer)
    self.ln1, self.attn, self.ln2 = LayerNorm(cfg.d_model), SelfAttention(gcfg), LayerNorm(cfg.d_model)
    self.use_moe = use_moe
    if use_moe:
      with split:
        self.ffn = MoEFFN(cfg.d_model, cfg.d_ff, cfg.num_experts, cfg.capacity_factor, cfg.gating)
    else:
      from easyparallellibrary_b200.models.gpt2 import MLP
      self.ffn = MLP(GPT2Config(n_embd=cfg.d_model, n_layer=cfg.n_layer))

  def forward(self, x):
    x = x + self.attn(self.ln1(x))
    return x + self.ffn(self.ln2(x))


class MoETransformer(nn.Module):
  def __init__(self, cfg: MoEConfig, expert_parallel: int = 1):
    super().__init__()
    import easyparallellibrary_b200 as epl
    self.cfg = cfg
    self._split = epl.split(device_count=expert_parallel)
    self.wte = nn.Embedding(cfg.vocab_size, cfg.d_model)
    self.wpe = nn.Embedding(cfg.n_positions, cfg.d_model)
    self.blocks = nn.ModuleList([MoEBlock(cfg, (i + 1) % cfg.moe_every == 0, self._split) for i in range(cfg.n_layer)])
    self.ln_f = LayerNorm(cfg.d_model)
    nn.init.normal_(self.wte.weight, std=0.02)
    nn.init.normal_(self.wpe.weight, std=0.02)

  def forward(self, idx, labels=None):
    pos = torch.arange(idx.shape[1], device=idx.device)
    x = self.wte(idx) + self.wpe(pos)
    for b in self.blocks:
      x = b(x)
    logits = torch.nn.functional.linear(self.ln_f(x), self.wte.weight)
    if labels is None:
      return logits
    loss = softmax_cross_entropy(logits, labels)
    aux = sum(b.ffn.aux_loss for b in self.blocks if b.use_moe and b.ffn.aux_loss is not None)
    return loss + self.cfg.aux_weight * aux
