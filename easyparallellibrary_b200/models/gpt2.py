"""GPT-2 (117M ... 1.5B "XL") on the EPL-B200 ops.

The flagship benchmark model (``BASELINE.json``: GPT-2-XL, 48 layers, d=1600,
25 heads, sequence 1024).  Every hot op is an in-tree sm_100a kernel: tcgen05
GEMMs with bias/GELU/dGELU epilogues, LayerNorm fwd/bwd, fused softmax
cross-entropy, flash attention; the optimizer is the fused flat AdamW.

``epl_sequential()`` exposes ``[embed, block_0 .. block_{L-1}, head]`` so the
engine can cut pipeline stages at block boundaries — either where the user's
``epl.replicate`` scopes / ``epl.set_default_strategy`` calls put them (the
reference's BERT idiom, ``examples/bert/modeling.py:829-834``) or automatically.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch
from torch import nn

from easyparallellibrary_b200.ops.attention import attention
from easyparallellibrary_b200.ops.cross_entropy import softmax_cross_entropy
from easyparallellibrary_b200.ops.layernorm import LayerNorm
from easyparallellibrary_b200.ops.linear import Linear, mlp


@dataclass
class GPT2Config:
  vocab_size: int = 50304          # 50257 padded to a multiple of 64
  n_positions: int = 1024
  n_embd: int = 768
  n_layer: int = 12
  n_head: int = 12
  dropout: float = 0.0
  tie_embeddings: bool = True
  num_pipeline_stages: int = 1     # >1: blocks are spread over that many replicate taskgraphs

  @staticmethod
  def named(name: str, **kw) -> "GPT2Config":
    table = {
        "tiny": dict(n_embd=128, n_layer=2, n_head=4, n_positions=128, vocab_size=512),
        "small": dict(n_embd=768, n_layer=12, n_head=12),
        "medium": dict(n_embd=1024, n_layer=24, n_head=16),
        "large": dict(n_embd=1280, n_layer=36, n_head=20),
        "xl": dict(n_embd=1600, n_layer=48, n_head=25),
    }
    cfg = dict(table[name.lower().replace("gpt2-", "").replace("gpt2", "small") if name.lower() != "gpt2" else "small"])
    cfg.update(kw)
    return GPT2Config(**cfg)

  @property
  def num_params(self) -> int:
    d, L, V, P = self.n_embd, self.n_layer, self.vocab_size, self.n_positions
    per_layer = 12 * d * d + 13 * d
    return V * d + P * d + L * per_layer + 2 * d + (0 if self.tie_embeddings else V * d)

  def flops_per_token(self, seq_len: int) -> float:
    """Training FLOPs per token (fwd + bwd = 3x forward), matmuls only, causal attention counted at half."""
    d, L, V = self.n_embd, self.n_layer, self.vocab_size
    fwd = L * (24 * d * d + 2 * seq_len * d) + 2 * d * V
    return 3.0 * fwd


class Embedding(nn.Module):
  def __init__(self, cfg: GPT2Config):
    super().__init__()
    self.wte = nn.Embedding(cfg.vocab_size, cfg.n_embd)
    self.wpe = nn.Embedding(cfg.n_positions, cfg.n_embd)
    self.drop = nn.Dropout(cfg.dropout)
    nn.init.normal_(self.wte.weight, std=0.02)
    nn.init.normal_(self.wpe.weight, std=0.02)

  def forward(self, idx):
    pos = torch.arange(idx.shape[1], device=idx.device)
    return self.drop(self.wte(idx) + self.wpe(pos))


class SelfAttention(nn.Module):
  def __init__(self, cfg: GPT2Config):
    super().__init__()
    d = cfg.n_embd
    self.n_head = cfg.n_head
    self.qkv = Linear(d, 3 * d, init_std=0.02)
    self.proj = Linear(d, d, init_std=0.02 / math.sqrt(2 * cfg.n_layer))

  def forward(self, x, residual=None):
    B, S, d = x.shape
    from easyparallellibrary_b200.ops.attention import attention_packed
    from easyparallellibrary_b200.ops.linear import linear
    qkv = self.qkv(x).view(B, S, 3, self.n_head, d // self.n_head)
    y = attention_packed(qkv, causal=True)                    # [B, S, d] — no permutes / copies on either side
    return linear(y, self.proj.weight, self.proj.bias, residual=residual)


class MLP(nn.Module):
  def __init__(self, cfg: GPT2Config):
    super().__init__()
    d = cfg.n_embd
    self.fc = Linear(d, 4 * d, init_std=0.02)
    self.proj = Linear(4 * d, d, init_std=0.02 / math.sqrt(2 * cfg.n_layer))

  def forward(self, x, residual=None):
    return mlp(x, self.fc.weight, self.fc.bias, self.proj.weight, self.proj.bias, residual)


class Block(nn.Module):
  def __init__(self, cfg: GPT2Config):
    super().__init__()
    self.ln_1 = LayerNorm(cfg.n_embd)
    self.attn = SelfAttention(cfg)
    self.ln_2 = LayerNorm(cfg.n_embd)
    self.mlp = MLP(cfg)

  def forward(self, x):
    # pre-LN residual block; both residual adds run in GEMM epilogues, both gradient joins in the LN backward
    skip, h = self.ln_1.fork(x)
    x = self.attn(h, residual=skip)
    skip, h = self.ln_2.fork(x)
    return self.mlp(h, residual=skip)


class Head(nn.Module):
  def __init__(self, cfg: GPT2Config, tied: Optional[nn.Parameter]):
    super().__init__()
    self.ln_f = LayerNorm(cfg.n_embd)
    if tied is None:
      self.weight = nn.Parameter(torch.empty(cfg.vocab_size, cfg.n_embd))
      nn.init.normal_(self.weight, std=0.02)
    else:
      self.weight = tied

  def forward(self, x):
    from easyparallellibrary_b200.ops.linear import linear
    return linear(self.ln_f(x), self.weight)


def lm_loss(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
  return softmax_cross_entropy(logits, labels, ignore_index=-100, reduction="mean")


class GPT2(nn.Module):
  def __init__(self, cfg: GPT2Config):
    super().__init__()
    import easyparallellibrary_b200 as epl
    self.cfg = cfg
    stages = max(cfg.num_pipeline_stages, 1)
    if stages > 1 and cfg.tie_embeddings:
      raise ValueError("tied embeddings cannot be split across pipeline stages; set tie_embeddings=False")
    # stage boundaries balance FLOPs, not block counts: the vocabulary projection costs as much as V / (12 d) blocks
    # (2.6 blocks for GPT-2-XL) and sits on the last stage, so that stage gets fewer blocks
    head_blocks = cfg.vocab_size / (12.0 * cfg.n_embd)
    target = (cfg.n_layer + head_blocks) / stages
    starts = {}
    for k in range(1, stages):
      starts[min(max(int(round(k * target)), k), cfg.n_layer - (stages - k))] = k
    blocks = []
    if stages > 1:
      epl.set_default_strategy(epl.replicate(1, name="stage_0"))
    self.embed = Embedding(cfg)
    for i in range(cfg.n_layer):
      if stages > 1 and i in starts:
        epl.set_default_strategy(epl.replicate(1, name="stage_%d" % starts[i]))   # opens the next taskgraph
      blocks.append(Block(cfg))
    self.h = nn.ModuleList(blocks)
    self.head = Head(cfg, self.embed.wte.weight if cfg.tie_embeddings else None)

  def epl_sequential(self):
    return [self.embed] + list(self.h) + [self.head]

  def forward(self, idx, labels=None):
    x = self.embed(idx)
    for blk in self.h:
      x = blk(x)
    logits = self.head(x)
    if labels is None:
      return logits
    return lm_loss(logits, labels)
