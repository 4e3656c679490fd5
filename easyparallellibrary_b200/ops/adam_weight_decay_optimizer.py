"""``AdamWeightDecayOptimizer`` — the name users of the reference import (``epl/ops/adam_weight_decay_optimizer.py``).

BERT-style AdamW: decoupled weight decay, no bias correction, ``exclude_from_weight_decay`` name patterns.  It is a thin
description object: pass it to ``Trainer(model, optimizer=AdamWeightDecayOptimizer(...))`` and the engine runs the update
with the fused flat-shard kernel (``csrc/optim.cu``), so ZeRO sharding / grouped apply / gradient accumulation / offload all
keep working exactly as with the built-in ``"adamw"``.
"""
from __future__ import annotations

import re
from typing import Optional, Sequence


class AdamWeightDecayOptimizer(object):
  kind = "adamw"

  def __init__(self, learning_rate: float, weight_decay_rate: float = 0.0, beta_1: float = 0.9, beta_2: float = 0.999,
               epsilon: float = 1e-6, exclude_from_weight_decay: Optional[Sequence[str]] = None, name: str = "AdamWeightDecayOptimizer"):
    self.learning_rate, self.weight_decay_rate = learning_rate, weight_decay_rate
    self.beta_1, self.beta_2, self.epsilon = beta_1, beta_2, epsilon
    self.exclude_from_weight_decay = list(exclude_from_weight_decay or [])
    self.name = name

  def trainer_kwargs(self, model=None) -> dict:
    kw = dict(lr=self.learning_rate, weight_decay=self.weight_decay_rate, betas=(self.beta_1, self.beta_2), eps=self.epsilon,
              bias_correction=False)
    if self.exclude_from_weight_decay and model is not None:
      names = {id(p): n for n, p in model.named_parameters()}
      pats = [re.compile(p) for p in self.exclude_from_weight_decay]

      def no_decay(p):
        n = names.get(id(p), "")
        return any(r.search(n) for r in pats)
      kw["no_decay"] = no_decay
    return kw
