"""Exponential moving average of the weights as a Trainer hook.

The reference's users wrap their optimizer in ``tf.contrib.opt.MovingAverageOptimizer`` (``tests/multi_optimizer_test.py:30-36``:
Adam -> clip_gradients_by_norm -> MovingAverageOptimizer must keep working under EPL).  Here optimizer nesting does not exist —
clipping is ``Trainer(max_grad_norm=...)``, the update is one fused kernel — and the moving average is a hook::

    ema = WeightEMA(decay=0.999, num_updates=True)
    trainer.hooks.append(ema)
    ...train...
    with ema.swapped(trainer):          # evaluate / export with the averaged weights
      trainer.eval_step(x)

The shadow copies are fp32 and live next to the parameters this rank holds (every rank of a data-parallel group computes the same
average; each pipeline stage averages its own layers); ``torch._foreach`` keeps the update at two launches per dtype.  Parameter
partitioning (ZeRO-3 / weight offload) releases the weights between steps and is not supported.
"""
from __future__ import annotations

import contextlib
from typing import Dict, List, Optional

import torch


class WeightEMA(object):
  def __init__(self, decay: float = 0.999, num_updates: bool = False):
    """``num_updates``: TF's warm-up of the decay, ``min(decay, (1 + n) / (10 + n))`` with n = steps so far."""
    if not 0.0 <= decay < 1.0:
      raise ValueError("decay must be in [0, 1)")
    self.decay, self.num_updates = decay, num_updates
    self.shadow: Dict[str, torch.Tensor] = {}
    self._params: Optional[List[torch.nn.Parameter]] = None
    self._names: List[str] = []
    self.updates = 0

  def _bind(self, trainer) -> None:
    if getattr(trainer, "zero3", None):
      raise RuntimeError("WeightEMA needs resident weights: not available with zero.level=v3 / offload.weights")
    self._params, self._names = [], []
    for s in trainer.plan.local_stages:
      for name, p in trainer.stage_modules[s].named_parameters():
        if p.requires_grad:
          self._params.append(p)
          self._names.append("%d.%s" % (s, name))
    for n, p in zip(self._names, self._params):
      self.shadow.setdefault(n, p.detach().float().clone())

  def before_step(self, trainer) -> None:
    if self._params is None:
      self._bind(trainer)

  @torch.no_grad()
  def after_step(self, trainer, out) -> None:
    if out is not None and getattr(out, "skipped", False):
      return                                     # an overflow step did not move the weights
    d = self.decay
    if self.num_updates:
      d = min(d, (1.0 + self.updates) / (10.0 + self.updates))
    self.updates += 1
    shadows = [self.shadow[n] for n in self._names]
    current = [p.detach().float() for p in self._params]
    torch._foreach_mul_(shadows, d)
    torch._foreach_add_(shadows, current, alpha=1.0 - d)

  @contextlib.contextmanager
  def swapped(self, trainer):
    """Temporarily load the averaged weights into the model (and put the trained ones back afterwards)."""
    if self._params is None:
      self._bind(trainer)
    saved = [p.detach().clone() for p in self._params]
    with torch.no_grad():
      for n, p in zip(self._names, self._params):
        p.copy_(self.shadow[n].to(p.dtype))
    try:
      yield self
    finally:
      with torch.no_grad():
        for p, v in zip(self._params, saved):
          p.copy_(v)

  def state_dict(self):
    return {"updates": self.updates, "shadow": self.shadow}

  def load_state_dict(self, sd) -> None:
    self.updates = int(sd["updates"])
    for k, v in sd["shadow"].items():
      self.shadow[k] = v.clone()
