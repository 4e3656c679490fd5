"""Mixed precision.

Reference: O1 graph colouring to fp16 + fixed/dynamic loss scale
(``epl/runtime/amp/auto_mixed_precision.py``, ``loss_scale_tf.py:200-406``).
On B200 the natural policy is *bf16 compute weights + fp32 master weights held
by the flat optimizer* (no scaling needed); ``amp.level="O1"`` keeps the
reference's fp16 + loss-scale behaviour.  The dynamic scale starts at 2**15,
doubles after 2000 consecutive finite steps and halves (floor 1) while
skipping the update on any non-finite gradient.  Unlike the reference (TODO at
``loss_scale_tf.py:346-360``) the finite flag is all-reduced so every replica
takes the same decision.
"""
from __future__ import annotations

from typing import Optional

import torch


def compute_dtype(level: str) -> Optional[torch.dtype]:
  level = (level or "").lower()
  if level == "o1":
    return torch.float16
  if level == "bf16":
    return torch.bfloat16
  return None


class LossScaler(object):
  def scale(self, loss: torch.Tensor) -> torch.Tensor:
    return loss * self.loss_scale if self.loss_scale != 1.0 else loss

  @property
  def inv_scale(self) -> float:
    return 1.0 / self.loss_scale

  def update(self, found_inf: bool) -> bool:
    """Returns True when the optimizer step must be skipped."""
    return False


class NoLossScale(LossScaler):
  loss_scale = 1.0


class FixedLossScale(LossScaler):
  def __init__(self, value: float):
    if value < 1:
      raise ValueError("loss scale must be >= 1")
    self.loss_scale = float(value)

  def update(self, found_inf: bool) -> bool:
    return False          # reference: fixed scale never skips (loss_scale_tf.py:200-244)


class DynamicLossScale(LossScaler):
  def __init__(self, initial: float = 2.0 ** 15, increment_period: int = 2000, multiplier: float = 2.0):
    self.loss_scale = float(initial)
    self.increment_period = increment_period
    self.multiplier = multiplier
    self.good_steps = 0
    self.skipped = 0

  def update(self, found_inf: bool) -> bool:
    if found_inf:
      self.loss_scale = max(self.loss_scale / self.multiplier, 1.0)
      self.good_steps = 0
      self.skipped += 1
      return True
    self.good_steps += 1
    if self.good_steps >= self.increment_period:
      self.loss_scale *= self.multiplier
      self.good_steps = 0
    return False


def make_scaler(amp_level: str, loss_scale) -> LossScaler:
  if (amp_level or "").lower() != "o1":
    return NoLossScale()
  if isinstance(loss_scale, str) and loss_scale.lower() == "dynamic":
    return DynamicLossScale()
  return FixedLossScale(float(loss_scale))


_KEEP_FP32 = (torch.nn.modules.batchnorm._BatchNorm,)


def cast_module(model: torch.nn.Module, dtype: torch.dtype, debug_log: bool = False) -> None:
  """Cast floating parameters/buffers to ``dtype``; batch-norm statistics stay fp32
  (the reference's deny/gray lists keep normalisation statistics in fp32 too)."""
  from easyparallellibrary_b200.utils.logging import get_logger
  for name, mod in model.named_modules():
    keep = isinstance(mod, _KEEP_FP32) or getattr(mod, "epl_keep_fp32", False)
    if debug_log:
      get_logger().info("amp: %s -> %s", name or "<root>", "fp32" if keep else dtype)
    if keep:
      continue
    for p in mod.parameters(recurse=False):
      if p.is_floating_point():
        p.data = p.data.to(dtype)
    for bname, b in mod.named_buffers(recurse=False):
      if b is not None and b.is_floating_point():
        mod._buffers[bname] = b.to(dtype)
