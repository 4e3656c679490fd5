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
  """Dtype the *parameters* are cast to.  ``bf16``: bf16 weights + fp32 masters in the flat optimizer.  ``O1``: None — the
  parameters stay fp32 (they are the master weights) and precision is decided per op, see :class:`o1_autocast`."""
  level = (level or "").lower()
  if level in ("bf16", "fp8"):                 # fp8: bf16 weights / activations, forward GEMMs quantised per tensor to e4m3 (ops/fp8.py)
    return torch.bfloat16
  return None


# ------------------------------------------------------------------------------------------------
# O1: op-level colouring (reference auto_mixed_precision.py:35-85 allow / deny / gray / clear lists, :174-415 the four
# colouring passes, :160-172 just-in-time fp16 casts of fp32 variables).  The reference rewrites the graph; eager code decides
# at each op call: ops on the ALLOW list cast their inputs (and, just in time, their fp32 weights) to fp16; ops on the DENY
# list cast to fp32; GRAY ops (element-wise, residual adds, reshapes) run in whatever dtype their inputs have — which is what
# the reference's gray propagation converges to; CLEAR ops are dtype-agnostic.  Stock torch ops inside the region follow
# ``torch.autocast``'s own lists (same design: matmul/conv fp16, softmax/norm/loss fp32).
# ------------------------------------------------------------------------------------------------
ALLOW_OPS = ("linear", "mlp", "matmul", "conv", "attention")
DENY_OPS = ("layer_norm", "rms_norm", "softmax", "cross_entropy", "exp", "log", "pow", "sum", "mean", "loss", "batch_norm")
GRAY_OPS = ("add", "mul", "gelu", "relu", "tanh", "dropout", "residual", "concat", "where")
CLEAR_OPS = ("reshape", "transpose", "slice", "gather", "identity")
_O1_DEPTH = 0
_O1_LOG = False


def o1_active() -> bool:
  return _O1_DEPTH > 0


def op_dtype(kind: str) -> Optional[torch.dtype]:
  """fp16 / fp32 / None (= keep the input dtype) for an op of ``kind`` inside an O1 region; always None outside."""
  if _O1_DEPTH == 0:
    return None
  if kind in ALLOW_OPS:
    return torch.float16
  if kind in DENY_OPS:
    return torch.float32
  return None


def cast_args(kind: str, *tensors):
  """Cast floating tensors to the dtype the O1 policy assigns to ``kind`` (identity outside O1 regions / for gray ops)."""
  dt = op_dtype(kind)
  if dt is None:
    return tensors if len(tensors) != 1 else tensors[0]
  if _O1_LOG:
    from easyparallellibrary_b200.utils.logging import get_logger
    get_logger().info("amp O1: %s -> %s", kind, dt)
  out = tuple(t.to(dt) if isinstance(t, torch.Tensor) and t.is_floating_point() and t.dtype != dt else t for t in tensors)
  return out if len(out) != 1 else out[0]


class o1_autocast(object):
  """Context of one forward pass under ``amp.level = O1``."""

  def __init__(self, device_type: str, enabled: bool = True, debug_log: bool = False):
    self.enabled, self.debug_log = enabled, debug_log
    self.ctx = torch.autocast(device_type, dtype=torch.float16, enabled=enabled) if enabled else None

  def __enter__(self):
    global _O1_DEPTH, _O1_LOG
    if self.enabled:
      _O1_DEPTH += 1
      _O1_LOG = self.debug_log
      self.ctx.__enter__()
    return self

  def __exit__(self, *exc):
    global _O1_DEPTH
    if self.enabled:
      self.ctx.__exit__(*exc)
      _O1_DEPTH -= 1
    return False


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
