"""Whole-step CUDA graph: forward + backward + gradient reduction + optimizer of one training step captured once and
replayed (the hot loop belongs to the runtime, not to a Python interpreter issuing ~1 600 launches a step).

What makes the step capturable:

* every kernel of the model path is an in-tree launch on the current stream with device pointers only; the per-step
  values the optimizer needs (learning rate, Adam bias corrections, gradient scale) live in a small device tensor
  (``FlatOptimizer.dyn`` / ``FusedDataParallel.dyn``) that the host refreshes before each replay;
* the fused data-parallel kernel keeps its barrier epoch on the device (``csrc/symm.cu``), so the per-bucket launches the
  gradient hooks issue on the side stream during backward are captured as forks of the graph and replay unchanged;
* gradients and weights live in persistent flat buckets, activations in the graph's private pool.

Not capturable (the trainer stays eager): pipelines (separate programs per stage), ZeRO-3 / offload (host-driven
streaming), dynamic loss scaling and clipping (host decisions per step), library collectives, collections.
"""
from __future__ import annotations

from typing import Any, Optional, Tuple

import torch

from easyparallellibrary_b200.ops import _lib
from easyparallellibrary_b200.utils.logging import get_logger


class GraphedStep(object):
  WARMUP = 3

  @staticmethod
  def eligible(tr) -> bool:
    from easyparallellibrary_b200.runtime import amp as amp_lib
    cfg = tr.config
    if tr.device.type != "cuda" or tr.plan.pipeline or tr.plan.num_stages > 1 or tr.zero3 or cfg.offload.level:
      return False
    if cfg.pipeline.num_micro_batch != 1 or tr.max_grad_norm is not None or tr._sparse or tr.baseline:
      return False
    if not isinstance(tr.scaler, amp_lib.NoLossScale) or tr.opt_kind not in ("adam", "adamw") or cfg.optimizer.num_apply_group != 1:
      return False
    if cfg.gradient_checkpoint.check_gradients:
      return False
    multi = any(c.size > 1 for c in tr.dp_comms.values())
    if not multi:
      return True
    # multi-rank: every bucket must go through the fused kernel (library collectives are not captured)
    return tr.fused is not None and all(tr.dp_comms[s].size == 1 or all(tr.fused.handles(s, b) for b in tr.flats[s].buckets)
                                        for s in tr.group_keys)

  def __init__(self, trainer):
    self.tr = trainer
    self.graph: Optional[torch.cuda.CUDAGraph] = None
    self.static_in: Optional[Tuple[torch.Tensor, ...]] = None
    self.static_loss: Optional[torch.Tensor] = None
    self.calls = 0
    self.enabled = True                                  # False: run eagerly (measurement aid; the device-side step values stay live)
    self.failed = False
    self.launches_per_replay = 0
    self.dyn = {}
    for s in trainer.group_keys:
      if trainer.fused is not None and s in trainer.fused.pads:
        continue                                         # the fused path owns its own dyn tensor
      d = torch.zeros(4, dtype=torch.float32, device=trainer.device)
      self.dyn[s] = d
      for o in trainer.optimizers[s]:
        o.dyn = d

  # ------------------------------------------------------------------ per-step host work (outside the graph)
  def _refresh(self, mean: bool, count: bool) -> None:
    tr = self.tr
    for s, d in self.dyn.items():
      opts = tr.optimizers[s]
      if count:
        for o in opts:
          o.step_count += 1
      h, t = opts[0].hyper, max(opts[0].step_count + (0 if count else 1), 1)
      if h.bias_correction:
        inv_c1, inv_c2 = 1.0 / (1.0 - h.beta1 ** t), 1.0 / (1.0 - h.beta2 ** t)
      else:
        inv_c1 = inv_c2 = 1.0
      scale = tr.scaler.inv_scale / (tr.mean_divisor(s) if mean else 1)
      d.copy_(torch.tensor([h.lr, inv_c1, inv_c2, scale], dtype=torch.float32), non_blocking=True)
    if tr.fused is not None and count:
      tr.fused.begin_step(mean)

  def _same_signature(self, batch) -> bool:
    return (self.static_in is not None and len(batch) == len(self.static_in) and
            all((isinstance(b, torch.Tensor) and b.shape == s.shape and b.dtype == s.dtype) for b, s in zip(batch, self.static_in)))

  # ------------------------------------------------------------------ capture
  def _capture(self, batch, kwargs) -> None:
    tr = self.tr
    self.static_in = tuple(b.clone() for b in batch)
    mean = tr._mean
    g = torch.cuda.CUDAGraph()
    l0 = _lib.launches
    tr._capturing = True
    counts = {id(o): o.step_count for s in tr.group_keys for o in tr.optimizers[s]}
    try:
      self._refresh(mean, count=False)
      with torch.cuda.graph(g):
        loss = tr._eager_body(self.static_in, kwargs, in_graph=True)
        self.static_loss = loss
    finally:
      tr._capturing = False
      for s in tr.group_keys:                            # the capture executed nothing: undo the host-side step counting
        for o in tr.optimizers[s]:
          o.step_count = counts[id(o)]
    self.launches_per_replay = _lib.launches - l0
    self.graph = g

  # ------------------------------------------------------------------ the step
  def step(self, batch, kwargs):
    tr = self.tr
    self.calls += 1
    if self.failed or not self.enabled or not all(isinstance(b, torch.Tensor) for b in batch) or kwargs:
      return None
    if self.calls <= self.WARMUP:
      return None                                        # eager warm-up steps (allocator, lazily configured kernels, NCCL)
    if self.graph is None:
      try:
        torch.cuda.synchronize(tr.device)
        self._capture(batch, kwargs)
        torch.cuda.synchronize(tr.device)
        get_logger().info("CUDA graph of the training step captured: %d kernel launches per replay", self.launches_per_replay)
      except Exception as e:      # pragma: no cover - depends on the CUDA runtime
        self.failed = True
        self.graph = None
        torch.cuda.synchronize(tr.device)
        get_logger().warning("CUDA graph capture of the training step failed (%s); staying eager", e)
        return None
    if not self._same_signature(batch):
      return None
    for s, b in zip(self.static_in, batch):
      s.copy_(b, non_blocking=True)
    self._refresh(tr._mean, count=True)
    self.graph.replay()
    _lib.launches += self.launches_per_replay
    return self.static_loss
