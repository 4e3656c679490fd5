"""Throughput / step-time meter.

The reference measures step time with wall-clock prints in its example scripts (``examples/resnet/resnet_dp.py:69-72``)
and has a one-shot PAI metric hook (``utils/metric.py:23-36``).  Here the meter is a small object the examples and
``bench.py`` share: device-side timing with CUDA events when a GPU is present (host wall-clock otherwise), a warm-up
window, and a whole-job aggregate (max over ranks) when ``torch.distributed`` is initialised.
"""
from __future__ import annotations

import time
from typing import Dict, Optional

import torch


class ThroughputMeter(object):
  def __init__(self, items_per_step: int, warmup: int = 3, unit: str = "tokens", device: Optional[torch.device] = None):
    self.items_per_step, self.warmup, self.unit = items_per_step, warmup, unit
    self.device = device
    self._use_cuda = bool(device is not None and device.type == "cuda" and torch.cuda.is_available())
    self._steps = 0
    self._t0 = None
    self._ev0 = None
    self._timed = 0

  def step(self) -> None:
    """Call once after every training step."""
    self._steps += 1
    if self._steps == self.warmup:
      if self._use_cuda:
        self._ev0 = torch.cuda.Event(enable_timing=True)
        self._ev0.record()
      self._t0 = time.perf_counter()
    elif self._steps > self.warmup:
      self._timed += 1

  def summary(self, world_items_multiplier: int = 1) -> Dict[str, float]:
    """ms/step and items/s over the timed window; the slowest rank defines the job's step time."""
    if self._timed == 0:
      return {"ms_per_step": float("nan"), "per_second": float("nan"), "steps": 0}
    if self._use_cuda:
      ev1 = torch.cuda.Event(enable_timing=True)
      ev1.record()
      ev1.synchronize()
      ms = self._ev0.elapsed_time(ev1)
    else:
      ms = (time.perf_counter() - self._t0) * 1e3
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
      t = torch.tensor([ms], dtype=torch.float64, device=self.device if self._use_cuda else "cpu")
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
      ms = float(t.item())
    per_step = ms / self._timed
    return {"ms_per_step": per_step, "per_second": self.items_per_step * world_items_multiplier / (per_step / 1e3),
            "steps": self._timed, "unit": self.unit}
