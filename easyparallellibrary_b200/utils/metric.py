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


class PhaseTimer(object):
  """Time one method of an object on the device (CUDA events on the current stream) or on the host (CPU runs).

  ``bench.py`` wraps ``Trainer._reduce_and_apply`` with it to report the *exposed* gradient-reduction + optimizer time per step
  (the fused bucket kernels run after backward on the main stream, so their event span is exactly what the step pays) next to
  that phase's roofline — the "exposed comm ms/step; fused-path % of roofline" part of the headline metric."""

  def __init__(self, obj, method: str, use_cuda: bool):
    self.obj, self.method, self.use_cuda = obj, method, bool(use_cuda)
    self.orig = getattr(obj, method)
    self.spans = []
    setattr(obj, method, self._call)

  def _call(self, *args, **kwargs):
    if self.use_cuda:
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      out = self.orig(*args, **kwargs)
      e1.record()
      self.spans.append((e0, e1))
    else:
      t0 = time.perf_counter()
      out = self.orig(*args, **kwargs)
      self.spans.append(time.perf_counter() - t0)
    return out

  def reset(self) -> None:
    self.spans = []

  def total_ms(self) -> float:
    if self.use_cuda:
      torch.cuda.synchronize()
      return float(sum(a.elapsed_time(b) for a, b in self.spans))
    return float(sum(self.spans) * 1e3)

  def restore(self) -> None:
    setattr(self.obj, self.method, self.orig)


def fused_dp_roofline_ms(num_params: int, world: int, hbm_gbs: float = 6400.0, nvlink_gbs: float = 770.0, grad_bytes: int = 2) -> float:
  """Lower bound of the WHOLE reduce-scatter + AdamW + all-gather phase for ``num_params`` parameters over ``world`` NVLink
  peers (it is an all-reduce's traffic: per direction and per GPU, (W-1)/W of the gradient bytes cross the link for the
  reduce-scatter — inbound: the peers' partials of my shard, outbound: my partials of their shards — and (W-1)/W of the
  weight bytes for the all-gather).  NVLink is full duplex and the HBM streams overlap the link traffic, so the bound is the
  MAX of the per-direction link time (at the measured 770 GB/s peer bandwidth, B200_PROFILING.md) and the HBM time
  (24 B/param of fp32 optimizer state for the 1/W shard + every gradient byte read once + every weight byte written once).
  With one GPU it is the local AdamW stream (30 B/param).  The *exposed* part of the phase can be shorter than this bound when
  buckets are processed while backward is still running."""
  if world <= 1:
    return num_params * 30.0 / (hbm_gbs * 1e9) * 1e3
  per_direction = 2.0 * (world - 1) / world * num_params * grad_bytes
  link = per_direction / (nvlink_gbs * 1e9)
  hbm = ((num_params / world) * 24.0 + 2.0 * num_params * grad_bytes) / (hbm_gbs * 1e9)
  return max(link, hbm) * 1e3
