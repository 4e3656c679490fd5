"""FLOPs profiling (reference ``epl/profiler/flops.py``: ``FlopsProfilerHook`` + ``profile_flops``).

The reference asks ``tf.profiler`` for per-op float operations of a traced step.  Here one traced forward
pass (``ir/capture.py``) yields per-module forward FLOPs; a training step costs 3x (forward + two backward
GEMMs per forward GEMM).  The hook turns measured step time into achieved TFLOP/s and a fraction of the
measured B200 peak (``MEASURED_PEAKS.json``).
"""
from __future__ import annotations

import json
import os
import time
from typing import Any, Dict, List, Sequence

import torch

from easyparallellibrary_b200.ir.capture import trace_module_costs


def profile_flops(model: torch.nn.Module, example_inputs: Sequence[Any], by: str = "scope", depth: int = 2) -> Dict[str, float]:
  """Forward FLOPs grouped by module scope (``by="scope"``), module type (``"op"``) or in total (``"graph"``)."""
  nodes = trace_module_costs(model, example_inputs)
  out: Dict[str, float] = {}
  for n in nodes:
    key = {"scope": ".".join(n.name.split(".")[:depth]), "op": n.type}.get(by, "total")
    out[key] = out.get(key, 0.0) + n.flops
  out["__total__"] = sum(n.flops for n in nodes)
  return out


def measured_peaks() -> Dict[str, float]:
  here = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
  try:
    return json.load(open(os.path.join(here, "MEASURED_PEAKS.json")))
  except Exception:
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "fallback": True}


class FlopsProfilerHook(object):
  """``trainer.hooks.append(FlopsProfilerHook(flops_per_step))``; ``summary()`` after training."""

  def __init__(self, flops_per_step: float, every: int = 1, use_cuda_events: bool = True):
    self.flops_per_step, self.every = float(flops_per_step), every
    self.use_events = use_cuda_events and torch.cuda.is_available()
    self.records: List[float] = []
    self._t0 = None

  def before_step(self, trainer) -> None:
    if self.use_events:
      self._t0 = torch.cuda.Event(enable_timing=True)
      self._t0.record()
    else:
      self._t0 = time.perf_counter()

  def after_step(self, trainer, out) -> None:
    if self.use_events:
      t1 = torch.cuda.Event(enable_timing=True)
      t1.record()
      t1.synchronize()
      self.records.append(self._t0.elapsed_time(t1) / 1e3)
    else:
      self.records.append(time.perf_counter() - self._t0)

  def summary(self, skip: int = 1) -> Dict[str, float]:
    xs = self.records[skip:] or self.records
    if not xs:
      return {}
    t = sorted(xs)[len(xs) // 2]
    tf = self.flops_per_step / t / 1e12
    peaks = measured_peaks()
    return {"median_step_s": t, "tflops": tf, "fraction_of_measured_bf16_sustained": tf / peaks["bf16_tflops_sustained"]}
