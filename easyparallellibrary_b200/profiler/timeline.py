"""Kernel / operator timeline hook.

The reference's examples attach the stock ``tf.train.ProfilerHook`` (``examples/bert/run_squad.py:1246-1252``) to get a
Chrome trace of a few steps.  ``TimelineHook`` is the equivalent for :class:`Trainer`: it records ``steps`` training steps
with ``torch.profiler`` (CUPTI on a GPU, CPU activities otherwise) starting at ``start_step``, then writes

* ``<dir>/kernel_table.txt`` — per-kernel total time, call count and share of GPU busy time (the table ``bench.py --profile``
  prints; the one used to steer round-1 optimisation, see ``profiles/r1_step_kernel_timeline_*``), and
* ``<dir>/trace.json`` — a Chrome / Perfetto trace (optional).

A number measured while this hook is active is never a benchmark value.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch


def kernel_table(events, top: int = 40) -> Tuple[str, Dict[str, Tuple[float, int]]]:
  """Aggregate profiler events by name.  Returns (text table, {name: (total_ms, calls)})."""
  cuda = [e for e in events if getattr(e.device_type, "name", "") == "CUDA"]
  evs = cuda if cuda else [e for e in events if e.time_range.elapsed_us() > 0]
  agg: Dict[str, List[float]] = {}
  for e in evs:
    a = agg.setdefault(e.name, [0.0, 0])
    a[0] += e.time_range.elapsed_us() / 1e3
    a[1] += 1
  if not evs:
    return "no events recorded\n", {}
  span = (max(e.time_range.end for e in evs) - min(e.time_range.start for e in evs)) / 1e3
  busy = sum(v[0] for v in agg.values())
  lines = ["%s timeline: span %.2f ms, sum of event time %.2f ms (%.1f%% of span), %d events" % (
      "GPU kernel" if cuda else "CPU operator", span, busy, 100.0 * busy / max(span, 1e-9), len(evs))]
  for name, (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    lines.append("%6.2f%% %10.3f ms %7d  %s" % (100.0 * ms / max(busy, 1e-9), ms, n, name[:110]))
  return "\n".join(lines) + "\n", {k: (v[0], int(v[1])) for k, v in agg.items()}


class TimelineHook(object):
  def __init__(self, output_dir: str = "./timeline", start_step: int = 3, steps: int = 2, chrome_trace: bool = False):
    self.output_dir, self.start_step, self.steps, self.chrome_trace = output_dir, start_step, steps, chrome_trace
    self._step = 0
    self._prof = None
    self.table: Optional[str] = None
    self.by_name: Dict[str, Tuple[float, int]] = {}

  def before_step(self, trainer) -> None:
    if self._step == self.start_step and self._prof is None and self.table is None:
      from torch.profiler import ProfilerActivity, profile
      acts = [ProfilerActivity.CPU] + ([ProfilerActivity.CUDA] if torch.cuda.is_available() else [])
      self._prof = profile(activities=acts)
      self._prof.__enter__()

  def after_step(self, trainer, out) -> None:
    self._step += 1
    if self._prof is not None and self._step == self.start_step + self.steps:
      if torch.cuda.is_available():
        torch.cuda.synchronize()
      self._prof.__exit__(None, None, None)
      self.table, self.by_name = kernel_table(self._prof.events())
      os.makedirs(self.output_dir, exist_ok=True)
      with open(os.path.join(self.output_dir, "kernel_table.txt"), "w") as f:
        f.write(self.table)
      if self.chrome_trace:
        self._prof.export_chrome_trace(os.path.join(self.output_dir, "trace.json"))
      self._prof = None
