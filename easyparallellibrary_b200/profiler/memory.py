"""Memory profiling (reference ``epl/profiler/memory_profiler_hook.py:207-271``: per-device allocation
timeline from RunMetadata, peak bytes, CSV + PNG with persistent / gradient / optimizer phases).

Eager equivalent: the caching allocator's counters are sampled at the phase boundaries of every profiled step — before the
step (*persistent*: weights, gradient buckets, optimizer state), at the end of every forward (*forward*: + activations), of
every backward (*backward*: gradients produced, activations freed) and of the apply phase (*apply*: optimizer temporaries) —
through the engine's phase scopes (``ir/phase.py`` listeners); the peak counter is reset at each boundary, so every row also
carries the peak *inside* its phase.  ``save()`` writes the CSV and — when matplotlib is importable — a PNG coloured by phase.
"""
from __future__ import annotations

import csv
import os
from typing import Dict, List, Optional

import torch


class MemoryProfilerHook(object):
  def __init__(self, save_steps: int = 1, max_steps: int = 10, output_dir: str = "./memory_profile", device=None):
    self.save_steps, self.max_steps, self.output_dir = save_steps, max_steps, output_dir
    self.device = device
    self.rows: List[Dict[str, float]] = []
    self._step = 0
    self._active = False
    from easyparallellibrary_b200.ir import phase as phase_lib
    self._phase_lib = phase_lib
    phase_lib.add_phase_listener(self._on_phase)

  def close(self) -> None:
    self._phase_lib.remove_phase_listener(self._on_phase)

  def _on_phase(self, phase, edge: str) -> None:
    if not self._active or edge != "exit" or phase.value not in ("forward", "backward", "apply"):
      return
    self.rows.append(dict(step=self._step, phase=phase.value, **self._stats()))
    if torch.cuda.is_available():
      torch.cuda.reset_peak_memory_stats(self.device)

  def _stats(self) -> Dict[str, float]:
    if not torch.cuda.is_available():
      return {"allocated": 0.0, "reserved": 0.0, "peak": 0.0}
    return {"allocated": float(torch.cuda.memory_allocated(self.device)), "reserved": float(torch.cuda.memory_reserved(self.device)),
            "peak": float(torch.cuda.max_memory_allocated(self.device))}

  def before_step(self, trainer) -> None:
    self._active = self._step % self.save_steps == 0 and len({r["step"] for r in self.rows}) < self.max_steps
    if self._active:
      if torch.cuda.is_available():
        torch.cuda.reset_peak_memory_stats(self.device)
      self.rows.append(dict(step=self._step, phase="persistent", **self._stats()))

  def after_step(self, trainer, out) -> None:
    if self._active:
      self.rows.append(dict(step=self._step, phase="after_step", **self._stats()))
    self._active = False
    self._step += 1

  def phase_peaks(self) -> Dict[str, float]:
    """Largest in-phase peak (bytes) per phase over the profiled steps."""
    out: Dict[str, float] = {}
    for r in self.rows:
      out[r["phase"]] = max(out.get(r["phase"], 0.0), r["peak"])
    return out

  @property
  def peak_bytes(self) -> float:
    return max((r["peak"] for r in self.rows), default=0.0)

  def save(self) -> Optional[str]:
    if not self.rows:
      return None
    os.makedirs(self.output_dir, exist_ok=True)
    path = os.path.join(self.output_dir, "memory_timeline.csv")
    with open(path, "w", newline="") as f:
      w = csv.DictWriter(f, fieldnames=list(self.rows[0].keys()))
      w.writeheader()
      w.writerows(self.rows)
    try:
      import matplotlib
      matplotlib.use("Agg")
      import matplotlib.pyplot as plt
      fig, ax = plt.subplots(figsize=(8, 3))
      xs = list(range(len(self.rows)))
      ax.plot(xs, [r["allocated"] / 2 ** 30 for r in self.rows], label="allocated")
      ax.plot(xs, [r["peak"] / 2 ** 30 for r in self.rows], label="peak")
      ax.set_ylabel("GiB")
      ax.legend()
      fig.savefig(os.path.join(self.output_dir, "memory_timeline.png"), dpi=100)
      plt.close(fig)
    except Exception:
      pass
    return path


def profile_memory(trainer) -> Dict[str, float]:
  """Static accounting of what the engine holds per rank (bytes)."""
  trainer.build()
  out = {"weights": 0, "gradients": 0, "optimizer_state_device": 0, "optimizer_state_host": 0}
  for s in trainer.group_keys:
    flat = trainer.flats[s]
    out["weights"] += sum(t.numel() * t.element_size() for t in flat.flat_params.values())
    out["gradients"] += sum(t.numel() * t.element_size() for t in flat.flat_grads.values())
    for o in trainer.optimizers[s]:
      for t in (o.master, o.m, o.v):
        if t is not None:
          key = "optimizer_state_device" if t.device.type != "cpu" or trainer.device.type == "cpu" else "optimizer_state_host"
          out[key] += t.numel() * t.element_size()
  return out
