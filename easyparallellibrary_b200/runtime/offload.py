"""CPU offload.

Reference ``offload.level="v0"`` (``graph_editor.py:727-751``): variables *and* the optimizer/apply ops are
pinned to the CPU; every variable read becomes a just-in-time host->device copy.  On a B200 node the host
link (PCIe Gen5, ~55 GB/s) is two orders of magnitude slower than HBM, so the design is:

* **optimizer-state offload** (:class:`OffloadedOptimizer`): the fp32 master weights and both Adam moments
  of a flat shard live in *pinned* host memory; per step they stream through a small double-buffered
  device window on a dedicated copy stream (``cudaMemcpyAsync``), the fused AdamW kernel runs on the
  window, and the results stream back — H2D of chunk *i+1* and D2H of chunk *i-1* overlap the kernel on
  chunk *i*.  Device memory drops by 12 B/param.
* **weight offload**: with ``offload.weights`` (default) the model runs through the per-layer engine of
  ``parallel/zero3.py`` with its weight shards in pinned host memory: a layer's weights are copied in (and, across
  ranks, all-gathered) one layer ahead of their use on a copy stream and released after use — the literal "weights on
  host, read just in time" behaviour of the reference, with the copy hidden behind the previous layer's compute.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
from torch import nn

from easyparallellibrary_b200.runtime.optimizer import FlatOptimizer

CHUNK_ELEMS = 32 << 20        # 128 MiB of fp32 per window buffer


class OffloadedOptimizer(FlatOptimizer):
  def __init__(self, kind: str, hyper, master_shard: torch.Tensor, decay_mask: Optional[torch.Tensor], device: torch.device):
    self.compute_device = device
    host = torch.empty(master_shard.numel(), dtype=torch.float32, pin_memory=device.type == "cuda")
    host.copy_(master_shard)
    super().__init__(kind, hyper, host, None, state_device=torch.device("cpu"))
    if device.type == "cuda":
      self.m = self.m.pin_memory() if self.m is not None else None
      self.v = self.v.pin_memory() if self.v is not None else None
      self.copy_stream = torch.cuda.Stream(device=device)
      n = min(CHUNK_ELEMS, max(master_shard.numel(), 1))
      self.window = [[torch.empty(n, dtype=torch.float32, device=device) for _ in range(3)] for _ in range(2)]
    self.device_mask = decay_mask

  def step(self, grad_shard, model_shard=None, grad_scale: float = 1.0, lo: int = 0, hi: Optional[int] = None,
           count_step: bool = True) -> None:
    if count_step:
      self.step_count += 1
    hi = self.master.numel() if hi is None else hi
    if hi <= lo:
      return
    if self.compute_device.type != "cuda":
      from easyparallellibrary_b200.runtime.optimizer import adamw_reference, sgd_reference
      sl = slice(lo, hi)
      mask = self.device_mask[sl] if self.device_mask is not None else None
      out = model_shard[sl] if model_shard is not None else None
      if self.kind == "sgd":
        sgd_reference(self.master[sl], grad_shard[sl], None if self.m is None else self.m[sl], self.hyper, grad_scale, out)
      else:
        adamw_reference(self.master[sl], grad_shard[sl], self.m[sl], self.v[sl], self.step_count, self.hyper, grad_scale, mask, out)
      return
    from easyparallellibrary_b200.ops import fused_optim
    main = torch.cuda.current_stream()
    cs = self.copy_stream
    chunks = [(a, min(a + self.window[0][0].numel(), hi)) for a in range(lo, hi, self.window[0][0].numel())]
    ready = [torch.cuda.Event() for _ in chunks]
    done = [torch.cuda.Event() for _ in chunks]
    free = [None, None]
    states = [self.master, self.m, self.v] if self.kind != "sgd" else [self.master, self.m, None]

    def upload(i):
      a, b = chunks[i]
      with torch.cuda.stream(cs):
        if free[i % 2] is not None:
          cs.wait_event(free[i % 2])                       # window slot still being written back
        for k, st in enumerate(states):
          if st is not None:
            self.window[i % 2][k][:b - a].copy_(st[a:b], non_blocking=True)
        ready[i].record(cs)

    upload(0)
    for i, (a, b) in enumerate(chunks):
      if i + 1 < len(chunks):
        upload(i + 1)
      main.wait_event(ready[i])
      wm, wmm, wv = (self.window[i % 2][k][:b - a] for k in range(3))
      mask = self.device_mask[a:b] if self.device_mask is not None else None
      out = model_shard[a:b] if model_shard is not None else None
      if self.kind == "sgd":
        fused_optim.sgd_step(wm, grad_shard[a:b], wmm if self.m is not None else None, self.hyper, grad_scale, out)
      else:
        fused_optim.adamw_step(wm, grad_shard[a:b], wmm, wv, self.step_count, self.hyper, grad_scale, mask, out)
      done[i].record(main)
      with torch.cuda.stream(cs):
        cs.wait_event(done[i])
        for k, st in enumerate(states):
          if st is not None:
            st[a:b].copy_(self.window[i % 2][k][:b - a], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(cs)
        free[i % 2] = ev
    main.wait_stream(cs)
