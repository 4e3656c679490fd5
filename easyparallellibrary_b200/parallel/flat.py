"""Flat parameter / gradient storage.

Every parameter of a stage lives as a view into one flat buffer per dtype and
its ``.grad`` is a view into a matching flat gradient buffer, cut into
*buckets* (``communication.max_splits``, planned by
``communicators.coalescing.plan_buckets``).  Consequences, all B200-driven:

* gradient fusion costs nothing per step (the reference copies every gradient
  into and out of a fused buffer each step, ``coalescing.py:212-240``);
* a bucket is one contiguous range, so reduce-scatter / the fused
  reduce-scatter+Adam kernel shard it with pure pointer arithmetic, and the
  bucket buffers can be allocated as NVLink symmetric memory;
* optimizer state is a flat fp32 shard — one fused kernel launch per bucket.

Buckets are padded so that each of ``shard_world`` ranks owns an equal,
16-byte-aligned range.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence

import torch

from easyparallellibrary_b200.communicators.coalescing import plan_buckets

ALIGN_ELEMS = 8           # 16 B for 2-byte types, 32 B for fp32


@dataclass
class Bucket:
  index: int
  dtype: torch.dtype
  start: int                    # element offset in the flat buffer of its dtype
  numel: int                    # padded length (multiple of shard_world * ALIGN_ELEMS)
  params: List[torch.nn.Parameter]
  offsets: List[int]            # element offset of each param inside the bucket
  flat_param: torch.Tensor = None
  flat_grad: torch.Tensor = None
  ready: int = 0

  def shard_range(self, rank: int, world: int):
    n = self.numel // world
    return rank * n, (rank + 1) * n


class FlatParameters(object):
  def __init__(self, params: Sequence[torch.nn.Parameter], max_splits: int, shard_world: int = 1,
               grad_dtype: Optional[torch.dtype] = None,
               allocator: Optional[Callable[[int, torch.dtype, torch.device], torch.Tensor]] = None):
    params = [p for p in params if p.requires_grad]
    self.params = params
    self.shard_world = max(int(shard_world), 1)
    self.buckets: List[Bucket] = []
    self.flat_params: Dict[torch.dtype, torch.Tensor] = {}
    self.flat_grads: Dict[torch.dtype, torch.Tensor] = {}
    self.grad_dtype = grad_dtype
    if not params:
      return
    device = params[0].device
    alloc = allocator or (lambda n, dt, dev: torch.zeros(n, dtype=dt, device=dev))
    plan = plan_buckets([p.numel() * p.element_size() for p in params], [p.dtype for p in params], max_splits)
    quantum = self.shard_world * ALIGN_ELEMS
    per_dtype_off: Dict[torch.dtype, int] = {}
    layout = []
    for bi, idxs in enumerate(plan):
      dt = params[idxs[0]].dtype
      off, offsets = 0, []
      for i in idxs:
        offsets.append(off)
        off += (params[i].numel() + ALIGN_ELEMS - 1) // ALIGN_ELEMS * ALIGN_ELEMS
      numel = (off + quantum - 1) // quantum * quantum
      start = per_dtype_off.get(dt, 0)
      per_dtype_off[dt] = start + numel
      layout.append((bi, dt, start, numel, [params[i] for i in idxs], offsets))
    for dt, total in per_dtype_off.items():
      self.flat_params[dt] = alloc(total, dt, device)
      self.flat_grads[dt] = alloc(total, grad_dtype or dt, device)
    for bi, dt, start, numel, ps, offsets in layout:
      b = Bucket(bi, dt, start, numel, ps, offsets)
      b.flat_param = self.flat_params[dt][start:start + numel]
      b.flat_grad = self.flat_grads[dt][start:start + numel]
      with torch.no_grad():
        for p, o in zip(ps, offsets):
          view = b.flat_param[o:o + p.numel()].view(p.shape)
          view.copy_(p.data)
          p.data = view
          p.grad = b.flat_grad[o:o + p.numel()].view(p.shape) if (grad_dtype or dt) == p.dtype else None
      self.buckets.append(b)
    self.grad_dtype = grad_dtype

  @property
  def total_numel(self) -> int:
    return sum(b.numel for b in self.buckets)

  def zero_grad(self) -> None:
    for g in self.flat_grads.values():
      g.zero_()
    for b in self.buckets:
      b.ready = 0
      for p, o in zip(b.params, b.offsets):
        if p.grad is None and (self.grad_dtype or b.dtype) == p.dtype:
          p.grad = b.flat_grad[o:o + p.numel()].view(p.shape)

  def rebind_grads(self) -> None:
    """Re-attach ``.grad`` views (autograd or user code may have replaced them)."""
    for b in self.buckets:
      for p, o in zip(b.params, b.offsets):
        view = b.flat_grad[o:o + p.numel()].view(p.shape)
        if p.grad is None:
          continue
        if p.grad.data_ptr() != view.data_ptr():
          view.copy_(p.grad)
          p.grad = view if view.dtype == p.dtype else None

  def decay_mask(self, bucket: Bucket, no_decay: Callable[[torch.nn.Parameter], bool]) -> Optional[torch.Tensor]:
    """fp32 0/1 mask over the bucket: 0 for params that take no weight decay (biases, norms)."""
    if not any(no_decay(p) for p in bucket.params):
      return None
    mask = torch.ones(bucket.numel, dtype=torch.float32, device=bucket.flat_param.device)
    for p, o in zip(bucket.params, bucket.offsets):
      if no_decay(p):
        mask[o:o + p.numel()] = 0
    return mask
