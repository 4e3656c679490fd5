"""Tensor fusion ("coalescing") for collectives.

Policy parity with ``epl/communicators/rewriters/coalescing.py:121-199``:
tensors keep their given order, are grouped by dtype first; when the number of
dtypes reaches ``max_splits`` there is one bucket per dtype; otherwise every
dtype gets a share of the ``max_splits`` budget proportional to its span and
its tensors are cut greedily at ``bytes / (k - 1)``.  Optional 16-bit wire
compression with a scale (341-342, 374-378).

What differs on B200: buckets are *persistent flat buffers* (``FlatBucket``)
whose member tensors are **views** — gradients are produced directly inside the
bucket, so the per-step "copy into and out of every fused buffer"
(``coalescing.py:212-240``) disappears, and the same buffers can be registered
as symmetric memory for the in-kernel NVLink reduce-scatter.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch


def plan_buckets(nbytes: Sequence[int], dtypes: Sequence[object], max_splits: int, use_native: bool = True) -> List[List[int]]:
  """Return buckets as lists of indices into the input (order preserved inside a dtype).

  The planning loop runs in the native runtime (``csrc/runtime.cpp: epl_plan_buckets``) when the library is
  built; the Python body below is the reference implementation and the fallback (they are cross-checked in
  ``tests/test_native_runtime.py``)."""
  n = len(nbytes)
  if use_native and n > 1:
    try:
      from easyparallellibrary_b200.runtime import native
      if native.available():
        ids, table = [], {}
        for dt in dtypes:
          ids.append(table.setdefault(dt, len(table)))
        return native.plan_buckets([int(b) for b in nbytes], ids, int(max_splits))
    except Exception:  # pragma: no cover - fall back to the Python planner
      pass
  if n == 0:
    return []
  if n == 1:
    return [[0]]
  by_dtype: "Dict[object, List[int]]" = {}
  for i, dt in enumerate(dtypes):
    by_dtype.setdefault(dt, []).append(i)
  groups = list(by_dtype.values())
  if len(groups) >= max_splits:
    return groups
  # budget per dtype, proportional to how many tensors it spans (the reference scores by tick span)
  spans = [max(len(g) - 1, 0) for g in groups]
  total_span = sum(spans) or 1
  budget = [max(int(max_splits * s / total_span), 1) for s in spans]
  budget[0] += max_splits - sum(budget)
  budget = [max(b, 1) for b in budget]
  buckets: List[List[int]] = []
  for g, k in zip(groups, budget):
    sizes = [nbytes[i] for i in g]
    nonzero = [s for s in sizes if s] or [1]
    mean = float(sum(nonzero)) / len(nonzero)
    sizes = [s if s else mean for s in sizes]
    tot = 0.0
    for x in sizes:           # plain left-to-right accumulation (Python's sum() is compensated; the native planner is not)
      tot += x
    limit = tot if k == 1 else tot / (k - 1)
    cur: List[int] = []
    acc = 0.0
    for idx, s in zip(g, sizes):
      if cur and acc + s > limit:
        buckets.append(cur)
        cur, acc = [], 0.0
      cur.append(idx)
      acc += s
    if cur:
      buckets.append(cur)
  return buckets


def estimate_split_num_for_comm(tensors, split_bytes: int = 32 << 20) -> int:
  """ceil(bytes / 32 MiB) summed per dtype (reference ``collective_communicator.py:183-204``)."""
  if isinstance(tensors, torch.Tensor):
    tensors = [tensors]
  per_dtype: Dict[object, int] = {}
  for t in tensors:
    if t.numel():
      per_dtype[t.dtype] = per_dtype.get(t.dtype, 0) + t.numel() * t.element_size()
  total = sum((b + split_bytes - 1) // split_bytes for b in per_dtype.values())
  return total if total > 1 else 1


def flatten(tensors: Sequence[torch.Tensor]) -> torch.Tensor:
  return torch.cat([t.reshape(-1) for t in tensors]) if len(tensors) != 1 else tensors[0].reshape(-1).clone()


def unflatten(flat: torch.Tensor, like: Sequence[torch.Tensor]) -> List[torch.Tensor]:
  out, off = [], 0
  for t in like:
    n = t.numel()
    out.append(flat[off:off + n].view(t.shape))
    off += n
  return out


def compress(flat: torch.Tensor, scale: float, wire_dtype: torch.dtype = torch.float16) -> torch.Tensor:
  if flat.dtype in (torch.float32, torch.float64):
    return (flat * scale).to(wire_dtype)
  return flat


def decompress(flat: torch.Tensor, scale: float, dtype: torch.dtype) -> torch.Tensor:
  if flat.dtype != dtype:
    return flat.to(dtype) / scale
  return flat


class FlatBucket(object):
  """A persistent flat buffer; ``views[i]`` aliases the storage of member ``i``.

  ``align`` pads every member to a multiple of ``align`` elements so that each
  rank's shard boundary in a reduce-scatter never splits a 16-byte vector.
  """

  def __init__(self, shapes: Sequence[torch.Size], dtype: torch.dtype, device, align: int = 1, pad_to: int = 1,
               buffer: Optional[torch.Tensor] = None):
    self.shapes = [torch.Size(s) for s in shapes]
    self.dtype = dtype
    self.offsets: List[int] = []
    off = 0
    for s in self.shapes:
      self.offsets.append(off)
      n = int(torch.Size(s).numel())
      off += (n + align - 1) // align * align
    self.numel = (off + pad_to - 1) // pad_to * pad_to if off else 0
    if buffer is None:
      buffer = torch.zeros(self.numel, dtype=dtype, device=device)
    elif buffer.numel() < self.numel:
      raise ValueError("provided buffer too small")
    self.buffer = buffer
    self.views = [self.buffer[o:o + int(s.numel())].view(s) for o, s in zip(self.offsets, self.shapes)]

  def zero_(self) -> None:
    self.buffer.zero_()

  def shard(self, rank: int, world: int) -> torch.Tensor:
    n = self.numel // world
    return self.buffer[rank * n:(rank + 1) * n]
