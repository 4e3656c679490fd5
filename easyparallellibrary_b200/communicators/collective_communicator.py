"""The communicator façade used by every parallel feature.

Parity: ``epl/communicators/collective_communicator.py:33-181`` —
``CollectiveCommunicator(name, devices, max_splits, num_communicators,
enable_fp16, fp16_scale)`` with ``batch_allreduce / broadcast / allgather /
alltoall / reduce`` — plus the verbs the reference only reaches through
autodiff or not at all (``reduce_scatter``, ``allgatherv``, ``alltoallv``,
``send/recv``).  Factory helpers mirror ``parallel/ops.py:421-451``.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from easyparallellibrary_b200.communicators import coalescing
from easyparallellibrary_b200.communicators.backend import make_backend
from easyparallellibrary_b200.communicators.pool import CommunicationPool
from easyparallellibrary_b200.utils import constant

_REGISTRY: "Dict[str, CollectiveCommunicator]" = {}


class CollectiveCommunicator(object):
  def __init__(self, name: str, ranks: Sequence[int], max_splits: Optional[int] = None,
               num_communicators: Optional[int] = None, enable_fp16: Optional[bool] = None,
               fp16_scale: Optional[float] = None, device: Optional[torch.device] = None,
               prefer_native: bool = False, wire_dtype: torch.dtype = torch.float16):
    from easyparallellibrary_b200.env import Env
    cfg = Env.get().config.communication
    self.name = name
    self.ranks = [int(getattr(r, "rank", r)) for r in ranks]
    self.max_splits = cfg.max_splits if max_splits is None else max_splits
    self.num_communicators = cfg.num_communicators if num_communicators is None else num_communicators
    self.enable_fp16 = cfg.fp16 if enable_fp16 is None else enable_fp16
    self.fp16_scale = float(cfg.fp16_scale if fp16_scale is None else fp16_scale)
    self.wire_dtype = wire_dtype
    self.device = device
    n = max(1, self.num_communicators if len(self.ranks) > 1 else 1)
    self.pool = CommunicationPool([make_backend(self.ranks, prefer_native, device, copy=c) for c in range(n)])
    Env.get().comm_resources["%s#%d" % (name, len(Env.get().comm_resources))] = self

  # -- introspection -------------------------------------------------------------------
  @property
  def size(self) -> int:
    return self.pool.backends[0].size

  @property
  def rank(self) -> int:
    return self.pool.backends[0].rank

  @property
  def primary(self):
    return self.pool.backends[0]

  # -- fused all-reduce ----------------------------------------------------------------
  def plan(self, tensors: Sequence[torch.Tensor]) -> List[List[int]]:
    return coalescing.plan_buckets([t.numel() * t.element_size() for t in tensors], [t.dtype for t in tensors],
                                   self.max_splits)

  def batch_allreduce(self, tensors: Sequence[torch.Tensor], mean: bool = False, op: str = "sum") -> List[torch.Tensor]:
    """Fuse, (compress), all-reduce on the pool, (decompress), un-fuse; order preserved."""
    tensors = list(tensors)
    if not tensors:
      return []
    if self.size == 1:
      return [t.clone() for t in tensors]
    plan = self.plan(tensors)
    flats = [coalescing.flatten([tensors[i] for i in b]) for b in plan]
    dtypes = [f.dtype for f in flats]
    if self.enable_fp16:
      flats = [coalescing.compress(f, self.fp16_scale, self.wire_dtype) for f in flats]
    reduced = self.pool.communicate(flats, lambda be, f: be.all_reduce(f, op))
    out: List[torch.Tensor] = [None] * len(tensors)
    for b, f, dt in zip(plan, reduced, dtypes):
      if self.enable_fp16:
        f = coalescing.decompress(f, self.fp16_scale, dt)
      if mean:
        f = f / self.size if f.is_floating_point() else f // self.size
      for i, t in zip(b, coalescing.unflatten(f, [tensors[i] for i in b])):
        out[i] = t
    return out

  def allreduce(self, t: torch.Tensor, mean: bool = False, op: str = "sum") -> torch.Tensor:
    return self.batch_allreduce([t], mean=mean, op=op)[0]

  # -- the other verbs -----------------------------------------------------------------
  def broadcast(self, tensors, root: int = 0):
    single = isinstance(tensors, torch.Tensor)
    ts = [tensors] if single else list(tensors)
    if self.size > 1 and ts:
      saved = self.max_splits
      self.max_splits = min(constant.SERIAL_COMM_MAX_SPLITS, coalescing.estimate_split_num_for_comm(ts))
      plan = self.plan(ts)
      self.max_splits = saved
      for b in plan:
        flat = coalescing.flatten([ts[i] for i in b])
        self.primary.broadcast(flat, root)
        for i, v in zip(b, coalescing.unflatten(flat, [ts[i] for i in b])):
          ts[i].copy_(v)
    return ts[0] if single else ts

  def reduce(self, t: torch.Tensor, root: int = 0, op: str = "sum", mean: bool = False) -> torch.Tensor:
    out = self.primary.reduce(t.clone(), root, op)
    if mean and self.rank == root:
      out = out / self.size
    return out

  def allgather(self, t: torch.Tensor) -> torch.Tensor:
    return self.primary.all_gather(t)

  def allgatherv(self, t: torch.Tensor):
    return self.primary.all_gatherv(t)

  def reduce_scatter(self, t: torch.Tensor, op: str = "sum") -> torch.Tensor:
    return self.primary.reduce_scatter(t, op)

  def alltoall(self, t: torch.Tensor) -> torch.Tensor:
    if self.enable_fp16 and t.dtype == torch.float32:
      return self.primary.all_to_all(t.to(self.wire_dtype)).to(torch.float32)
    return self.primary.all_to_all(t)

  def alltoallv(self, t: torch.Tensor, send_counts: torch.Tensor):
    return self.primary.all_to_allv(t, send_counts)

  def send(self, t, dst): return self.primary.send(t, dst)
  def recv(self, t, src): return self.primary.recv(t, src)
  def barrier(self): return self.primary.barrier()

  def close(self) -> None:
    self.pool.close()


# ---------------------------------------------------------------------------------------------
# factories (reference parallel/ops.py:421-451)
# ---------------------------------------------------------------------------------------------
def create_communicator(name: str, ranks: Sequence[int], **kw) -> CollectiveCommunicator:
  return CollectiveCommunicator(name, ranks, **kw)


def create_serial_communicator(name: str, ranks: Sequence[int], **kw) -> CollectiveCommunicator:
  """One transport, generous split budget: weight broadcast and ZeRO traffic."""
  kw.setdefault("max_splits", constant.SERIAL_COMM_MAX_SPLITS)
  kw.setdefault("num_communicators", 1)
  kw.setdefault("enable_fp16", False)
  return CollectiveCommunicator(name, ranks, **kw)


def create_simple_communicator(name: str, ranks: Sequence[int], **kw) -> CollectiveCommunicator:
  """One transport, no fusion: tiny latency-bound messages (TP softmax statistics, metrics)."""
  kw.setdefault("max_splits", 1)
  kw.setdefault("num_communicators", 1)
  kw.setdefault("enable_fp16", False)
  return CollectiveCommunicator(name, ranks, **kw)


def get_or_create(name: str, ranks: Sequence[int], kind: str = "simple", **kw) -> CollectiveCommunicator:
  """Process-wide cache so repeated layers share one communicator (reference: shared_name)."""
  key = "%s/%s" % (name, ",".join(str(int(getattr(r, "rank", r))) for r in ranks))
  comm = _REGISTRY.get(key)
  if comm is None:
    maker = {"simple": create_simple_communicator, "serial": create_serial_communicator,
             "pooled": create_communicator}[kind]
    comm = _REGISTRY[key] = maker(name, ranks, **kw)
  return comm


def reset_registry() -> None:
  for c in _REGISTRY.values():
    c.close()
  _REGISTRY.clear()
