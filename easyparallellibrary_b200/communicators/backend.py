"""Transport backends behind :class:`CollectiveCommunicator`.

* :class:`TorchBackend` — ``torch.distributed`` process groups: ``nccl`` on
  GPUs, ``gloo`` on the CPU (the reference has no CPU path at all —
  ``epl/communicators/nccl.py:71-73`` raises on non-GPU devices; the CPU path
  here is what lets the plumbing tests run without hardware).  Verbs that gloo
  lacks (reduce-scatter, all-to-all) are composed from ones it has.
* :class:`NativeBackend` — the in-tree C++ communicator
  (``csrc/communicator.cpp``), one ``ncclComm_t`` + one side stream per pool
  slot, event-fenced against the caller's stream (the B200 equivalent of the
  reference's ``CudaStreamAsyncOpKernel``, ``tensorflow_cuda.h:50-136``).
* :class:`LocalBackend` — world size 1; every verb is the identity.

All verbs take and return plain tensors; autograd rules live in
``functional.py``.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

REDUCE_OPS = ("sum", "prod", "max", "min")


def _torch_op(op: str):
  return {"sum": dist.ReduceOp.SUM, "prod": dist.ReduceOp.PRODUCT, "max": dist.ReduceOp.MAX,
          "min": dist.ReduceOp.MIN}[op.lower()]


class LocalBackend(object):
  name = "local"

  def __init__(self, ranks: Sequence[int] = (0,)):
    self.ranks = list(ranks)
    self.size = 1
    self.rank = 0

  def all_reduce(self, t, op="sum"): return t
  def reduce(self, t, root=0, op="sum"): return t
  def broadcast(self, t, root=0): return t
  def all_gather(self, t): return t.clone()
  def all_gatherv(self, t): return t.clone(), torch.tensor([t.shape[0]], dtype=torch.int64)
  def reduce_scatter(self, t, op="sum"): return t.clone()
  def all_to_all(self, t): return t.clone()
  def all_to_allv(self, t, send_counts): return t.clone(), send_counts.clone()
  def all_reduce_async(self, t, op="sum"): return None
  def reduce_scatter_into(self, out, inp, op="sum", async_op=False):
    if out.data_ptr() != inp.data_ptr(): out.copy_(inp)
    return None
  def all_gather_into(self, out, inp, async_op=False):
    if out.data_ptr() != inp.data_ptr(): out.copy_(inp)
    return None
  def send(self, t, dst): raise RuntimeError("send on a single-rank communicator")
  def recv(self, t, src): raise RuntimeError("recv on a single-rank communicator")
  def barrier(self): return None
  def close(self): return None


_GROUPS = {}       # (ranks tuple, copy index) -> ProcessGroup, filled collectively by register_groups()


def register_groups(rank_lists: Sequence[Sequence[int]], copies: int = 1) -> None:
  """``dist.new_group`` is collective over the whole world and order sensitive: every rank must call this with
  the same list (the engine derives it from the plan, which is identical everywhere)."""
  if not dist.is_initialized():
    return
  world = list(range(dist.get_world_size()))
  for ranks in rank_lists:
    ranks = list(ranks)
    for c in range(copies):
      key = (tuple(ranks), c)
      if key in _GROUPS or len(ranks) <= 1:
        continue
      _GROUPS[key] = dist.group.WORLD if (ranks == world and c == 0) else dist.new_group(ranks)


def reset_groups() -> None:
  _GROUPS.clear()


class TorchBackend(object):
  """A communicator over ``ranks`` (global ranks, ordered)."""
  name = "torch"

  def __init__(self, ranks: Sequence[int], group=None, copy: int = 0):
    if not dist.is_initialized():
      raise RuntimeError("torch.distributed is not initialised; call epl.init() under a launcher "
                         "(torchrun / epl-launch) or use a single-rank communicator")
    self.ranks = list(ranks)
    self.size = len(self.ranks)
    me = dist.get_rank()
    self.rank = self.ranks.index(me) if me in self.ranks else -1
    if group is None:
      group = _GROUPS.get((tuple(self.ranks), copy))
    if group is None:
      if self.ranks == list(range(dist.get_world_size())) and copy == 0:
        group = dist.group.WORLD
      else:
        group = dist.new_group(self.ranks)
    self.group = group
    self._gloo = dist.get_backend(group) == "gloo" if self.rank >= 0 else False

  # -- helpers ------------------------------------------------------------------------
  def _g(self, group_rank: int) -> int:
    return self.ranks[group_rank]

  # -- verbs --------------------------------------------------------------------------
  def all_reduce(self, t: torch.Tensor, op: str = "sum") -> torch.Tensor:
    dist.all_reduce(t, op=_torch_op(op), group=self.group)
    return t

  def reduce(self, t: torch.Tensor, root: int = 0, op: str = "sum") -> torch.Tensor:
    dist.reduce(t, dst=self._g(root), op=_torch_op(op), group=self.group)
    return t

  def broadcast(self, t: torch.Tensor, root: int = 0) -> torch.Tensor:
    dist.broadcast(t, src=self._g(root), group=self.group)
    return t

  def all_gather(self, t: torch.Tensor) -> torch.Tensor:
    t = t.contiguous()
    out = t.new_empty((self.size * t.shape[0],) + tuple(t.shape[1:])) if t.dim() else t.new_empty((self.size,))
    dist.all_gather_into_tensor(out, t if t.dim() else t.reshape(1), group=self.group)
    return out

  def all_gatherv(self, t: torch.Tensor):
    """Variable first dimension.  Returns ``(concatenated, counts)``; counts stay on the
    device — no host round-trip in the data path (the reference blocks the host here,
    ``nccl_all_gather.cc:150-204``)."""
    t = t.contiguous()
    counts = self.all_gather(torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device))
    cmax = int(counts.max())          # one scalar sync, needed only to size the padded buffer
    pad = t.new_zeros((cmax,) + tuple(t.shape[1:]))
    pad[:t.shape[0]] = t
    gathered = self.all_gather(pad).reshape((self.size, cmax) + tuple(t.shape[1:]))
    pieces = [gathered[r, :int(counts[r])] for r in range(self.size)]
    return torch.cat(pieces, 0), counts

  def reduce_scatter(self, t: torch.Tensor, op: str = "sum") -> torch.Tensor:
    t = t.contiguous()
    if t.shape[0] % self.size:
      raise ValueError("reduce_scatter: dim 0 (%d) must be divisible by the communicator size (%d)" % (t.shape[0], self.size))
    n = t.shape[0] // self.size
    if self._gloo:
      full = t.clone()
      dist.all_reduce(full, op=_torch_op(op), group=self.group)
      return full[self.rank * n:(self.rank + 1) * n].clone()
    out = t.new_empty((n,) + tuple(t.shape[1:]))
    dist.reduce_scatter_tensor(out, t, op=_torch_op(op), group=self.group)
    return out

  def all_to_all(self, t: torch.Tensor) -> torch.Tensor:
    """Equal split of dim 0 into ``size`` segments; segment j goes to rank j."""
    t = t.contiguous()
    if t.shape[0] % self.size:
      raise ValueError("all_to_all: dim 0 (%d) must be divisible by the communicator size (%d)" % (t.shape[0], self.size))
    out = torch.empty_like(t)
    if self._gloo:
      ins = list(t.chunk(self.size, 0))
      outs = list(out.chunk(self.size, 0))
      self._p2p_exchange(ins, outs)
      return out
    dist.all_to_all_single(out, t, group=self.group)
    return out

  def all_to_allv(self, t: torch.Tensor, send_counts: torch.Tensor):
    """Rows ``[sum(send_counts[:j]), ...)`` go to rank j.  Returns ``(received, recv_counts)``."""
    t = t.contiguous()
    send_counts = send_counts.to(torch.int64)
    recv_counts = self.all_to_all(send_counts.to(t.device) if not self._gloo else send_counts.cpu()).cpu()
    sc = [int(c) for c in send_counts.cpu()]
    rc = [int(c) for c in recv_counts]
    out = t.new_empty((sum(rc),) + tuple(t.shape[1:]))
    if self._gloo:
      ins = list(t.split(sc, 0))
      outs = list(out.split(rc, 0))
      self._p2p_exchange(ins, outs)
    else:
      dist.all_to_all_single(out, t, output_split_sizes=rc, input_split_sizes=sc, group=self.group)
    return out, recv_counts

  def _p2p_exchange(self, ins: List[torch.Tensor], outs: List[torch.Tensor]) -> None:
    outs[self.rank].copy_(ins[self.rank])
    ops = []
    for peer in range(self.size):
      if peer == self.rank:
        continue
      if ins[peer].numel():
        ops.append(dist.P2POp(dist.isend, ins[peer].contiguous(), self._g(peer), group=self.group))
      if outs[peer].numel():
        ops.append(dist.P2POp(dist.irecv, outs[peer], self._g(peer), group=self.group))
    if ops:
      for w in dist.batch_isend_irecv(ops):
        w.wait()

  # -- in-place / asynchronous forms used by the engine's flat buckets -------------------
  def all_reduce_async(self, t: torch.Tensor, op: str = "sum"):
    return dist.all_reduce(t, op=_torch_op(op), group=self.group, async_op=True)

  def reduce_scatter_into(self, out: torch.Tensor, inp: torch.Tensor, op: str = "sum", async_op: bool = False):
    """``out`` may alias ``inp[rank*n:(rank+1)*n]`` (NCCL in-place form)."""
    if self._gloo:
      w = dist.all_reduce(inp, op=_torch_op(op), group=self.group, async_op=False)
      n = inp.numel() // self.size
      if out.data_ptr() != inp[self.rank * n:].data_ptr():
        out.copy_(inp[self.rank * n:(self.rank + 1) * n])
      return None
    return dist.reduce_scatter_tensor(out, inp, op=_torch_op(op), group=self.group, async_op=async_op)

  def all_gather_into(self, out: torch.Tensor, inp: torch.Tensor, async_op: bool = False):
    """``inp`` may alias ``out[rank*n:(rank+1)*n]``."""
    if self._gloo:
      inp = inp.clone()
    return dist.all_gather_into_tensor(out, inp, group=self.group, async_op=async_op)

  def send(self, t: torch.Tensor, dst: int):
    return dist.isend(t.contiguous(), self._g(dst), group=self.group)

  def recv(self, t: torch.Tensor, src: int):
    return dist.irecv(t, self._g(src), group=self.group)

  def barrier(self) -> None:
    dist.barrier(group=self.group)

  def close(self) -> None:
    self.group = None


def _no_grad_verbs(cls):
  """Collectives move raw data; autograd rules live in communicators/functional.py."""
  for name in ("all_reduce", "reduce", "broadcast", "all_gather", "all_gatherv", "reduce_scatter", "all_to_all", "all_to_allv",
               "all_reduce_async", "reduce_scatter_into", "all_gather_into", "send", "recv"):
    fn = getattr(cls, name)

    def wrapped(self, *a, __fn=fn, **kw):
      with torch.no_grad():
        a = tuple(x.detach() if isinstance(x, torch.Tensor) else x for x in a)
        return __fn(self, *a, **kw)
    wrapped.__name__ = name
    wrapped.__doc__ = fn.__doc__
    setattr(cls, name, wrapped)
  return cls


_no_grad_verbs(TorchBackend)


def make_backend(ranks: Sequence[int], prefer_native: bool = False, device: Optional[torch.device] = None, copy: int = 0):
  ranks = list(ranks)
  if len(ranks) <= 1:
    return LocalBackend(ranks or [0])
  if prefer_native and device is not None and device.type == "cuda":
    from easyparallellibrary_b200.communicators.native import NativeBackend
    return NativeBackend(ranks, device)
  return TorchBackend(ranks, copy=copy)
