"""The in-tree NCCL communicator (``csrc/communicator.cpp``) as a transport backend.

Bootstrap (reference: ``EplNcclCommunicatorGetId`` + TF collective broadcast of the id,
``nccl_communicator.cc:25-57``, ``nccl_ops.py:126-131``): group rank 0 creates the 128-byte NCCL
unique id and publishes it through the ``torch.distributed`` store; everyone calls
``ncclCommInitRank``.  Each backend owns a side stream; collectives are fenced with events
against the caller's current stream, never against the host.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence

import torch
import torch.distributed as dist

from easyparallellibrary_b200.runtime import native

_NCCL_DT = {torch.int8: 0, torch.uint8: 1, torch.int32: 2, torch.int64: 4, torch.float16: 6, torch.float32: 7,
            torch.float64: 8, torch.bfloat16: 9, torch.bool: 1}
_NCCL_OP = {"sum": 0, "prod": 1, "max": 2, "min": 3, "avg": 4}
_counter = {}          # per rank-set creation counter: identical on every member of the set
_LIVE: list = []            # every open NativeBackend of this process


def _find_nccl() -> str:
  base = os.path.dirname(os.path.dirname(torch.__file__))
  for cand in (os.path.join(base, "nvidia", "nccl", "lib", "libnccl.so.2"), os.path.join(os.path.dirname(torch.__file__), "lib", "libnccl.so.2"),
               "libnccl.so.2"):
    if os.path.sep not in cand or os.path.exists(cand):
      return cand
  return "libnccl.so.2"


class NativeBackend(object):
  name = "native"

  def __init__(self, ranks: Sequence[int], device: torch.device):
    self.lib = native.lib()
    if self.lib.epl_nccl_load(_find_nccl().encode()) != 0:
      raise RuntimeError("cannot load NCCL: %s" % self.lib.epl_comm_last_error().decode())
    self.ranks = list(ranks)
    self.size = len(self.ranks)
    me = dist.get_rank()
    self.rank = self.ranks.index(me)
    self.device = device
    rk = "-".join(map(str, self.ranks))
    _counter[rk] = _counter.get(rk, 0) + 1
    key = "epl_nccl_id/%s/%d" % (rk, _counter[rk])
    store = dist.distributed_c10d._get_default_store()
    ident = ctypes.create_string_buffer(128)
    if self.rank == 0:
      self._check(self.lib.epl_comm_get_unique_id(ident), "get_unique_id")
      store.set(key, ident.raw)
    else:
      ident.raw = bytes(store.get(key))[:128]
    h = ctypes.c_void_p()
    self._check(self.lib.epl_comm_create(ident, self.size, self.rank, device.index or 0, ctypes.byref(h)), "comm_create")
    self.handle = h
    self.stream_ptr = self.lib.epl_comm_stream(h)
    self.stream = torch.cuda.ExternalStream(self.stream_ptr, device=device)
    _LIVE.append(self)                      # the step watchdog aborts every live communicator on a hang

  def abort(self) -> None:
    """``ncclCommAbort``: unblocks kernels of this communicator that wait for a dead peer (used by runtime/watchdog.py;
    the reference wraps Abort but never calls it, tensorflow_nccl.h:119-123)."""
    if getattr(self, "handle", None):
      self.lib.epl_comm_abort(self.handle)

  def _check(self, rc: int, what: str) -> None:
    if rc != 0:
      raise RuntimeError("native communicator %s failed: %s" % (what, self.lib.epl_comm_last_error().decode()))

  @staticmethod
  def _cur() -> int:
    return torch.cuda.current_stream().cuda_stream

  def _p(self, t: torch.Tensor):
    return ctypes.c_void_p(t.data_ptr())

  def wait(self) -> None:
    """The caller's current stream waits (on the device) for everything issued so far."""
    self._check(self.lib.epl_comm_wait(self.handle, ctypes.c_void_p(self._cur())), "wait")

  # -- verbs (synchronous with respect to the caller's stream unless *_async) --------------------------
  def all_reduce(self, t: torch.Tensor, op: str = "sum") -> torch.Tensor:
    self.all_reduce_async(t, op)
    self.wait()
    return t

  def all_reduce_async(self, t: torch.Tensor, op: str = "sum"):
    self._check(self.lib.epl_comm_all_reduce(self.handle, self._p(t), self._p(t), ctypes.c_int64(t.numel()),
                                             _NCCL_DT[t.dtype], _NCCL_OP[op], ctypes.c_void_p(self._cur())), "all_reduce")
    return self

  def reduce(self, t: torch.Tensor, root: int = 0, op: str = "sum") -> torch.Tensor:
    self._check(self.lib.epl_comm_reduce(self.handle, self._p(t), self._p(t), ctypes.c_int64(t.numel()), _NCCL_DT[t.dtype],
                                         _NCCL_OP[op], root, ctypes.c_void_p(self._cur())), "reduce")
    self.wait()
    return t

  def broadcast(self, t: torch.Tensor, root: int = 0) -> torch.Tensor:
    self._check(self.lib.epl_comm_broadcast(self.handle, self._p(t), self._p(t), ctypes.c_int64(t.numel()), _NCCL_DT[t.dtype],
                                            root, ctypes.c_void_p(self._cur())), "broadcast")
    self.wait()
    return t

  def all_gather(self, t: torch.Tensor) -> torch.Tensor:
    t = t.contiguous()
    out = t.new_empty((self.size * t.shape[0],) + tuple(t.shape[1:])) if t.dim() else t.new_empty((self.size,))
    self.all_gather_into(out, t)
    return out

  def all_gather_into(self, out: torch.Tensor, inp: torch.Tensor, async_op: bool = False):
    self._check(self.lib.epl_comm_all_gather(self.handle, self._p(inp), self._p(out), ctypes.c_int64(inp.numel()),
                                             _NCCL_DT[inp.dtype], ctypes.c_void_p(self._cur())), "all_gather")
    if async_op:
      return self
    self.wait()
    return None

  def reduce_scatter(self, t: torch.Tensor, op: str = "sum") -> torch.Tensor:
    t = t.contiguous()
    n = t.shape[0] // self.size
    out = t.new_empty((n,) + tuple(t.shape[1:]))
    self.reduce_scatter_into(out, t, op)
    return out

  def reduce_scatter_into(self, out: torch.Tensor, inp: torch.Tensor, op: str = "sum", async_op: bool = False):
    self._check(self.lib.epl_comm_reduce_scatter(self.handle, self._p(inp), self._p(out), ctypes.c_int64(out.numel()),
                                                 _NCCL_DT[inp.dtype], _NCCL_OP[op], ctypes.c_void_p(self._cur())), "reduce_scatter")
    if async_op:
      return self
    self.wait()
    return None

  def all_to_all(self, t: torch.Tensor) -> torch.Tensor:
    t = t.contiguous()
    out = torch.empty_like(t)
    self._check(self.lib.epl_comm_all_to_all(self.handle, self._p(t), self._p(out), ctypes.c_int64(t.numel() // self.size),
                                             _NCCL_DT[t.dtype], ctypes.c_void_p(self._cur())), "all_to_all")
    self.wait()
    return out

  def all_to_allv(self, t: torch.Tensor, send_counts: torch.Tensor):
    t = t.contiguous()
    row = t.numel() // max(t.shape[0], 1) if t.shape[0] else int(torch.tensor(t.shape[1:]).prod()) if t.dim() > 1 else 1
    recv_counts = self.all_to_all(send_counts.to(t.device, torch.int64)).cpu()
    sc, rc = [int(c) for c in send_counts.cpu()], [int(c) for c in recv_counts]
    out = t.new_empty((sum(rc),) + tuple(t.shape[1:]))
    arr = lambda xs: (ctypes.c_int64 * self.size)(*xs)
    so = [sum(sc[:i]) * row for i in range(self.size)]
    ro = [sum(rc[:i]) * row for i in range(self.size)]
    self._check(self.lib.epl_comm_all_to_allv(self.handle, self._p(t), arr([c * row for c in sc]), arr(so), self._p(out),
                                              arr([c * row for c in rc]), arr(ro), _NCCL_DT[t.dtype], ctypes.c_void_p(self._cur())),
                "all_to_allv")
    self.wait()
    return out, recv_counts

  def all_gatherv(self, t: torch.Tensor, max_rows: Optional[int] = None):
    """Variable first dimension.  With ``max_rows`` given there is no host synchronisation at all: the result is
    a padded ``[size, max_rows, ...]`` tensor plus device-resident counts."""
    t = t.contiguous()
    cnt = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    counts = torch.empty(self.size, dtype=torch.int64, device=t.device)
    padded_mode = max_rows is not None
    if max_rows is None:
      allc = self.all_gather(cnt)
      max_rows = int(allc.max())
    pad = t.new_zeros((max_rows,) + tuple(t.shape[1:]))
    pad[:t.shape[0]] = t
    out = t.new_empty((self.size, max_rows) + tuple(t.shape[1:]))
    row = pad.numel() // max(max_rows, 1)
    self._check(self.lib.epl_comm_all_gatherv_padded(self.handle, self._p(pad), self._p(cnt), self._p(out), self._p(counts),
                                                     ctypes.c_int64(max_rows), ctypes.c_int64(row), _NCCL_DT[t.dtype],
                                                     ctypes.c_void_p(self._cur())), "all_gatherv")
    self.wait()
    if padded_mode:
      return out, counts
    return torch.cat([out[r, :int(c)] for r, c in enumerate(counts.cpu())], 0), counts

  def send(self, t: torch.Tensor, dst: int):
    self._check(self.lib.epl_comm_send(self.handle, self._p(t), ctypes.c_int64(t.numel()), _NCCL_DT[t.dtype], dst,
                                       ctypes.c_void_p(self._cur())), "send")
    return self

  def recv(self, t: torch.Tensor, src: int):
    self._check(self.lib.epl_comm_recv(self.handle, self._p(t), ctypes.c_int64(t.numel()), _NCCL_DT[t.dtype], src,
                                       ctypes.c_void_p(self._cur())), "recv")
    return self

  def barrier(self) -> None:
    self.all_reduce(torch.zeros(1, device=self.device))
    torch.cuda.current_stream().synchronize()

  def close(self) -> None:
    if getattr(self, "handle", None):
      torch.cuda.synchronize(self.device)       # nothing may still be queued on the side stream when it is destroyed
      self.stream = None
      self.lib.epl_comm_destroy(self.handle)
      self.handle = None
      if self in _LIVE:
        _LIVE.remove(self)
