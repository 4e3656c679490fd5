"""Symmetric memory: identically shaped device buffers on every rank of a group, each mapped into
every peer's address space over NVLink (``csrc/symm.cu``).

``SymmetricBuffer(nbytes, group_ranks)`` allocates locally, exchanges the CUDA IPC handles through
the control plane and imports the peers' buffers.  ``tensor(dtype, numel, offset)`` views the local
buffer as a ``torch.Tensor`` (zero-copy, via ``__cuda_array_interface__``); ``peer_table(offset)``
returns the ``void*[world]`` table a kernel needs.  ``SignalPad`` is a tiny symmetric buffer of
release/acquire flags used for device-side barriers.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from easyparallellibrary_b200.ops import _lib

_TYPESTR = {torch.float32: "<f4", torch.float16: "<f2", torch.bfloat16: "<V2", torch.int32: "<i4", torch.uint8: "|u1",
            torch.int64: "<i8"}


class _CudaArray(object):
  def __init__(self, ptr: int, nbytes: int):
    self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def wrap_pointer(ptr: int, nbytes: int, dtype: torch.dtype, device: torch.device) -> torch.Tensor:
  raw = torch.as_tensor(_CudaArray(ptr, nbytes), device=device)
  return raw.view(dtype)


def _sym_lib():
  lib = _lib.require()
  if not hasattr(lib, "_symm_ready"):
    lib.epl_symm_alloc.argtypes = [ctypes.c_int64, ctypes.POINTER(ctypes.c_void_p)]
    lib.epl_symm_free.argtypes = [ctypes.c_void_p]
    lib.epl_symm_export.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    lib.epl_symm_import.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]
    lib.epl_symm_unimport.argtypes = [ctypes.c_void_p]
    lib.epl_symm_barrier.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_uint, ctypes.c_void_p]
    lib.epl_peer_copy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]
    lib.epl_fused_rs_adam_ag.argtypes = ([ctypes.c_void_p] * 8 + [ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_uint, ctypes.c_int] + [ctypes.c_float] * 8 + [ctypes.c_int, ctypes.c_void_p])
    lib.epl_fused_rs_adam_ag_v2.argtypes = ([ctypes.c_void_p] * 8 + [ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_uint, ctypes.c_int, ctypes.c_void_p] + [ctypes.c_float] * 4 +
                                            [ctypes.c_int, ctypes.c_void_p])
    lib._symm_ready = True
  return lib


_EXCHANGES: dict = {}     # per rank-set exchange counter: identical on every member (buffers are created in program order)


def _exchange_handles(raw: bytes, ranks: Sequence[int], my_index: int, group=None) -> List[bytes]:
  """All-gather of the 64-byte IPC handles among ``ranks``.

  Goes through the rendezvous store (each member publishes under a key unique to (rank set, creation index, member) and
  reads the others'), NOT through a process-group collective: the rank set is often a strict subset of the world (the
  data-parallel group of one pipeline stage, one of several tensor-parallel groups) whose communicator is an in-tree NCCL
  communicator without a torch ProcessGroup, and a WORLD collective would need every rank of the job to create the same
  number of buffers at the same time.  Falls back to ``all_gather_object`` over ``group`` when no store is available."""
  tag = "-".join(str(int(r)) for r in ranks)
  n = _EXCHANGES.get(tag, 0)
  _EXCHANGES[tag] = n + 1
  store = None
  try:
    store = dist.distributed_c10d._get_default_store()
  except Exception:  # pragma: no cover
    store = None
  if store is None:                                      # pragma: no cover
    out: List[Optional[bytes]] = [None] * len(ranks)
    dist.all_gather_object(out, raw, group=group)
    return out  # type: ignore[return-value]
  base = "epl_symm/%s/%d/" % (tag, n)
  store.set(base + str(my_index), raw)
  return [raw if i == my_index else bytes(store.get(base + str(i))) for i in range(len(ranks))]


class SymmetricBuffer(object):
  def __init__(self, nbytes: int, ranks: Sequence[int], device: torch.device, group=None):
    self.lib = _sym_lib()
    self.nbytes = (int(nbytes) + 255) // 256 * 256
    self.ranks = list(ranks)
    self.world = len(self.ranks)
    self.rank = self.ranks.index(dist.get_rank()) if dist.is_initialized() else 0
    self.device = device
    p = ctypes.c_void_p()
    rc = self.lib.epl_symm_alloc(self.nbytes, ctypes.byref(p))
    if rc:
      raise RuntimeError("symmetric alloc of %d bytes failed (cuda error %d)" % (self.nbytes, rc))
    self.local_ptr = p.value
    self.peer_ptrs: List[int] = [0] * self.world
    self.peer_ptrs[self.rank] = self.local_ptr
    self._imported: List[int] = []
    if self.world > 1:
      h = ctypes.create_string_buffer(64)
      rc = self.lib.epl_symm_export(ctypes.c_void_p(self.local_ptr), h)
      if rc:
        raise RuntimeError("cudaIpcGetMemHandle failed (%d)" % rc)
      handles = _exchange_handles(h.raw, self.ranks, self.rank, group)
      for r, raw in enumerate(handles):
        if r == self.rank:
          continue
        q = ctypes.c_void_p()
        rc = self.lib.epl_symm_import(ctypes.create_string_buffer(raw, 64), ctypes.byref(q))
        if rc:
          raise RuntimeError("cudaIpcOpenMemHandle for peer %d failed (%d): NVLink peer access unavailable?" % (r, rc))
        self.peer_ptrs[r] = q.value
        self._imported.append(q.value)

  def tensor(self, dtype: torch.dtype, numel: int, byte_offset: int = 0) -> torch.Tensor:
    key = (dtype, numel, byte_offset)
    views = self.__dict__.setdefault("_views", {})
    t = views.get(key)                 # wrapping a raw pointer costs ~50 us of host time: the hot ops ask for the same views every call
    if t is None:
      nbytes = numel * torch.empty(0, dtype=dtype).element_size()
      if byte_offset + nbytes > self.nbytes:
        raise ValueError("view exceeds the symmetric buffer")
      t = wrap_pointer(self.local_ptr + byte_offset, nbytes, dtype, self.device)
      t._epl_symm_owner = self          # keep the allocation alive as long as a view exists
      if len(views) < 64:
        views[key] = t
    return t

  def peer_table(self, byte_offset: int = 0):
    tables = self.__dict__.setdefault("_tables", {})
    arr = tables.get(byte_offset)
    if arr is None:
      arr = (ctypes.c_void_p * 8)()
      for r in range(self.world):
        arr[r] = self.peer_ptrs[r] + byte_offset
      if len(tables) < 256:
        tables[byte_offset] = arr
    return arr

  def close(self) -> None:
    for q in self._imported:
      self.lib.epl_symm_unimport(ctypes.c_void_p(q))
    self._imported = []
    self.__dict__.pop("_views", None)
    self.__dict__.pop("_tables", None)
    if self.local_ptr:
      self.lib.epl_symm_free(ctypes.c_void_p(self.local_ptr))
      self.local_ptr = 0


class SignalPad(SymmetricBuffer):
  """``slots`` independent flag groups of 2 x 8 uint32 each (start / end barrier per bucket)."""
  SLOT_BYTES = 2 * 8 * 4

  def __init__(self, slots: int, ranks: Sequence[int], device: torch.device, group=None):
    super().__init__(max(slots, 1) * self.SLOT_BYTES, ranks, device, group)
    self.epoch = 0

  def slot_table(self, slot: int):
    return self.peer_table(slot * self.SLOT_BYTES)

  def barrier(self, slot: int = 0) -> None:
    """Device-side barrier on the current stream (no host involvement)."""
    self.epoch += 1
    rc = self.lib.epl_symm_barrier(self.slot_table(slot), self.rank, self.world, self.epoch, _lib.stream())
    _lib.check(rc, "symm_barrier")
