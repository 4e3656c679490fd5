"""NVLS: NVSwitch multicast / in-switch reduction.

A :class:`MulticastBuffer` is a symmetric allocation that is ADDITIONALLY bound to a multicast object
(``cuMemCreate`` + ``cuMulticastCreate`` / ``cuMulticastBindMem``): one virtual address (``multicast_ptr``) that fans a
store out to every rank's copy and lets ``multimem.ld_reduce`` return the switch-side sum of all copies.  The VMM /
multicast plumbing (POSIX file-descriptor hand-off between the processes included) is ``torch.distributed``'s
symmetric-memory rendezvous — control plane, like the process group itself; the data plane is the in-tree kernel
``csrc/symm.cu::nvls_allreduce_kernel`` (``multimem.ld_reduce`` + ``multimem.st``, SASS ``LDGMC.E.HPADD`` for the reduce-load; the multicast store is a ``STG.E.128.STRONG.SYS`` on the multicast address).

Availability is a property of the platform (NVSwitch fabric + driver + container permissions): ``MulticastBuffer.supported``
tells; callers fall back to the peer-pointer path (``runtime/symmetric.py``) when it is False.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch
import torch.distributed as dist

from easyparallellibrary_b200.ops import _lib
from easyparallellibrary_b200.runtime.symmetric import SignalPad, _sym_lib


class MulticastBuffer(object):
  def __init__(self, nbytes: int, device: torch.device, group=None):
    import torch.distributed._symmetric_memory as symm_mem
    self.group = group or dist.group.WORLD
    self.device = device
    self.nbytes = (int(nbytes) + 255) // 256 * 256
    name = self.group.group_name
    try:
      if hasattr(symm_mem, "is_symm_mem_enabled_for_group") and not symm_mem.is_symm_mem_enabled_for_group(name):
        symm_mem.enable_symm_mem_for_group(name)
    except Exception:          # newer torch enables groups lazily
      pass
    self.storage = symm_mem.empty(self.nbytes, dtype=torch.uint8, device=device)
    self.storage.zero_()
    self.handle = symm_mem.rendezvous(self.storage, self.group)
    self.rank, self.world = self.handle.rank, self.handle.world_size
    self.multicast_ptr = int(self.handle.multicast_ptr or 0)
    self.supported = self.multicast_ptr != 0
    ranks = dist.get_process_group_ranks(self.group)
    self.pad = SignalPad(1, ranks, device, group=self.group)
    self.sync = torch.zeros(4, dtype=torch.int32, device=device)
    lib = _sym_lib()
    if not hasattr(lib, "_nvls_ready"):
      lib.epl_nvls_allreduce.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_uint, ctypes.c_int, ctypes.c_void_p]
      lib._nvls_ready = True
    self.lib = lib

  def tensor(self, dtype: torch.dtype, numel: int, byte_offset: int = 0) -> torch.Tensor:
    return self.storage[byte_offset:byte_offset + numel * torch.empty(0, dtype=dtype).element_size()].view(dtype)

  # -- drop-in for runtime.symmetric.SymmetricBuffer (the peer-pointer kernels work on the same allocation) ------------------
  @property
  def peer_ptrs(self):
    return [int(p) for p in self.handle.buffer_ptrs]

  @property
  def local_ptr(self) -> int:
    return int(self.handle.buffer_ptrs[self.rank])

  def peer_table(self, byte_offset: int = 0):
    tables = self.__dict__.setdefault("_tables", {})
    arr = tables.get(byte_offset)
    if arr is None:
      arr = (ctypes.c_void_p * 8)()
      for r, p in enumerate(self.handle.buffer_ptrs):
        arr[r] = int(p) + byte_offset
      if len(tables) < 256:
        tables[byte_offset] = arr
    return arr

  def close(self) -> None:
    self.storage = None

  def all_reduce_(self, dtype: torch.dtype, numel: int, blocks: int = 32) -> torch.Tensor:
    """In-place sum over the ranks of the first ``numel`` elements (bf16 or fp32; byte count a multiple of 16)."""
    if not self.supported:
      raise RuntimeError("NVLS multicast is not available on this platform")
    nbytes = numel * torch.empty(0, dtype=dtype).element_size()
    rc = self.lib.epl_nvls_allreduce(self.multicast_ptr, self.pad.slot_table(0), self.sync.data_ptr(), nbytes, _lib.dtype_code(dtype),
                                     self.rank, self.world, 0, blocks, _lib.stream())
    _lib.check(rc, "nvls_allreduce")
    return self.tensor(dtype, numel)


def probe(device: torch.device, group=None) -> Optional[MulticastBuffer]:
  """A small multicast buffer, or None when the platform has no NVLS (the reason is logged)."""
  from easyparallellibrary_b200.utils.logging import get_logger
  try:
    buf = MulticastBuffer(1 << 20, device, group)
  except Exception as e:
    get_logger().info("NVLS probe failed: %s", e)
    return None
  if not buf.supported:
    get_logger().info("NVLS probe: symmetric memory works but the multicast pointer is null (no NVSwitch multicast in this container)")
  return buf
