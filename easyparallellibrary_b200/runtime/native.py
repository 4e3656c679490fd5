"""ctypes bindings for ``lib/libepl_runtime.so`` (``csrc/runtime.cpp`` + ``csrc/communicator.cpp``)."""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Sequence, Tuple

_HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_PATH = os.path.join(_HERE, "lib", "libepl_runtime.so")
_lib: Optional[ctypes.CDLL] = None

POLICIES = {"preferforward": 0, "preferbackward": 1, "preferbackwardoptimizer": 2}


def available() -> bool:
  return os.path.exists(_PATH)


def lib() -> ctypes.CDLL:
  global _lib
  if _lib is None:
    if not os.path.exists(_PATH):
      from easyparallellibrary_b200.build import build_all
      build_all()
    _lib = ctypes.CDLL(_PATH)
    _lib.epl_comm_last_error.restype = ctypes.c_char_p
    _lib.epl_comm_stream.restype = ctypes.c_void_p
    _lib.epl_schedule_simulate.argtypes = [ctypes.c_int] * 4 + [ctypes.c_double] * 3 + [ctypes.POINTER(ctypes.c_double)] * 2 + [ctypes.POINTER(ctypes.c_int32)]
  return _lib


def schedule_stage(policy: str, stage: int, num_stages: int, num_micro_batch: int, prefetch: int = 1) -> List[Tuple[int, int]]:
  cap = 8 * num_micro_batch + 8
  buf = (ctypes.c_int32 * (2 * cap))()
  n = lib().epl_schedule_stage(POLICIES[policy.lower()], stage, num_stages, num_micro_batch, prefetch, buf, cap)
  if n < 0:
    raise RuntimeError("schedule buffer too small")
  return [(buf[2 * i], buf[2 * i + 1]) for i in range(n)]


def schedule_simulate(policy: str, num_stages: int, num_micro_batch: int, prefetch: int = 1, t_fwd: float = 1.0,
                      t_bwd: float = 2.0, t_p2p: float = 0.0):
  mk, bub = ctypes.c_double(), ctypes.c_double()
  infl = (ctypes.c_int32 * num_stages)()
  rc = lib().epl_schedule_simulate(POLICIES[policy.lower()], num_stages, num_micro_batch, prefetch, t_fwd, t_bwd, t_p2p,
                                   ctypes.byref(mk), ctypes.byref(bub), infl)
  return rc == 0, mk.value, bub.value, list(infl)


def plan_buckets(nbytes: Sequence[int], dtype_ids: Sequence[int], max_splits: int) -> List[List[int]]:
  n = len(nbytes)
  if n == 0:
    return []
  a = (ctypes.c_int64 * n)(*nbytes)
  d = (ctypes.c_int32 * n)(*dtype_ids)
  out = (ctypes.c_int32 * n)()
  nb = lib().epl_plan_buckets(a, d, n, max_splits, out)
  buckets: List[List[int]] = [[] for _ in range(nb)]
  for i in range(n):
    buckets[out[i]].append(i)
  return [b for b in buckets if b]


def partition_stages(weights: Sequence[float], parts: int) -> List[int]:
  n = len(weights)
  w = (ctypes.c_double * n)(*[max(float(x), 1e-12) for x in weights])
  out = (ctypes.c_int32 * (parts + 1))()
  if lib().epl_partition_stages(w, n, parts, out) != 0:
    raise ValueError("partition_stages requires parts >= 1")
  return list(out)
