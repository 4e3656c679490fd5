"""Python side of the fused tensor-parallel kernels (``csrc/gemm_tcgen05.cu``, ``epl_gemm_fused``).

A :class:`TPWorkspace` per split group owns the NVLink symmetric memory the kernels need:

* ``shard``  — where a rank exposes its token shard to its peers (all-gather source);
* ``stage``  — ``[world, rows_per_rank, N]`` slots that peers fill with their partial output tiles
  (reduce-scatter destination);
* two signal pads + device-local sync words (one set per kernel kind, epochs count launches).

Buffers grow on demand; growth is deterministic (same shapes on every rank), so the collective
handle exchange stays in lock-step.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import torch

from easyparallellibrary_b200.ops import _lib
from easyparallellibrary_b200.runtime.symmetric import SignalPad, SymmetricBuffer

COPY_CTAS = 20           # CTAs of the all-gather->GEMM grid that drive NVLink (148 - 20 run the GEMM)
_WS: Dict[int, "TPWorkspace"] = {}
_sig_ready = [False]


def _lib_fused():
  lib = _lib.require()
  if not _sig_ready[0]:
    p, i, u = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint
    lib.epl_gemm_fused.argtypes = [i, p, p, i, i, i, i, i, i, i, p, p, i, p, i, i, u, i, p, p, p, p, p, i, p]
    lib.epl_gemm_fused.restype = i
    _sig_ready[0] = True
  return lib


def available(group) -> bool:
  return torch.cuda.is_available() and 1 < group.size <= 8 and _lib.available()


class TPWorkspace(object):
  def __init__(self, group, device: torch.device):
    self.group, self.device = group, device
    pg = getattr(group.comm.primary, "group", None)
    self.pg = pg
    self.pad_ag = SignalPad(1, group.ranks, device, group=pg)
    self.pad_rs = SignalPad(1, group.ranks, device, group=pg)
    self.sync_ag = torch.zeros(32, dtype=torch.int32, device=device)
    self.sync_rs = torch.zeros(32, dtype=torch.int32, device=device)
    self.epoch_ag = 0
    self.epoch_rs = 0
    self.shard: Optional[SymmetricBuffer] = None
    self.stage: Optional[SymmetricBuffer] = None

  def shard_buffer(self, nbytes: int) -> SymmetricBuffer:
    if self.shard is None or self.shard.nbytes < nbytes:
      torch.cuda.synchronize(self.device)
      self.shard = SymmetricBuffer(max(nbytes, 1 << 20), self.group.ranks, self.device, group=self.pg)
    return self.shard

  def stage_buffer(self, nbytes: int) -> SymmetricBuffer:
    if self.stage is None or self.stage.nbytes < nbytes:
      torch.cuda.synchronize(self.device)
      self.stage = SymmetricBuffer(max(nbytes, 1 << 20), self.group.ranks, self.device, group=self.pg)
    return self.stage


def workspace(group, device) -> TPWorkspace:
  key = id(group)
  ws = _WS.get(key)
  if ws is None:
    ws = _WS[key] = TPWorkspace(group, device)
  return ws


def ag_gemm(x_shard2: torch.Tensor, w: torch.Tensor, group, bias=None, gelu: bool = False, b_mn_major: bool = False):
  """y = gather(x_shard) @ op(w) (+bias)(gelu) in one kernel.  Returns (y, pre, x_full)."""
  from easyparallellibrary_b200.ops import linear as L
  lib = _lib_fused()
  ws = workspace(group, x_shard2.device)
  rows, K = x_shard2.shape
  M = rows * group.size
  N = w.shape[1] if b_mn_major else w.shape[0]
  es = x_shard2.element_size()
  src = ws.shard_buffer(rows * K * es)
  src.tensor(x_shard2.dtype, rows * K).copy_(x_shard2.reshape(-1))          # expose the shard to the peers
  x_full = torch.empty((M, K), dtype=x_shard2.dtype, device=x_shard2.device)
  y = torch.empty((M, N), dtype=x_shard2.dtype, device=x_shard2.device)
  pre = torch.empty_like(y) if gelu else None
  epi = L.EPI_BIAS_GELU if gelu else (L.EPI_BIAS if bias is not None else L.EPI_NONE)
  rc = lib.epl_gemm_fused(1, x_full.data_ptr(), w.data_ptr(), M, N, K, K, w.stride(0), N, int(b_mn_major), _lib.ptr(bias),
                          _lib.ptr(pre), epi, y.data_ptr(), group.rank, group.size, 0, COPY_CTAS,      # epoch 0: kept on the device
                          ws.pad_ag.slot_table(0), ws.sync_ag.data_ptr(), src.peer_table(0), None, None,
                          int(x_shard2.dtype == torch.float16), _lib.stream())
  _lib.check(rc, "ag_gemm")
  return y, pre, x_full


def gemm_rs(a: torch.Tensor, w: torch.Tensor, group, b_mn_major: bool = False) -> torch.Tensor:
  """reduce_scatter_rows(a @ op(w)) in one kernel.  ``a``: [T, K]; returns [T/world, N]."""
  lib = _lib_fused()
  ws = workspace(group, a.device)
  M, K = a.shape
  N = w.shape[1] if b_mn_major else w.shape[0]
  rows = M // group.size
  stage = ws.stage_buffer(group.size * rows * N * 2)
  out = torch.empty((rows, N), dtype=a.dtype, device=a.device)
  rc = lib.epl_gemm_fused(2, a.data_ptr(), w.data_ptr(), M, N, K, a.stride(0), w.stride(0), N, int(b_mn_major), None, None, 0,
                          None, group.rank, group.size, 0, 0, ws.pad_rs.slot_table(0), ws.sync_rs.data_ptr(), None,
                          stage.peer_table(0), out.data_ptr(), 0, _lib.stream())
  _lib.check(rc, "gemm_rs")
  return out


def ag_weight_gemm(x2: torch.Tensor, w_shard: torch.Tensor, group, bias=None, gelu: bool = False, out_w_full: Optional[torch.Tensor] = None):
  """K2 — ZeRO-3 weight all-gather fused with the GEMM that consumes it.

  ``w_shard``: this rank's ``[N/world, K]`` rows of the weight.  Returns ``(y, pre, w_full)`` with
  ``y = x2 @ gather(w_shard)^T (+bias)(gelu)``; the gathered weight is kept for the backward GEMMs (written into
  ``out_w_full`` — the ZeRO-3 unit's transient buffer — when given).
  """
  from easyparallellibrary_b200.ops import linear as L
  lib = _lib_fused()
  ws = workspace(group, x2.device)
  if not hasattr(ws, "pad_agb"):
    ws.pad_agb = SignalPad(1, group.ranks, x2.device, group=ws.pg)
    ws.sync_agb = torch.zeros(32, dtype=torch.int32, device=x2.device)
    ws.epoch_agb = 0
    ws.wshard = None
  rows, K = w_shard.shape
  N = rows * group.size
  M = x2.shape[0]
  es = w_shard.element_size()
  nbytes = rows * K * es
  if ws.wshard is None or ws.wshard.nbytes < nbytes:
    torch.cuda.synchronize(x2.device)
    ws.wshard = SymmetricBuffer(max(nbytes, 1 << 20), group.ranks, x2.device, group=ws.pg)
  ws.wshard.tensor(w_shard.dtype, rows * K).copy_(w_shard.reshape(-1))
  if out_w_full is not None:
    if out_w_full.numel() != N * K or out_w_full.dtype != w_shard.dtype or not out_w_full.is_contiguous():
      raise ValueError("ag_weight_gemm: out_w_full must be a contiguous [N, K] buffer of the weight dtype")
    w_full = out_w_full.view(N, K)
  else:
    w_full = torch.empty((N, K), dtype=w_shard.dtype, device=x2.device)
  y = torch.empty((M, N), dtype=x2.dtype, device=x2.device)
  pre = torch.empty_like(y) if gelu else None
  epi = L.EPI_BIAS_GELU if gelu else (L.EPI_BIAS if bias is not None else L.EPI_NONE)
  rc = lib.epl_gemm_fused(3, x2.data_ptr(), w_full.data_ptr(), M, N, K, x2.stride(0), K, N, 0, _lib.ptr(bias), _lib.ptr(pre), epi,
                          y.data_ptr(), group.rank, group.size, 0, COPY_CTAS, ws.pad_agb.slot_table(0),
                          ws.sync_agb.data_ptr(), ws.wshard.peer_table(0), None, None, int(x2.dtype == torch.float16), _lib.stream())
  _lib.check(rc, "ag_weight_gemm")
  return y, pre, w_full
