"""Rotary position embedding on a packed QKV tensor (``csrc/rope.cu``), in place, with the exact inverse as backward."""
from __future__ import annotations

import ctypes

import torch

from easyparallellibrary_b200.ops import _lib

_ready = [False]


def _lib_rope():
  lib = _lib.require()
  if not _ready[0]:
    lib.epl_rope.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                             ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    _ready[0] = True
  return lib


def rope_reference(qkv: torch.Tensor, base: float = 10000.0, pos_offset: int = 0) -> torch.Tensor:
  """Plain PyTorch (rotate-half convention): Q and K slices rotated, V untouched."""
  B, S, _, H, D = qkv.shape
  half = D // 2
  freq = base ** (-torch.arange(0, D, 2, device=qkv.device, dtype=torch.float32) / D)
  ang = (torch.arange(S, device=qkv.device, dtype=torch.float32) + pos_offset)[:, None] * freq[None, :]
  cs, sn = ang.cos()[None, :, None, None, :], ang.sin()[None, :, None, None, :]
  x = qkv.float()
  lo, hi = x[..., :half], x[..., half:]
  rot = torch.cat([lo * cs - hi * sn, hi * cs + lo * sn], -1)
  out = x.clone()
  out[:, :, :2] = rot[:, :, :2]
  return out.to(qkv.dtype)


class _Rope(torch.autograd.Function):
  @staticmethod
  def forward(ctx, qkv, base, pos_offset):
    lib = _lib_rope()
    B, S, _, H, D = qkv.shape
    out = qkv.contiguous().clone()
    rc = lib.epl_rope(out.data_ptr(), B, S, H, D, base, 1.0, pos_offset, _lib.dtype_code(out.dtype), _lib.stream())
    _lib.check(rc, "rope")
    ctx.args = (base, pos_offset)
    return out

  @staticmethod
  def backward(ctx, g):
    lib = _lib_rope()
    B, S, _, H, D = g.shape
    g = g.contiguous().clone()
    rc = lib.epl_rope(g.data_ptr(), B, S, H, D, ctx.args[0], -1.0, ctx.args[1], _lib.dtype_code(g.dtype), _lib.stream())
    _lib.check(rc, "rope_bwd")
    return g, None, None


def apply_rope(qkv: torch.Tensor, base: float = 10000.0, pos_offset: int = 0) -> torch.Tensor:
  D = qkv.shape[-1]
  if qkv.is_cuda and (D // 2) % (16 // qkv.element_size()) == 0:
    return _Rope.apply(qkv, base, pos_offset)
  return rope_reference(qkv, base, pos_offset)
