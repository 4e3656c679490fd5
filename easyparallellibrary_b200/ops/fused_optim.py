"""Fused optimizer kernels (``csrc/optim.cu``)."""
from __future__ import annotations

from typing import Optional

import torch

from easyparallellibrary_b200.ops import _lib


def adamw_step(master: torch.Tensor, grad: torch.Tensor, m: torch.Tensor, v: torch.Tensor, step: int, hyper,
               grad_scale: float = 1.0, decay_mask: Optional[torch.Tensor] = None,
               model_out: Optional[torch.Tensor] = None, dyn: Optional[torch.Tensor] = None) -> None:
  """``dyn``: device tensor {lr, 1/(1-b1^t), 1/(1-b2^t), grad scale}; when given the launch carries no per-step value
  (CUDA-graph replays read the current ones from memory)."""
  lib = _lib.require()
  assert master.dtype == torch.float32 and m.dtype == torch.float32 and v.dtype == torch.float32
  assert master.is_contiguous() and grad.is_contiguous() and grad.numel() == master.numel()
  if dyn is not None:
    out_dt = _lib.dtype_code(model_out.dtype) if model_out is not None else _lib.F32
    rc = lib.epl_adamw_dyn(master.data_ptr(), grad.data_ptr(), _lib.dtype_code(grad.dtype), m.data_ptr(), v.data_ptr(),
                           _lib.ptr(model_out), out_dt, _lib.ptr(decay_mask), master.numel(), dyn.data_ptr(), hyper.beta1,
                           hyper.beta2, hyper.eps, hyper.weight_decay, _lib.stream())
    _lib.check(rc, "adamw")
    return
  if hyper.bias_correction:
    inv_c1 = 1.0 / (1.0 - hyper.beta1 ** step)
    inv_c2 = 1.0 / (1.0 - hyper.beta2 ** step)
  else:
    inv_c1 = inv_c2 = 1.0
  out_dt = _lib.dtype_code(model_out.dtype) if model_out is not None else _lib.F32
  rc = lib.epl_adamw(master.data_ptr(), grad.data_ptr(), _lib.dtype_code(grad.dtype), m.data_ptr(), v.data_ptr(),
                     _lib.ptr(model_out), out_dt, _lib.ptr(decay_mask), master.numel(), hyper.lr, hyper.beta1,
                     hyper.beta2, hyper.eps, hyper.weight_decay, grad_scale, inv_c1, inv_c2, _lib.stream())
  _lib.check(rc, "adamw")


def sgd_step(master: torch.Tensor, grad: torch.Tensor, mom: Optional[torch.Tensor], hyper, grad_scale: float = 1.0,
             model_out: Optional[torch.Tensor] = None) -> None:
  lib = _lib.require()
  out_dt = _lib.dtype_code(model_out.dtype) if model_out is not None else _lib.F32
  rc = lib.epl_sgd(master.data_ptr(), grad.data_ptr(), _lib.dtype_code(grad.dtype), _lib.ptr(mom), _lib.ptr(model_out),
                   out_dt, master.numel(), hyper.lr, hyper.momentum, hyper.weight_decay, grad_scale, _lib.stream())
  _lib.check(rc, "sgd")


def sumsq_and_finite(x: torch.Tensor, out2: Optional[torch.Tensor] = None) -> torch.Tensor:
  """Returns a 2-float tensor: [sum of squares, 1.0 if any non-finite else 0.0] (accumulates into ``out2``)."""
  lib = _lib.require()
  if out2 is None:
    out2 = torch.zeros(2, dtype=torch.float32, device=x.device)
  rc = lib.epl_sumsq(x.data_ptr(), _lib.dtype_code(x.dtype), x.numel(), out2.data_ptr(), _lib.stream())
  _lib.check(rc, "sumsq")
  return out2
