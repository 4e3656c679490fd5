"""LayerNorm / RMSNorm with hand-written sm_100a forward/backward (``csrc/norm.cu``)."""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from easyparallellibrary_b200.ops import _lib


class _NormFn(torch.autograd.Function):
  @staticmethod
  def forward(ctx, x, gamma, beta, eps, rms, fork=False):
    lib = _lib.require()
    D = x.shape[-1]
    x2 = x.contiguous().view(-1, D)
    rows = x2.shape[0]
    y = torch.empty_like(x2)
    mean = torch.empty(rows, dtype=torch.float32, device=x.device) if not rms else None
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
    rc = lib.epl_norm_fwd(x2.data_ptr(), gamma.data_ptr(), _lib.ptr(beta), y.data_ptr(), _lib.ptr(mean), rstd.data_ptr(),
                          rows, D, eps, _lib.dtype_code(x.dtype), int(rms), _lib.stream())
    _lib.check(rc, "norm_fwd")
    ctx.save_for_backward(x2, gamma, mean if mean is not None else rstd, rstd)
    ctx.rms, ctx.has_beta, ctx.shape, ctx.fork = rms, beta is not None, x.shape, fork
    if fork:
      return x.view_as(x), y.view(x.shape)
    return y.view(x.shape)

  @staticmethod
  def backward(ctx, *grads):
    dy, dres = (grads[1], grads[0]) if ctx.fork else (grads[0], None)
    lib = _lib.require()
    x2, gamma, mean, rstd = ctx.saved_tensors
    rows, D = x2.shape
    if dy is None:
      dy = torch.zeros(ctx.shape, dtype=x2.dtype, device=x2.device)
    dy2 = dy.contiguous().view(rows, D)
    if dres is not None:
      dres = dres.contiguous()
    dx = torch.empty_like(x2)
    dgamma = torch.empty_like(gamma)
    dbeta = torch.empty_like(gamma) if ctx.has_beta else None
    grid = lib.epl_norm_bwd_grid(rows)
    ws = torch.empty(2 * grid * D, dtype=torch.float32, device=x2.device)
    rc = lib.epl_norm_bwd(x2.data_ptr(), dy2.data_ptr(), gamma.data_ptr(), None if ctx.rms else mean.data_ptr(),
                          rstd.data_ptr(), dx.data_ptr(), dgamma.data_ptr(), _lib.ptr(dbeta), ws.data_ptr(), rows, D,
                          _lib.dtype_code(x2.dtype), int(ctx.rms), 0, _lib.ptr(dres), _lib.stream())
    _lib.check(rc, "norm_bwd")
    _lib.launches += 2
    return dx.view(ctx.shape), dgamma, dbeta, None, None, None


def _o1(x, gamma, beta):
  from easyparallellibrary_b200.runtime import amp
  if amp.o1_active():                            # O1 deny-list op: statistics and output in fp32
    return amp.cast_args("layer_norm", x, gamma, beta)
  return x, gamma, beta


def layer_norm(x: torch.Tensor, gamma: torch.Tensor, beta: Optional[torch.Tensor], eps: float = 1e-5) -> torch.Tensor:
  x, gamma, beta = _o1(x, gamma, beta)
  if x.is_cuda and x.shape[-1] % (16 // x.element_size()) == 0:
    return _NormFn.apply(x, gamma, beta, eps, False)
  return torch.nn.functional.layer_norm(x, (x.shape[-1],), gamma, beta, eps)


def layer_norm_fork(x: torch.Tensor, gamma: torch.Tensor, beta: Optional[torch.Tensor], eps: float = 1e-5):
  """Returns ``(x, LN(x))`` for a pre-LN residual block.  The backward kernel adds the gradient arriving on the
  skip connection to the LayerNorm input gradient in the same pass (no stand-alone add kernel)."""
  x, gamma, beta = _o1(x, gamma, beta)
  if x.is_cuda and x.shape[-1] % (16 // x.element_size()) == 0:
    return _NormFn.apply(x, gamma, beta, eps, False, True)
  return x, torch.nn.functional.layer_norm(x, (x.shape[-1],), gamma, beta, eps)


def rms_norm(x: torch.Tensor, gamma: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
  x, gamma, _ = _o1(x, gamma, None)
  if x.is_cuda and x.shape[-1] % (16 // x.element_size()) == 0:
    return _NormFn.apply(x, gamma, None, eps, True)
  xf = x.float()
  return (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).to(x.dtype) * gamma


class LayerNorm(nn.Module):
  def __init__(self, dim: int, eps: float = 1e-5, bias: bool = True):
    super().__init__()
    self.weight = nn.Parameter(torch.ones(dim))
    self.bias = nn.Parameter(torch.zeros(dim)) if bias else None
    self.eps = eps

  def reset_parameters(self):
    nn.init.ones_(self.weight)
    if self.bias is not None:
      nn.init.zeros_(self.bias)

  def forward(self, x):
    return layer_norm(x, self.weight, self.bias, self.eps)

  def fork(self, x):
    return layer_norm_fork(x, self.weight, self.bias, self.eps)


class RMSNorm(nn.Module):
  def __init__(self, dim: int, eps: float = 1e-6):
    super().__init__()
    self.weight = nn.Parameter(torch.ones(dim))
    self.eps = eps

  def reset_parameters(self):
    nn.init.ones_(self.weight)

  def forward(self, x):
    return rms_norm(x, self.weight, self.eps)
