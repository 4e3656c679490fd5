"""All-gather -> GEMM and GEMM -> reduce-scatter for ``epl.split`` tensor parallelism.

Reference: ``Replica2Split`` all-gather *then* a cuBLAS GEMM; backward GEMM *then* a
reduce-scatter — always two library calls in sequence (``bridging_layer.py:46-58``,
``distributed_dense.py:127-137``, ``nccl_ops.py:53-62``).

Here each pair is ONE sm_100a kernel (``csrc/gemm_tcgen05.cu`` with the ``ag_*`` / ``rs_*``
parameters): the persistent tcgen05 GEMM gets a few extra "copy" CTAs that pull the peers' token
shards over NVLink into a local buffer chunk by chunk and publish per-chunk flags that the TMA
producer waits on (all-gather -> GEMM); or its epilogue stores every output tile straight into the
owning rank's staging slot over NVLink, followed by an in-kernel cross-GPU barrier and the local
reduction (GEMM -> reduce-scatter).  Transfers therefore overlap the math tile by tile.

The NCCL + separate-GEMM path below is kept as the measured baseline (``fused=False``) and as the
CPU/gloo implementation.  Autograd pairs the two ops with each other: the backward of
all-gather->GEMM is a GEMM->reduce-scatter and vice versa.
"""
from __future__ import annotations


import torch

from easyparallellibrary_b200.ops.linear import linear

USE_FUSED = True         # flipped off for the baseline measurements


def _fused_ok(x: torch.Tensor, group) -> bool:
  if not (USE_FUSED and x.is_cuda and group.size > 1 and x.dtype == torch.bfloat16):
    return False
  try:
    from easyparallellibrary_b200.ops import tp_kernels
  except ImportError:
    return False
  return tp_kernels.available(group)


def _gemm_nt(x2, w, bias=None, gelu=False):
  """x2 @ w^T (+bias)(gelu) -> (y, pre)."""
  if x2.is_cuda and x2.dtype in (torch.bfloat16, torch.float16):
    from easyparallellibrary_b200.ops import linear as L
    if gelu:
      pre = torch.empty((x2.shape[0], w.shape[0]), dtype=x2.dtype, device=x2.device)
      return L.gemm(x2, w, bias=bias, epilogue=L.EPI_BIAS_GELU, pre=pre), pre
    return L.gemm(x2, w, bias=bias, epilogue=L.EPI_BIAS if bias is not None else L.EPI_NONE), None
  y = torch.nn.functional.linear(x2, w, bias)
  if gelu:
    return torch.nn.functional.gelu(y, approximate="tanh"), y
  return y, None


def _gemm_nn(a, w):
  """a[M,N] @ w[N,K] -> [M,K]."""
  if a.is_cuda and a.dtype in (torch.bfloat16, torch.float16):
    from easyparallellibrary_b200.ops import linear as L
    return L.gemm(a, w, b_mn_major=True)
  return a @ w


def _gemm_tn(a, b):
  """a[M,N]^T @ b[M,K] -> [N,K]."""
  if a.is_cuda and a.dtype in (torch.bfloat16, torch.float16):
    from easyparallellibrary_b200.ops import linear as L
    return L.gemm(a, b, a_mn_major=True, b_mn_major=True)
  return a.t() @ b


def _gelu_grad(pre, dy):
  if pre.is_cuda:
    from easyparallellibrary_b200.ops import _lib
    lib = _lib.require()
    out = torch.empty_like(dy)
    rc = lib.epl_gelu_bwd(pre.data_ptr(), dy.data_ptr(), out.data_ptr(), dy.numel(), _lib.dtype_code(dy.dtype), _lib.stream())
    _lib.check(rc, "gelu_bwd")
    return out
  x = pre.float().requires_grad_()
  with torch.enable_grad():
    torch.nn.functional.gelu(x, approximate="tanh").backward(dy.float())
  return x.grad.to(dy.dtype)


# ----------------------------------------------------------------------------------------------------
# primitive pairs
# ----------------------------------------------------------------------------------------------------
def ag_gemm(x_shard2: torch.Tensor, w: torch.Tensor, group, bias=None, gelu=False, b_mn_major=False):
  """Returns (y, pre, x_full).  y = gather(x_shard) @ op(w)."""
  if _fused_ok(x_shard2, group):
    from easyparallellibrary_b200.ops import tp_kernels
    return tp_kernels.ag_gemm(x_shard2, w, group, bias, gelu, b_mn_major)
  x_full = group.comm.allgather(x_shard2.contiguous()) if group.size > 1 else x_shard2
  if b_mn_major:
    return _gemm_nn(x_full, w), None, x_full
  y, pre = _gemm_nt(x_full, w, bias, gelu)
  return y, pre, x_full


def gemm_rs(a: torch.Tensor, w: torch.Tensor, group, b_mn_major=False) -> torch.Tensor:
  """reduce_scatter(a @ op(w)) over rows.  ``a``: [T, K]; result: [T/N, out]."""
  if _fused_ok(a, group):
    from easyparallellibrary_b200.ops import tp_kernels
    return tp_kernels.gemm_rs(a, w, group, b_mn_major)
  full = _gemm_nn(a, w) if b_mn_major else _gemm_nt(a, w)[0]
  return group.comm.reduce_scatter(full) if group.size > 1 else full


# ----------------------------------------------------------------------------------------------------
# autograd
# ----------------------------------------------------------------------------------------------------
class _AllGatherLinear(torch.autograd.Function):
  @staticmethod
  def forward(ctx, x_shard, w, bias, group, gelu):
    x2 = x_shard.reshape(-1, x_shard.shape[-1])
    y, pre, x_full = ag_gemm(x2, w, group, bias, gelu)
    ctx.save_for_backward(x_full, w, pre if pre is not None else x2.new_empty(0))
    ctx.group, ctx.gelu, ctx.has_bias, ctx.lead = group, gelu, bias is not None, x_shard.shape[:-1]
    return y

  @staticmethod
  def backward(ctx, dy):
    x_full, w, pre = ctx.saved_tensors
    dy = dy.contiguous()
    if ctx.gelu:
      dy = _gelu_grad(pre, dy)
    dx_shard = gemm_rs(dy, w, ctx.group, b_mn_major=True)            # (dY @ W) reduce-scattered over tokens
    dw = _gemm_tn(dy, x_full)
    db = dy.float().sum(0).to(dy.dtype) if ctx.has_bias else None
    return dx_shard.view(*ctx.lead, -1) if len(ctx.lead) > 1 else dx_shard, dw, db, None, None


class _LinearReduceScatter(torch.autograd.Function):
  @staticmethod
  def forward(ctx, x, w, group):
    x2 = x.reshape(-1, x.shape[-1]).contiguous()
    y = gemm_rs(x2, w, group)
    ctx.save_for_backward(x2, w)
    ctx.group = group
    return y

  @staticmethod
  def backward(ctx, dy_shard):
    x2, w = ctx.saved_tensors
    dx, _, dy_full = ag_gemm(dy_shard.contiguous(), w, ctx.group, b_mn_major=True)   # gather(dY) @ W
    dw = _gemm_tn(dy_full, x2)
    return dx, dw, None


def all_gather_linear(x_shard, w, bias, group, gelu: bool = False):
  if group.size == 1:
    return linear(x_shard, w, bias, gelu)
  return _AllGatherLinear.apply(x_shard, w, bias, group, gelu)


def linear_reduce_scatter(x, w, group):
  if group.size == 1:
    return linear(x, w, None)
  return _LinearReduceScatter.apply(x, w, group)
