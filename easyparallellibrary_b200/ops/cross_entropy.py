"""Fused softmax cross-entropy (``csrc/xent.cu``): one read of the logits, gradient written in place."""
from __future__ import annotations

import torch

from easyparallellibrary_b200.ops import _lib


class _XentFn(torch.autograd.Function):
  @staticmethod
  def forward(ctx, logits, labels, ignore_index, reduction):
    lib = _lib.require()
    V = logits.shape[-1]
    l2 = logits.view(-1, V)
    if not l2.is_contiguous():
      l2 = l2.contiguous()
    rows = l2.shape[0]
    lab = labels.reshape(-1).to(torch.int64).contiguous()
    loss = torch.empty(rows, dtype=torch.float32, device=logits.device)
    valid = (lab != ignore_index).sum().clamp(min=1).float() if reduction == "mean" else None
    # gradient is produced now (scaled for mean reduction by a device scalar later) and overwrites the logits
    rc = lib.epl_xent(l2.data_ptr(), lab.data_ptr(), loss.data_ptr(), l2.data_ptr(), None, None, rows, V, l2.stride(0),
                      1.0, ignore_index, 0, 0, _lib.dtype_code(l2.dtype), _lib.stream())
    _lib.check(rc, "xent")
    ctx.save_for_backward(l2, valid if valid is not None else loss.new_ones(()))
    ctx.shape, ctx.reduction = logits.shape, reduction
    if reduction == "mean":
      return loss.sum() / valid
    if reduction == "sum":
      return loss.sum()
    return loss.view(labels.shape)

  @staticmethod
  def backward(ctx, gout):
    lib = _lib.require()
    dl, valid = ctx.saved_tensors
    if ctx.reduction == "none":
      return (dl.view(ctx.shape) * gout.reshape(-1, 1).to(dl.dtype).view(*ctx.shape[:-1], 1)), None, None, None
    scale = (gout.float() / valid).reshape(1).contiguous() if ctx.reduction == "mean" else gout.float().reshape(1).contiguous()
    rc = lib.epl_scale_by_device_scalar(dl.data_ptr(), scale.data_ptr(), dl.numel(), _lib.dtype_code(dl.dtype), _lib.stream())
    _lib.check(rc, "xent_scale")
    return dl.view(ctx.shape), None, None, None


def softmax_cross_entropy(logits: torch.Tensor, labels: torch.Tensor, ignore_index: int = -100,
                          reduction: str = "mean") -> torch.Tensor:
  """Note: on CUDA the logits buffer is consumed (overwritten with its gradient)."""
  from easyparallellibrary_b200.runtime import amp
  if amp.o1_active():                            # O1 deny-list op
    logits = amp.cast_args("cross_entropy", logits)
  V = logits.shape[-1]
  if logits.is_cuda and V * logits.element_size() <= 200 * 1024 and logits.stride(-1) == 1 \
      and (logits.numel() // V == 0 or logits.view(-1, V).stride(0) % (16 // logits.element_size()) == 0):
    return _XentFn.apply(logits, labels, ignore_index, reduction)
  return torch.nn.functional.cross_entropy(logits.float().view(-1, V), labels.reshape(-1), ignore_index=ignore_index,
                                           reduction=reduction)
