"""Causal / bidirectional multi-head attention.

``attention(q, k, v, causal)`` takes ``[B, H, S, D]`` tensors.  On sm_100a it runs the hand-written
flash-style kernels in ``csrc/attention.cu`` when they are built for the head dimension; the
``torch`` SDPA call is the numerics reference and the CPU path.
"""
from __future__ import annotations

import math

import torch

from easyparallellibrary_b200.ops import _lib

import os

# "auto" | "epl" | "sdpa".  auto: the hand-written tcgen05 kernel whenever it supports the input (bf16, head dim 64 —
# measured faster end to end than the cuDNN path on GPT-2-XL because it consumes the packed QKV layout directly),
# else torch SDPA.  EPL_ATTENTION=epl makes an unsupported input an error, =sdpa forces the library path.
_IMPL = os.environ.get("EPL_ATTENTION", "auto")


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, causal: bool = True, dropout_p: float = 0.0) -> torch.Tensor:
  if _IMPL != "sdpa" and q.is_cuda and dropout_p == 0.0:
    try:
      from easyparallellibrary_b200.ops import attention_kernel
      if attention_kernel.supported(q, k, v):
        return attention_kernel.flash_attention(q, k, v, causal)
    except ImportError:
      if _IMPL == "epl":
        raise
  return torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=causal, dropout_p=dropout_p)


def attention_packed(qkv: torch.Tensor, causal: bool = True) -> torch.Tensor:
  """``qkv``: ``[B, S, 3, H, D]`` (output of a fused QKV projection) -> ``[B, S, H*D]``."""
  from easyparallellibrary_b200.runtime import amp
  if amp.o1_active():
    qkv = amp.cast_args("attention", qkv)
  if _IMPL != "sdpa" and qkv.is_cuda:
    from easyparallellibrary_b200.ops import attention_kernel
    if attention_kernel.supported_packed(qkv):
      return attention_kernel.flash_attention_packed(qkv, causal)
    if _IMPL == "epl":
      raise RuntimeError("hand-written attention kernel does not support this input (needs bf16, head dim 64, contiguous)")
  B, S, _, H, D = qkv.shape
  q, k, v = qkv.permute(2, 0, 3, 1, 4).unbind(0)
  y = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=causal)
  return y.transpose(1, 2).reshape(B, S, H * D)


def attention_reference(q, k, v, causal: bool = True) -> torch.Tensor:
  """Plain fp32 softmax(QK^T)V — the numerics reference for kernel tests."""
  qf, kf, vf = q.float(), k.float(), v.float()
  s = qf @ kf.transpose(-1, -2) / math.sqrt(q.shape[-1])
  if causal:
    S, T = s.shape[-2:]
    mask = torch.ones(S, T, dtype=torch.bool, device=q.device).tril(T - S)
    s = s.masked_fill(~mask, float("-inf"))
  return (torch.softmax(s, -1) @ vf).to(q.dtype)
