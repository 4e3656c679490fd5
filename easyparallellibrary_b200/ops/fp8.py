"""fp8 (e4m3) forward GEMMs on the 2-CTA tcgen05 kernel (``tcgen05.mma.kind::f8f6f4``, SASS ``UTCQMMA``).

``amp.level = "fp8"``: weights and activations stay bf16 (fp32 master weights in the flat optimizer); the *forward* GEMMs of
``ops.linear`` quantise both operands per tensor to e4m3 (scale = 448 / amax, computed on the device, no host round trip)
and multiply on the fp8 tensor-core path — twice the math per shared-memory byte of the bf16 kernel — with the
de-quantisation factors applied to the fp32 accumulator in the epilogue, in front of the usual bias / GELU / residual
epilogues.  Backward GEMMs (dX, dW) stay bf16: gradients need the range, and dW accumulates into the gradient buckets.
The reference has no 8-bit path (its AMP is fp16 O1, ``epl/runtime/amp``); this is the BASELINE.json "fp8 path".
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from easyparallellibrary_b200.ops import _lib

ENABLED = False          # set by the engine for amp.level = "fp8" (tests flip it directly)


def quantize_e4m3(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
  """``x`` (bf16 / fp16 / fp32, contiguous) -> ``(q, inv_scale)``: ``q`` uint8 storage of e4m3 values, ``x ~ q * inv_scale``."""
  lib = _lib.require()
  x = x.contiguous()
  q = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
  scratch = torch.empty(2, dtype=torch.float32, device=x.device)
  rc = lib.epl_quantize_e4m3(x.data_ptr(), _lib.dtype_code(x.dtype), x.numel(), q.data_ptr(), scratch[0:1].data_ptr(),
                             scratch[1:2].data_ptr(), _lib.stream())
  _lib.check(rc, "quantize_e4m3")
  _lib.launches += 1
  return q, scratch[1:2]


def supported(M: int, N: int, K: int) -> bool:
  return M >= 256 and N >= 256 and K % 16 == 0


def gemm_fp8(a: torch.Tensor, b: torch.Tensor, bias: Optional[torch.Tensor] = None, epilogue: int = 0,
             pre: Optional[torch.Tensor] = None, aux: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
             out_dtype: Optional[torch.dtype] = None, b_q: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> torch.Tensor:
  """``D[M,N] = a[M,K] @ b[N,K]^T`` with both operands quantised to e4m3 (``b_q``: a cached quantisation of ``b``)."""
  lib = _lib.require()
  from easyparallellibrary_b200.ops import linear as L
  M, K = a.shape
  N = b.shape[0]
  aq, sa = quantize_e4m3(a)
  bq, sb = b_q if b_q is not None else quantize_e4m3(b)
  if out is None:
    out = torch.empty((M, N), dtype=out_dtype or a.dtype, device=a.device)
  rc = lib.epl_gemm_fp8(aq.data_ptr(), bq.data_ptr(), out.data_ptr(), M, N, K, K, K, out.stride(0), _lib.ptr(bias), _lib.ptr(pre),
                        _lib.ptr(aux), epilogue, _lib.dtype_code(out.dtype), 1.0, sa.data_ptr(), sb.data_ptr(), L._NUM_SMS, _lib.stream())
  _lib.check(rc, "gemm_fp8")
  return out
