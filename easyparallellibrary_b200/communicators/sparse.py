"""Sparse gradient reduction: all-gather values and indices instead of a dense
all-reduce (reference ``rewriters/sparse_allreduce.py:127-160``)."""
from __future__ import annotations

import torch


def sparse_all_reduce(comm, grad: torch.Tensor, mean: bool = False) -> torch.Tensor:
  """``grad`` is a ``torch.sparse_coo`` tensor (e.g. from ``nn.Embedding(sparse=True)``)."""
  grad = grad.coalesce()
  idx = grad.indices().t().contiguous()           # [nnz, ndim_sparse]
  val = grad.values().contiguous()                # [nnz, ...]
  all_val, _ = comm.allgatherv(val)
  all_idx, _ = comm.allgatherv(idx)
  if mean:
    all_val = all_val / comm.size
  return torch.sparse_coo_tensor(all_idx.t(), all_val, grad.shape).coalesce()
