"""Communication pool: ``num_communicators`` independent transports per
logical communicator so several buckets are in flight at once.

Parity: ``epl/communicators/communication_pool.py:84-105`` — buckets are issued
**last to first** (the gradients of the last layers are ready first),
round-robin over the pool slots, serialised per slot.
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import torch


class CommunicationPool(object):
  def __init__(self, backends: Sequence[object]):
    if not backends:
      raise ValueError("a pool needs at least one communicator")
    self.backends = list(backends)

  @property
  def size(self) -> int:
    return len(self.backends)

  def issue_order(self, num_buckets: int) -> List[int]:
    return list(range(num_buckets - 1, -1, -1))

  def slot_of(self, issue_position: int) -> int:
    return issue_position % len(self.backends)

  def communicate(self, buckets: Sequence[torch.Tensor], fn: Callable[[object, torch.Tensor], torch.Tensor]) -> List[torch.Tensor]:
    """Apply ``fn(backend, bucket)`` to every bucket; results keep the input order."""
    out: List[torch.Tensor] = [None] * len(buckets)
    for pos, b in enumerate(self.issue_order(len(buckets))):
      out[b] = fn(self.backends[self.slot_of(pos)], buckets[b])
    return out

  def close(self) -> None:
    for b in self.backends:
      b.close()
