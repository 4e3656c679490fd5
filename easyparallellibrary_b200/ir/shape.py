"""Sharded shape types (reference ``epl/ir/shape.py:26-207`` — there "types only, unused by any transform").
Here they describe how a tensor-parallel parameter relates to its unsharded form; ``runtime/saver.ShardingLoader``
and ``ops/tensor_parallel.add_weight`` use them to slice checkpoints."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple


@dataclass(frozen=True)
class Dimension:
  size: int                      # global extent
  num_shards: int = 1
  shard_index: int = 0
  remainder_to_first: bool = False

  @property
  def bounds(self) -> Tuple[int, int]:
    q, r = divmod(self.size, self.num_shards)
    if self.remainder_to_first:
      sizes = [q + r] + [q] * (self.num_shards - 1)
    else:
      sizes = [q + (1 if i < r else 0) for i in range(self.num_shards)]
    lo = sum(sizes[:self.shard_index])
    return lo, lo + sizes[self.shard_index]

  @property
  def local_size(self) -> int:
    lo, hi = self.bounds
    return hi - lo

  @property
  def is_sharded(self) -> bool:
    return self.num_shards > 1


class Shape(object):
  def __init__(self, dims: Sequence[Dimension]):
    self.dims: List[Dimension] = list(dims)

  @staticmethod
  def replicated(sizes: Sequence[int]) -> "Shape":
    return Shape([Dimension(int(s)) for s in sizes])

  def shard(self, dim: int, num_shards: int, shard_index: int, remainder_to_first: bool = False) -> "Shape":
    dims = list(self.dims)
    dims[dim] = Dimension(dims[dim].size, num_shards, shard_index, remainder_to_first)
    return Shape(dims)

  @property
  def global_shape(self) -> Tuple[int, ...]:
    return tuple(d.size for d in self.dims)

  @property
  def local_shape(self) -> Tuple[int, ...]:
    return tuple(d.local_size for d in self.dims)

  def slices(self) -> Tuple[Tuple[int, int], ...]:
    """(begin, size) per dimension — the ``sharding_info`` format of ``ShardingLoader``."""
    return tuple((d.bounds[0], d.local_size) for d in self.dims)

  def __repr__(self) -> str:
    return "Shape(%s)" % ", ".join("%d%s" % (d.size, "/%d@%d" % (d.num_shards, d.shard_index) if d.is_sharded else "") for d in self.dims)
