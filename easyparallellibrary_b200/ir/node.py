"""IR nodes.

The reference wraps every ``tf.Operation``/``tf.Tensor`` (``epl/ir/operation.py``,
``tensor.py``).  Eager PyTorch has no op graph to wrap, so the IR is coarser and
cheaper: one :class:`Node` per *leaf module call* (or per free parameter), with
the static costs the planners need.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, List, Optional, Tuple

from easyparallellibrary_b200.ir.phase import ModelPhase


@dataclass
class TensorMeta:
  shape: Tuple[int, ...]
  dtype: Any
  producer: Optional["Node"] = None

  @property
  def numel(self) -> int:
    n = 1
    for s in self.shape:
      n *= int(s)
    return n

  @property
  def nbytes(self) -> int:
    size = getattr(self.dtype, "itemsize", None) or 4
    return self.numel * size


@dataclass
class Node:
  name: str                       # qualified module path, e.g. "h.3.mlp.fc"
  type: str                       # module class name (the repeated-block detector keys on this)
  taskgraph: int = -1
  phase: ModelPhase = ModelPhase.FORWARD
  param_count: int = 0
  param_bytes: int = 0
  flops: float = 0.0              # forward FLOPs for the captured example input
  act_bytes: int = 0              # bytes of outputs kept for backward
  depth: int = 0                  # nesting depth of the module path
  inputs: List[TensorMeta] = field(default_factory=list)
  outputs: List[TensorMeta] = field(default_factory=list)
  module: Any = None
  has_rng: bool = False           # dropout & friends: never recomputed (reference gradient_checkpoint.py:224-225)
  is_collective: bool = False     # all-to-all etc.: never recomputed (reference constant.py:97)

  @property
  def scope(self) -> str:
    return self.name.rsplit(".", 1)[0] if "." in self.name else ""

  def __repr__(self):
    return "Node(%s:%s tg=%d params=%d flops=%.3g)" % (self.name, self.type, self.taskgraph, self.param_count, self.flops)
