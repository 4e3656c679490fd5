"""Taskgraph = everything annotated by one strategy scope = one pipeline stage
or one tensor-parallel region (reference ``epl/ir/taskgraph.py:107-577``).

The reference's taskgraph exists to find stage entrance/exit ops for control
edges; here a taskgraph owns *modules and parameters*, and the stage boundary
is simply "the tensors passed between consecutive stage modules".
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

from easyparallellibrary_b200.ir.node import Node
from easyparallellibrary_b200.ir.phase import ModelPhase
from easyparallellibrary_b200.strategies.base import ParallelStrategy, Replicate, Split


class Taskgraph(object):
  def __init__(self, index: int, strategy: Optional[ParallelStrategy]):
    self.index = index
    self.strategy = strategy
    self.virtual_device = None
    self.nodes: Dict[ModelPhase, List[Node]] = {p: [] for p in ModelPhase}
    self.parameters: List[Any] = []     # torch Parameters (insertion ordered, unique)
    self._param_ids = set()
    self.modules: List[Any] = []
    self._module_ids = set()

  # -- membership ---------------------------------------------------------------
  def add_parameter(self, param) -> None:
    if id(param) not in self._param_ids:
      self._param_ids.add(id(param))
      self.parameters.append(param)

  def add_module(self, module) -> None:
    if id(module) not in self._module_ids:
      self._module_ids.add(id(module))
      self.modules.append(module)

  def add_node(self, node: Node) -> None:
    node.taskgraph = self.index
    self.nodes[node.phase].append(node)

  def owns(self, param) -> bool:
    return id(param) in self._param_ids

  # -- properties ----------------------------------------------------------------
  @property
  def is_split(self) -> bool:
    return isinstance(self.strategy, Split)

  @property
  def is_replicate(self) -> bool:
    return self.strategy is None or isinstance(self.strategy, Replicate)

  @property
  def num_device_per_replica(self) -> int:
    if self.strategy is None or self.strategy.device_count is None:
      return 1
    return self.strategy.device_count

  @property
  def num_replicas(self) -> int:
    return self.virtual_device.num_replicas if self.virtual_device is not None else 1

  @property
  def local_num_replicas(self) -> int:
    if self.virtual_device is None:
      return 1
    local = {d.rank for d in self.virtual_device.owned_devices}
    return sum(1 for s in self.virtual_device.slice_devices if any(d.rank in local for d in s))

  @property
  def pipeline_config(self):
    from easyparallellibrary_b200.env import Env
    return Env.get().config.pipeline

  @property
  def operations(self) -> List[Node]:
    return [n for nodes in self.nodes.values() for n in nodes]

  def get_variables(self) -> List[Any]:
    return list(self.parameters)

  @property
  def param_count(self) -> int:
    return sum(int(p.numel()) for p in self.parameters)

  @property
  def flops(self) -> float:
    return sum(n.flops for n in self.nodes[ModelPhase.FORWARD])

  def format(self, max_depth: int = 2) -> str:
    """Pretty tree: strategy -> module scopes with devices (reference taskgraph.py:485-529)."""
    name = self.strategy.name if self.strategy is not None and self.strategy.name else ""
    kind = type(self.strategy).__name__ if self.strategy is not None else "Default"
    head = "Taskgraph %d [%s%s devices/replica=%d params=%d]" % (
        self.index, kind, (":" + name) if name else "", self.num_device_per_replica, self.param_count)
    if self.virtual_device is not None:
      head += " on %s" % self.virtual_device.ranks()
    lines = [head]
    seen = []
    for n in self.nodes[ModelPhase.FORWARD]:
      scope = ".".join(n.name.split(".")[:max_depth])
      if scope not in seen:
        seen.append(scope)
    for s in seen:
      lines.append("  - " + s)
    return "\n".join(lines)

  def __repr__(self):
    return "Taskgraph(index=%d, strategy=%r, params=%d)" % (self.index, self.strategy, len(self.parameters))
