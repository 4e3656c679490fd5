"""Automatic pipeline-stage search over the module-level IR.

Parity: ``epl/parallel/planner.py`` — ``AutoStageGenerator.search`` (49-60)
with the policies BALANCE_OP_NUM / REPEATED_LAYERS / HEURISTIC (66-112).  The
weights come from the traced cost model (FLOPs, parameters, activation bytes)
rather than op counts, because on B200 a stage's time is its GEMM FLOPs and
its HBM traffic, not the number of framework ops.
"""
from __future__ import annotations

from typing import Any, List, Optional, Sequence

from easyparallellibrary_b200.parallel import partitioner
from easyparallellibrary_b200.utils import constant


def node_weight(node, policy: str) -> float:
  if policy == constant.STAGE_POLICY_BALANCE_OP_NUM:
    return 1.0
  flops = float(getattr(node, "flops", 0.0))
  params = float(getattr(node, "param_count", 0))
  # forward+backward FLOPs dominate; parameters add optimizer/HBM traffic (16 B/param against ~6.5 TB/s
  # vs ~1.4 PFLOP/s: 1 param ~ 3.5e3 FLOP-equivalents).
  return 3.0 * flops + 3.5e3 * params + 1.0


class AutoStageGenerator(object):
  def __init__(self, policy: str = constant.STAGE_POLICY_HEURISTIC, num_stages: int = 2):
    if num_stages < 1:
      raise ValueError("num_stages must be >= 1")
    self.policy = policy
    self.num_stages = num_stages

  def search(self, nodes: Optional[Sequence[Any]] = None) -> List[List[Any]]:
    """Return ``num_stages`` lists of nodes (execution order preserved)."""
    if nodes is None:
      from easyparallellibrary_b200.ir.graph import Graph
      nodes = Graph.get().operations
    nodes = list(nodes)
    if not nodes:
      return [[] for _ in range(self.num_stages)]
    if self.policy == constant.STAGE_POLICY_BALANCE_OP_NUM:
      return partitioner.partition_stages(nodes, [1.0] * len(nodes), self.num_stages)
    if self.policy == constant.STAGE_POLICY_REPEATED_LAYERS:
      stages = self._by_repeated_blocks(nodes)
      if stages is None:
        raise RuntimeError("No repeated blocks found; use the heuristic policy instead.")
      return stages
    stages = self._by_repeated_blocks(nodes)
    if stages is not None:
      return stages
    return partitioner.partition_stages(nodes, [node_weight(n, self.policy) for n in nodes], self.num_stages)

  def _by_repeated_blocks(self, nodes: List[Any]) -> Optional[List[List[Any]]]:
    blocks = partitioner.find_repeated_blocks(nodes, min_dup=max(constant.MIN_REPEAT_BLOCKS, self.num_stages))
    if len(blocks) < self.num_stages:
      return None
    index = {id(n): i for i, n in enumerate(nodes)}
    # units: [prefix + block0], block1, ..., [block_last + suffix]; cut only at block boundaries
    starts = [index[id(b[0])] for b in blocks]
    units: List[List[Any]] = []
    for k, s in enumerate(starts):
      a = 0 if k == 0 else s
      b = starts[k + 1] if k + 1 < len(starts) else len(nodes)
      units.append(nodes[a:b])
    weights = [sum(node_weight(n, self.policy) for n in u) for u in units]
    groups = partitioner.partition_stages(units, weights, self.num_stages)
    return [[n for u in g for n in u] for g in groups]


def stage_module_names(stages: Sequence[Sequence[Any]], depth: int = 2, sep: str = ".") -> List[List[str]]:
  """Collapse node lists to unique module-path prefixes (what the engine cuts at)."""
  out = []
  for st in stages:
    seen: List[str] = []
    for n in st:
      key = sep.join(n.name.split(sep)[:depth])
      if key not in seen:
        seen.append(key)
    out.append(seen)
  return out
