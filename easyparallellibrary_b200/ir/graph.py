"""The IR graph: taskgraphs, captured nodes and user collections.

Parity with ``epl/ir/graph.py``: ``GraphKeys`` (40-65), ``Graph.get()``
(162-171), taskgraph creation on scope entry (319-331, 532-534), collections
(600-649), ``pipeline_enabled`` (918-923: at least two taskgraphs *and*
``num_micro_batch > 1``), ``need_parallel`` (926-933: only training is
parallelised), ``set_default_strategy`` (942-949: replicate only), ``format()``
(587-598).

What is different: membership is decided when a ``torch.nn.Parameter`` /
sub-module is *registered* (PyTorch's global registration hooks) or when a
forward-time scope is active — not by classifying every op of a static graph.
"""
from __future__ import annotations

import weakref
from collections import OrderedDict
from typing import Any, Dict, List, Optional

from easyparallellibrary_b200.ir.node import Node
from easyparallellibrary_b200.ir.phase import ModelPhase
from easyparallellibrary_b200.ir.taskgraph import Taskgraph
from easyparallellibrary_b200.strategies.base import ParallelStrategy, Replicate, Split


class _IdMap(object):
  """Identity-keyed weak map (tensors cannot live in a WeakKeyDictionary: ``==`` is element-wise)."""

  def __init__(self):
    self._d: Dict[int, Any] = {}

  def _drop(self, key):
    self._d.pop(key, None)

  def __contains__(self, obj) -> bool:
    return id(obj) in self._d

  def __setitem__(self, obj, value) -> None:
    key = id(obj)
    try:
      ref = weakref.ref(obj, lambda _r, k=key: self._drop(k))
    except TypeError:
      ref = None
    self._d[key] = (ref, value)

  def __getitem__(self, obj):
    return self._d[id(obj)][1]

  def get(self, obj, default=None):
    hit = self._d.get(id(obj))
    return default if hit is None else hit[1]

  def clear(self) -> None:
    self._d.clear()


class GraphKeys(object):
  """Collections whose members are merged over micro-batches (LOCAL_*) and
  additionally over replicas (GLOBAL_*)."""
  GLOBAL_CONCAT_OBJECTS = "global_concat_objects"
  GLOBAL_MEAN_OBJECTS = "global_mean_objects"
  GLOBAL_SUM_OBJECTS = "global_sum_objects"
  LOCAL_CONCAT_OBJECTS = "local_concat_objects"
  LOCAL_MEAN_OBJECTS = "local_mean_objects"
  LOCAL_SUM_OBJECTS = "local_sum_objects"
  ALL_COLLECTION_KEYS = (GLOBAL_CONCAT_OBJECTS, GLOBAL_MEAN_OBJECTS, GLOBAL_SUM_OBJECTS,
                         LOCAL_CONCAT_OBJECTS, LOCAL_MEAN_OBJECTS, LOCAL_SUM_OBJECTS)
  GC_CHECKPOINTS = "checkpoints"


class Graph(object):
  """One per ``Env``; use ``Graph.get()``."""

  def __init__(self):
    self._taskgraphs: List[Taskgraph] = []
    self._by_strategy: Dict[int, Taskgraph] = {}
    self._param_tg = _IdMap()
    self._module_tg = _IdMap()
    self._collections: "OrderedDict[str, List[Any]]" = OrderedDict()
    self._nodes: List[Node] = []
    self.training = True
    self.parallel_information: Dict[str, Any] = {}
    self.current_micro_batch = None      # the (inputs, *targets) tuple of the micro-batch being executed, on every stage

  # ------------------------------------------------------------------ access
  @staticmethod
  def get(may_create: bool = True) -> "Graph":
    from easyparallellibrary_b200.env import Env
    env = Env.get()
    if env.graph is None and may_create:
      env.graph = Graph()
    return env.graph

  def reset(self) -> None:
    self.__init__()

  @property
  def taskgraphs(self) -> List[Taskgraph]:
    return self._taskgraphs

  @property
  def operations(self) -> List[Node]:
    return list(self._nodes)

  @property
  def num_stages(self) -> int:
    return len([t for t in self._taskgraphs if t.is_replicate]) or (1 if not self._taskgraphs else 0)

  @property
  def pipeline_enabled(self) -> bool:
    from easyparallellibrary_b200.env import Env
    stages = [t for t in self._taskgraphs if t.is_replicate]
    return len(stages) > 1 and Env.get().config.pipeline.num_micro_batch > 1

  @property
  def need_parallel(self) -> bool:
    return self.training

  @property
  def num_constructors(self) -> int:
    from easyparallellibrary_b200.env import Env
    c = Env.get().cluster
    return c.total_gpu_num if c is not None else 1

  # ------------------------------------------------------------------ taskgraph bookkeeping
  def _context(self):
    from easyparallellibrary_b200.env import Env
    return Env.get().strategy_context

  def _taskgraph_for(self, strategy: Optional[ParallelStrategy], create: bool = True) -> Optional[Taskgraph]:
    if strategy is None:
      if self._taskgraphs:
        # outside any scope: fall into the last replicate taskgraph (reference graph.py:339-345)
        for tg in reversed(self._taskgraphs):
          if tg.is_replicate:
            return tg
        return self._taskgraphs[-1]
      if not create:
        return None
      tg = Taskgraph(0, None)
      self._taskgraphs.append(tg)
      return tg
    tg = self._by_strategy.get(id(strategy))
    if tg is None and create:
      tg = Taskgraph(len(self._taskgraphs), strategy)
      self._taskgraphs.append(tg)
      self._by_strategy[id(strategy)] = tg
    return tg

  def current_taskgraph(self, create: bool = True) -> Optional[Taskgraph]:
    ctx = self._context()
    if ctx is None:
      return self._taskgraph_for(None, create)
    ctx.update_flag = False
    return self._taskgraph_for(ctx.current, create)

  # ------------------------------------------------------------------ tagging (construction time)
  def tag_parameter(self, param) -> Optional[Taskgraph]:
    if param is None:
      return None
    if param in self._param_tg:        # shared / tied weights stay where they were first created
      return self._taskgraphs[self._param_tg[param]]
    tg = self.current_taskgraph()
    self._param_tg[param] = tg.index
    tg.add_parameter(param)
    return tg

  def tag_module(self, module) -> Optional[Taskgraph]:
    ctx = self._context()
    if ctx is None or ctx.current is None:
      return None
    if module in self._module_tg:
      return self._taskgraphs[self._module_tg[module]]
    # a module is registered with its parent AFTER it was built (``self.h = nn.ModuleList(blocks)`` runs when the last
    # scope is already open): it belongs where its own parameters were created, not where it was attached
    tg = None
    if hasattr(module, "parameters"):
      for p in module.parameters():
        idx = self._param_tg.get(p)
        if idx is not None:
          tg = self._taskgraphs[idx]
          break
    if tg is None:
      tg = self.current_taskgraph()
    self._module_tg[module] = tg.index
    tg.add_module(module)
    return tg

  def taskgraph_of(self, obj) -> Optional[Taskgraph]:
    """Taskgraph of a Parameter or Module (modules fall back to their first tagged parameter)."""
    idx = self._param_tg.get(obj) if _is_tensor(obj) else self._module_tg.get(obj)
    if idx is None and not _is_tensor(obj) and hasattr(obj, "parameters"):
      for p in obj.parameters():
        idx = self._param_tg.get(p)
        if idx is not None:
          break
    return self._taskgraphs[idx] if idx is not None else None

  def assign_parameter(self, param, taskgraph_index: int) -> None:
    """Used by the auto partitioner to move ownership."""
    old = self._param_tg.get(param)
    if old is not None and old != taskgraph_index:
      tg = self._taskgraphs[old]
      tg.parameters = [p for p in tg.parameters if p is not param]
      tg._param_ids.discard(id(param))
    self._param_tg[param] = taskgraph_index
    self._taskgraphs[taskgraph_index].add_parameter(param)

  def new_taskgraph(self, strategy: Optional[ParallelStrategy]) -> Taskgraph:
    tg = Taskgraph(len(self._taskgraphs), strategy)
    self._taskgraphs.append(tg)
    if strategy is not None:
      self._by_strategy[id(strategy)] = tg
    return tg

  def add_node(self, node: Node, taskgraph: Optional[Taskgraph] = None) -> Node:
    tg = taskgraph or self.current_taskgraph()
    tg.add_node(node)
    self._nodes.append(node)
    return node

  def clear_nodes(self) -> None:
    self._nodes = []
    for tg in self._taskgraphs:
      for p in tg.nodes:
        tg.nodes[p] = []

  # ------------------------------------------------------------------ default strategy
  def set_default_strategy(self, strategy: Optional[ParallelStrategy]) -> None:
    if strategy is not None and not isinstance(strategy, Replicate):
      raise ValueError("Only replicate can be the default strategy, got %r" % (strategy,))
    self._context().default_strategy = strategy

  # ------------------------------------------------------------------ collections
  def add_to_collection(self, objs, key: str) -> None:
    if key not in GraphKeys.ALL_COLLECTION_KEYS and key != GraphKeys.GC_CHECKPOINTS:
      raise ValueError("Unknown collection %r; expected one of %s" % (key, GraphKeys.ALL_COLLECTION_KEYS))
    if not isinstance(objs, (list, tuple)):
      objs = [objs]
    self._collections.setdefault(key, []).extend(objs)

  def get_collection(self, key: str) -> List[Any]:
    return list(self._collections.get(key, []))

  def get_all_collections(self) -> List[Any]:
    return [o for k in GraphKeys.ALL_COLLECTION_KEYS for o in self._collections.get(k, [])]

  def pop_collections(self) -> "OrderedDict[str, List[Any]]":
    """Engine hook: take everything registered during one micro-batch forward."""
    out = OrderedDict((k, v) for k, v in self._collections.items() if k != GraphKeys.GC_CHECKPOINTS and v)
    for k in list(out):
      self._collections[k] = []
    return out

  # ------------------------------------------------------------------ printing
  def format(self, max_depth: int = 2) -> str:
    lines = ["Graph: %d taskgraph(s), pipeline=%s" % (len(self._taskgraphs), self.pipeline_enabled)]
    for tg in self._taskgraphs:
      lines.append(tg.format(max_depth))
    return "\n".join(lines)

  def __repr__(self) -> str:
    return self.format()


def _is_tensor(obj) -> bool:
  import torch
  return isinstance(obj, torch.Tensor)


# module level helpers re-exported from the package root (reference graph.py:952-961)
def add_to_collection(objs, key: str) -> None:
  Graph.get().add_to_collection(objs, key)


def get_collection(key: str):
  return Graph.get().get_collection(key)


def get_all_collections():
  return Graph.get().get_all_collections()


def current_micro_batch():
  """The ``(inputs, *targets)`` tuple of the micro-batch whose forward is running, or ``None`` outside a training step.

  Under pipeline parallelism only one tensor travels between stages, but every stage is handed the whole micro-batch, so a
  later stage reads side inputs (sequence length, masks, position ids) from here instead of re-deriving them from the
  activation (the reference gets this for free: every stage subgraph sees the same input tensors of its micro-batch clone,
  ``graph_editor.py:397-421``)."""
  g = Graph.get(may_create=False)
  return None if g is None else g.current_micro_batch
