"""Capture: how annotations reach the IR without touching PyTorch internals.

Construction time — PyTorch's *global registration hooks*
(``register_module_parameter_registration_hook`` and friends) fire for every
``Parameter`` / buffer / sub-module registered anywhere; while a strategy scope
(or a default strategy) is active the object is tagged with the current
taskgraph.  This replaces the reference's monkey-patched ``Graph._add_op``
(``epl/parallel/hooks.py:97-102, 1029``).

Trace time — :func:`trace_module_costs` runs one forward pass with forward
hooks on leaf modules and records a :class:`Node` per call with static costs
(parameters, FLOPs, activation bytes).  That is the input of the auto-stage
partitioner and of automatic gradient-checkpoint selection (the reference gets
the same quantities from ``tf.profiler`` on a shape-inferred graph,
``epl/profiler/profiler.py:36-60``).
"""
from __future__ import annotations

from typing import Any, Callable, List, Optional, Sequence

import torch
from torch import nn

from easyparallellibrary_b200.ir.graph import Graph
from easyparallellibrary_b200.ir.node import Node, TensorMeta
from easyparallellibrary_b200.ir.phase import ModelPhase

_handles: List[Any] = []


def _on_parameter(module: nn.Module, name: str, param):
  if param is not None:
    g = Graph.get(may_create=False)
    if g is not None:
      g.tag_parameter(param)
  return None


def _on_module(module: nn.Module, name: str, child):
  if child is not None:
    g = Graph.get(may_create=False)
    if g is not None:
      g.tag_module(child)
  return None


def install_hooks() -> None:
  """Idempotent (reference ``hooks.add_hooks`` is idempotent too, hooks.py:1000-1008)."""
  if _handles:
    return
  from torch.nn.modules import module as _m
  _handles.append(_m.register_module_parameter_registration_hook(_on_parameter))
  _handles.append(_m.register_module_module_registration_hook(_on_module))


def remove_hooks() -> None:
  while _handles:
    _handles.pop().remove()


# ------------------------------------------------------------------------------------------
# cost model
# ------------------------------------------------------------------------------------------
_RNG_TYPES = (nn.Dropout, nn.Dropout1d, nn.Dropout2d, nn.Dropout3d, nn.AlphaDropout)


def _numel(shape) -> int:
  n = 1
  for s in shape:
    n *= int(s)
  return n


def module_flops(module: nn.Module, inputs: Sequence[Any], output: Any) -> float:
  """Forward FLOPs of one leaf-module call (multiply-add = 2)."""
  custom = getattr(module, "epl_flops", None)
  if callable(custom):
    return float(custom(inputs, output))
  x = inputs[0] if inputs and isinstance(inputs[0], torch.Tensor) else None
  if isinstance(module, nn.Linear) and x is not None:
    return 2.0 * _numel(x.shape[:-1]) * module.in_features * module.out_features
  if isinstance(module, (nn.Conv1d, nn.Conv2d, nn.Conv3d)) and isinstance(output, torch.Tensor):
    k = _numel(module.kernel_size)
    return 2.0 * _numel(output.shape) * (module.in_channels // module.groups) * k
  if isinstance(module, nn.Embedding):
    return 0.0
  if isinstance(module, (nn.LayerNorm, nn.BatchNorm1d, nn.BatchNorm2d, nn.GroupNorm)) and x is not None:
    return 8.0 * _numel(x.shape)
  if isinstance(module, nn.MultiheadAttention) and x is not None:
    s, b, d = (x.shape + (1, 1))[:3]
    return 8.0 * s * b * d * d + 4.0 * s * s * b * d
  if x is not None:
    own = sum(p.numel() for p in module.parameters(recurse=False))
    if own:
      return 2.0 * own * (_numel(x.shape) / max(int(x.shape[-1]), 1))
    return float(_numel(x.shape))
  return 0.0


def _metas(obj) -> List[TensorMeta]:
  if isinstance(obj, torch.Tensor):
    return [TensorMeta(tuple(obj.shape), obj.dtype)]
  if isinstance(obj, (list, tuple)):
    return [m for o in obj for m in _metas(o)]
  if isinstance(obj, dict):
    return [m for o in obj.values() for m in _metas(o)]
  return []


def trace_module_costs(model: nn.Module, example_inputs: Sequence[Any], example_kwargs: Optional[dict] = None,
                       leaf_predicate: Optional[Callable[[nn.Module], bool]] = None) -> List[Node]:
  """Run ``model(*example_inputs)`` once (no grad) and record one Node per leaf call."""
  graph = Graph.get()
  graph.clear_nodes()
  names = {m: n for n, m in model.named_modules()}
  is_leaf = leaf_predicate or (lambda m: len(list(m.children())) == 0 or getattr(m, "epl_leaf", False))
  handles = []
  order: List[Node] = []

  def hook(mod, args, out):
    name = names.get(mod, type(mod).__name__)
    own = list(mod.parameters(recurse=getattr(mod, "epl_leaf", False)))
    outs = _metas(out)
    node = Node(name=name, type=type(mod).__name__, phase=ModelPhase.FORWARD,
                param_count=sum(p.numel() for p in own),
                param_bytes=sum(p.numel() * p.element_size() for p in own),
                flops=module_flops(mod, args, out),
                act_bytes=sum(m.nbytes for m in outs), depth=name.count(".") + 1 if name else 0,
                inputs=_metas(args), outputs=outs, module=mod,
                has_rng=isinstance(mod, _RNG_TYPES) and getattr(mod, "p", 0) > 0,
                is_collective=bool(getattr(mod, "epl_collective", False)))
    tg = graph.taskgraph_of(mod) or graph.current_taskgraph()
    graph.add_node(node, tg)
    order.append(node)

  for m in model.modules():
    if is_leaf(m):
      handles.append(m.register_forward_hook(hook))
  was_training = model.training
  try:
    with torch.no_grad():
      model(*example_inputs, **(example_kwargs or {}))
  finally:
    for h in handles:
      h.remove()
    model.train(was_training)
  return order
