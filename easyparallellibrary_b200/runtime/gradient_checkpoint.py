"""Activation recomputation.

Reference behaviour (``epl/runtime/gc/gradient_checkpoint.py:80-327``,
``auto_gradient_checkpoint.py:141-199``): checkpoints come from the user
collection ``"checkpoints"`` or are chosen automatically — boundaries of
repeated blocks when the model has them, otherwise a sqrt(n)
memory-balanced partition; ``gradient_checkpoint.end_taskgraph`` limits how
far the automatic choice reaches; RNG ops, loops and all-to-all are never
recomputed; ``check_gradients`` validates recompute gradients.

Eager translation: a *segment* is a module (or a run of consecutive modules)
whose forward is wrapped with ``torch.utils.checkpoint`` (non-reentrant).  RNG
state is preserved across recompute so dropout masks are identical (the
reference keeps masks by refusing to recompute them); modules flagged
``epl_collective`` (MoE dispatch/combine) are left outside segments.
"""
from __future__ import annotations

import math
from typing import Any, List, Optional, Sequence

import torch
from torch import nn
from torch.utils.checkpoint import checkpoint

from easyparallellibrary_b200.parallel import partitioner
from easyparallellibrary_b200.utils import constant


RECOMPUTE_ENABLED = True        # gradient_checkpoint.check_gradients flips this to obtain the plain (non-recomputed) gradients


class _Checkpointed(nn.Module):
  """Wraps ``inner`` so its activations are recomputed in backward."""

  def __init__(self, inner: nn.Module):
    super().__init__()
    self.inner = inner

  def forward(self, *args, **kwargs):
    if not torch.is_grad_enabled() or not self.training or not RECOMPUTE_ENABLED:
      return self.inner(*args, **kwargs)
    return checkpoint(self.inner, *args, use_reentrant=False, preserve_rng_state=True, **kwargs)


def wrap_module(parent: nn.Module, child_name: str) -> None:
  child = getattr(parent, child_name)
  if isinstance(child, _Checkpointed):
    return
  parent._modules[child_name] = _Checkpointed(child)


def _find_parent(model: nn.Module, target: nn.Module):
  for parent in model.modules():
    for name, child in parent._modules.items():
      if child is target:
        return parent, name
  return None, None


def select_auto(nodes: Sequence[Any], end_taskgraph: int = -1) -> List[nn.Module]:
  """Choose modules to checkpoint from traced IR nodes."""
  nodes = [n for n in nodes if end_taskgraph < 0 or n.taskgraph <= end_taskgraph]
  nodes = [n for n in nodes if not n.is_collective]
  blocks = partitioner.find_repeated_blocks(nodes, min_dup=constant.MIN_REPEAT_BLOCKS)
  chosen: List[nn.Module] = []
  if blocks:
    # the module that spans exactly one block = common scope prefix of its nodes
    for blk in blocks:
      scope = blk[0].name.split(".")
      for n in blk[1:]:
        parts = n.name.split(".")
        k = 0
        while k < min(len(scope), len(parts)) and scope[k] == parts[k]:
          k += 1
        scope = scope[:k]
      chosen.append(".".join(scope))
    return chosen
  # sqrt(n) memory-balanced partition over activation bytes
  if not nodes:
    return []
  k = max(int(math.sqrt(len(nodes))), 1)
  groups = partitioner.partition_stages(nodes, [max(n.act_bytes, 1) for n in nodes], k)
  return [g[0].name for g in groups if g]


def apply_gradient_checkpoint(model: nn.Module, gc_type: str, nodes: Optional[Sequence[Any]] = None,
                              collection: Optional[Sequence[Any]] = None, end_taskgraph: int = -1) -> List[str]:
  """Wrap the selected modules in place; returns their qualified names."""
  gc_type = (gc_type or "").lower()
  if not gc_type:
    return []
  names = {m: n for n, m in model.named_modules()}
  by_name = dict(model.named_modules())
  targets: List[nn.Module] = []
  if gc_type == constant.GC_COLLECTION:
    for obj in collection or []:
      if isinstance(obj, nn.Module):
        targets.append(obj)
      elif isinstance(obj, str) and obj in by_name:
        targets.append(by_name[obj])
    if not targets:
      raise RuntimeError('gradient_checkpoint.type="collection" but the "checkpoints" collection holds no module; '
                         'use epl.add_to_collection(module, epl.GraphKeys.GC_CHECKPOINTS)')
  elif gc_type == constant.GC_AUTO:
    if nodes is None:
      # no trace available: fall back on structure — checkpoint every element of the longest ModuleList/Sequential
      best = None
      for m in model.modules():
        if isinstance(m, (nn.ModuleList, nn.Sequential)) and len(m) >= constant.MIN_REPEAT_BLOCKS:
          if best is None or len(m) > len(best):
            best = m
      targets = list(best) if best is not None else []
    else:
      for name in select_auto(nodes, end_taskgraph):
        if name in by_name:
          targets.append(by_name[name])
  else:
    raise ValueError("unknown gradient_checkpoint.type %r" % gc_type)
  done = []
  for t in targets:
    if any(getattr(sub, "epl_collective", False) for sub in t.modules()):
      continue
    parent, cname = _find_parent(model, t)
    if parent is not None:
      wrap_module(parent, cname)
      done.append(names.get(t, cname))
  return done


def check_gradients(model_fn, params: Sequence[torch.Tensor], atol: float = 1e-5) -> bool:
  """Run ``model_fn(use_checkpoint)`` twice and compare parameter gradients."""
  grads = []
  for flag in (False, True):
    for p in params:
      p.grad = None
    model_fn(flag).backward()
    grads.append([p.grad.detach().clone() for p in params])
  return all(torch.allclose(a, b, atol=atol) for a, b in zip(*grads))
