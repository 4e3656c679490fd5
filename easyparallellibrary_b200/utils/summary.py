"""Named summaries of values computed inside the model, merged like the reference merges ``tf.summary`` tensors.

In the reference a ``tf.summary.scalar / histogram`` placed in the model is rewired to the *merged* tensor of its collection —
mean / sum / concatenation over the micro-batches and the replicas — and recorded in ``Graph.summary_map`` (tags, tensor name,
summary type; ``tests/summary_test.py:57-110``), and the first constructor's session writes the events.  Here::

    def forward(self, x, labels):
      logits = self.net(x)
      epl.summary.scalar("accuracy", (logits.argmax(-1) == labels).float().mean())      # GLOBAL_MEAN over micro-batches + replicas
      epl.summary.histogram("features", x)                                               # GLOBAL_CONCAT
      ...
    trainer.hooks.append(epl.summary.SummaryHook("/tmp/run1", every=10))

``scalar`` / ``histogram`` put the tensor into the matching collection (``ir/graph.py::GraphKeys``) and remember its tag and slot in
``Graph.summary_map``; after the step the merged values come back in ``StepOutput.collections`` and ``SummaryHook`` (rank 0 only, like
the reference's first constructor) writes them — together with loss, learning rate, loss scale and gradient norm — as TensorBoard
events when ``tensorboard`` is importable and always as one JSON line per step in ``<log_dir>/summaries.jsonl``.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from typing import Any, Dict, List, Optional

import torch

from easyparallellibrary_b200.ir.graph import Graph, GraphKeys

SUMMARY_SCALAR_TYPE = "SUMMARY_SCALAR_TYPE"
SUMMARY_HISTOGRAM_TYPE = "SUMMARY_HISTOGRAM_TYPE"

_REDUCE_KEYS = {
    ("mean", True): GraphKeys.GLOBAL_MEAN_OBJECTS, ("sum", True): GraphKeys.GLOBAL_SUM_OBJECTS,
    ("concat", True): GraphKeys.GLOBAL_CONCAT_OBJECTS, ("mean", False): GraphKeys.LOCAL_MEAN_OBJECTS,
    ("sum", False): GraphKeys.LOCAL_SUM_OBJECTS, ("concat", False): GraphKeys.LOCAL_CONCAT_OBJECTS,
}


@dataclass
class SummaryInfo:
  tags: str
  summary_type: str
  collection: str
  index: int                     # slot of the tensor inside its collection within one micro-batch forward


def _register(name: str, tensor: torch.Tensor, kind: str, reduce: str, across_replicas: bool) -> None:
  g = Graph.get()
  key = _REDUCE_KEYS[(reduce, across_replicas)]
  if not hasattr(g, "summary_map"):
    g.summary_map = {}
  index = len(g.get_collection(key))
  g.add_to_collection(tensor.detach(), key)
  g.summary_map[name] = SummaryInfo(tags=name, summary_type=kind, collection=key, index=index)


def scalar(name: str, tensor: torch.Tensor, reduce: str = "mean", across_replicas: bool = True) -> None:
  """Record a scalar; merged by ``reduce`` ("mean" | "sum") over micro-batches and (``across_replicas``) replicas."""
  if reduce not in ("mean", "sum"):
    raise ValueError("scalar summaries merge by 'mean' or 'sum'")
  _register(name, tensor, SUMMARY_SCALAR_TYPE, reduce, across_replicas)


def histogram(name: str, tensor: torch.Tensor, across_replicas: bool = True) -> None:
  """Record the distribution of ``tensor``: concatenated over micro-batches and replicas along dim 0."""
  _register(name, tensor, SUMMARY_HISTOGRAM_TYPE, "concat", across_replicas)


def merged_summaries(out) -> Dict[str, torch.Tensor]:
  """``{tag: merged value}`` of one ``StepOutput`` (empty on ranks that do not run the forward that produced them)."""
  g = Graph.get()
  res: Dict[str, torch.Tensor] = {}
  for name, info in getattr(g, "summary_map", {}).items():
    vals = out.collections.get(info.collection) if out is not None and out.collections else None
    if vals is not None and info.index < len(vals):
      res[name] = vals[info.index]
  return res


class SummaryHook(object):
  """Trainer hook: writes the step's summaries every ``every`` steps (rank 0 only)."""

  def __init__(self, log_dir: str, every: int = 1, tensorboard: Optional[bool] = None):
    self.log_dir, self.every = log_dir, max(int(every), 1)
    self.rank = int(os.environ.get("RANK", "0"))
    self._tb = None
    self._want_tb = tensorboard
    self._file = None

  def _open(self) -> None:
    os.makedirs(self.log_dir, exist_ok=True)
    self._file = open(os.path.join(self.log_dir, "summaries.jsonl"), "a")
    if self._want_tb is not False:
      try:
        from torch.utils.tensorboard import SummaryWriter
        self._tb = SummaryWriter(self.log_dir)
      except Exception:
        if self._want_tb:
          raise
        self._tb = None

  def before_step(self, trainer) -> None:
    pass

  def after_step(self, trainer, out) -> None:
    step = trainer.global_step
    if self.rank != 0 or out is None or step % self.every:
      return
    if self._file is None:
      self._open()
    row: Dict[str, Any] = {"step": step, "lr": trainer.lr, "loss_scale": out.loss_scale, "skipped": bool(out.skipped)}
    if out.loss is not None:
      row["loss"] = float(out.loss)
    if out.grad_norm is not None:
      row["grad_norm"] = float(out.grad_norm)
    hists: Dict[str, torch.Tensor] = {}
    g = Graph.get()
    for name, value in merged_summaries(out).items():
      info = g.summary_map[name]
      if info.summary_type == SUMMARY_SCALAR_TYPE:
        row[name] = float(value)
      else:
        v = value.detach().float().flatten().cpu()
        hists[name] = v
        row[name] = {"count": int(v.numel()), "min": float(v.min()), "max": float(v.max()), "mean": float(v.mean())} if v.numel() else {"count": 0}
    self._file.write(json.dumps(row) + "\n")
    self._file.flush()
    if self._tb is not None:
      for k, v in row.items():
        if k not in ("step", "skipped") and isinstance(v, (int, float)):
          self._tb.add_scalar(k, v, step)
      for k, v in hists.items():
        if v.numel():
          self._tb.add_histogram(k, v, step)
      self._tb.flush()

  def close(self) -> None:
    if self._file is not None:
      self._file.close()
      self._file = None
    if self._tb is not None:
      self._tb.close()
      self._tb = None
