"""``train`` / ``evaluate`` / ``train_and_evaluate`` around a :class:`Trainer`.

Users of the reference drive their models through ``tf.estimator`` (``train``, ``evaluate``, ``train_and_evaluate`` under EPL's
hooks, reference ``tests/estimator_test.py:95-175``, ``examples/bert/run_squad.py:1209-1252``).  The TF plumbing is not carried
over but the three entry points are: plain loops over any iterable of per-replica batches that

* stop at ``max_steps`` *global* steps, so a job restarted by ``epl-launch --max_restarts`` (or by hand) continues where its last
  checkpoint left off instead of starting over;
* save a checkpoint every ``save_every`` steps and at the end (``runtime/saver.py``: rank 0 writes the model, every rank its
  optimizer shard) and resume from ``checkpoint_dir`` when one is there;
* evaluate without parallelising (reference ``ir/graph.py:926-933``) and end every evaluation with a barrier — the reference's
  ``_sync_signal`` (``parallel/hooks.py:915-933``) — so no rank runs ahead into training collectives while another still evaluates;
* run the trainer's hooks (``before_step`` / ``after_step``: profilers, watchdog, timeline) exactly as ``Trainer.step`` does.
"""
from __future__ import annotations

import os
from typing import Any, Callable, Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import torch

from easyparallellibrary_b200.utils.logging import get_logger


def _as_batch(item) -> Tuple[Any, ...]:
  return tuple(item) if isinstance(item, (tuple, list)) else (item,)


def _cycle(data: Iterable) -> Iterator:
  """Iterate ``data`` forever, calling ``set_epoch`` (ShardedFileDataset, DistributedSampler) between passes when there is one."""
  epoch = 0
  while True:
    for target in (data, getattr(data, "dataset", None), getattr(data, "sampler", None)):
      if target is not None and hasattr(target, "set_epoch"):
        target.set_epoch(epoch)
    empty = True
    for item in data:
      empty = False
      yield item
    if empty:
      raise ValueError("training data is empty")
    epoch += 1


def _barrier() -> None:
  import torch.distributed as dist
  if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
    dist.barrier()


def train(trainer, data: Iterable, max_steps: int, checkpoint_dir: Optional[str] = None, save_every: int = 0,
          log_every: int = 0, on_step: Optional[Callable[[int, Any], None]] = None) -> List[float]:
  """Train until ``trainer.global_step == max_steps``.  Returns the losses of the steps run by this call."""
  from easyparallellibrary_b200.runtime import saver
  if checkpoint_dir and trainer.global_step == 0 and os.path.isdir(checkpoint_dir) and os.listdir(checkpoint_dir):
    step = saver.load_checkpoint(trainer, checkpoint_dir)
    get_logger().info("resumed from %s at global step %d", checkpoint_dir, step)
  rank = int(os.environ.get("RANK", "0"))
  losses: List[float] = []
  it = data if hasattr(data, "__next__") else _cycle(data)       # an iterator keeps its position across calls (train_and_evaluate)
  while trainer.global_step < max_steps:
    out = trainer.step(*_as_batch(next(it)))
    if out.loss is not None:
      losses.append(float(out.loss))
    step = trainer.global_step
    if on_step is not None:
      on_step(step, out)
    if log_every and rank == 0 and step % log_every == 0 and losses:
      print("step %d loss %.4f%s" % (step, losses[-1], " (skipped: loss scale %g)" % out.loss_scale if out.skipped else ""), flush=True)
    if checkpoint_dir and save_every and step % save_every == 0 and not out.skipped:
      saver.save_checkpoint(trainer, checkpoint_dir)
  if checkpoint_dir:
    saver.save_checkpoint(trainer, checkpoint_dir)
  return losses


@torch.no_grad()
def evaluate(trainer, data: Iterable, metric_fn: Optional[Callable[[Any, Tuple[Any, ...]], Dict[str, float]]] = None,
             max_batches: int = 0, reduce: bool = True) -> Dict[str, float]:
  """Forward-only pass over ``data``.  With ``metric_fn(output, batch) -> {name: value}`` the values are averaged over batches
  (``output`` is what ``Trainer.eval_step`` returns: the loss when the batch carries labels, the model output otherwise; ``None`` on
  the ranks of a pipeline that do not hold the last stage — those batches are skipped); without it the mean loss is reported.

  ``reduce`` (default): sums and batch counts are merged over all ranks, so every rank — also the first stage of a pipeline,
  also each replica that evaluated its own shard of the data — returns the job-wide averages.  Either way the call ends with a
  collective (the reference's ``_sync_signal`` barrier)."""
  sums: Dict[str, float] = {}
  n = 0
  for i, item in enumerate(data):
    if max_batches and i >= max_batches:
      break
    batch = _as_batch(item)
    out = trainer.eval_step(*batch)
    if out is None:
      continue
    vals = metric_fn(out, batch) if metric_fn is not None else {"loss": float(out)}
    for k, v in vals.items():
      sums[k] = sums.get(k, 0.0) + float(v)
    n += 1
  import torch.distributed as dist
  if reduce and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
    parts: List[Any] = [None] * dist.get_world_size()
    dist.all_gather_object(parts, (sums, n))
    sums, n = {}, 0
    for part_sums, part_n in parts:
      n += part_n
      for k, v in part_sums.items():
        sums[k] = sums.get(k, 0.0) + v
  else:
    _barrier()
  res = {k: v / max(n, 1) for k, v in sums.items()}
  res["batches"] = n
  return res


def train_and_evaluate(trainer, train_data: Iterable, eval_data: Iterable, max_steps: int, eval_every: int,
                       metric_fn: Optional[Callable[[Any, Tuple[Any, ...]], Dict[str, float]]] = None,
                       checkpoint_dir: Optional[str] = None, save_every: int = 0, log_every: int = 0,
                       max_eval_batches: int = 0) -> List[Dict[str, float]]:
  """Alternate ``eval_every`` training steps with an evaluation until ``max_steps``; returns the evaluation results in order
  (each with the ``global_step`` it was taken at)."""
  if eval_every <= 0:
    raise ValueError("eval_every must be positive")
  history: List[Dict[str, float]] = []
  train_data = train_data if hasattr(train_data, "__next__") else _cycle(train_data)
  while True:
    target = min(max_steps, (trainer.global_step // eval_every + 1) * eval_every)
    train(trainer, train_data, target, checkpoint_dir=checkpoint_dir, save_every=save_every, log_every=log_every)
    res = evaluate(trainer, eval_data, metric_fn, max_batches=max_eval_batches)
    res["global_step"] = trainer.global_step
    history.append(res)
    if trainer.global_step >= max_steps:
      return history
