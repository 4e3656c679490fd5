"""Input sharding driven by the ``io.*`` config keys.

Reference: with ``io.slicing`` the graph editor rewrites the dataset ops so that every worker reads its own part of the
file list, in proportion to the model replicas it hosts (``epl/parallel/graph_editor.py:149-215, 787-854``);
``io.drop_last_files`` / ``io.unbalanced_io_slicing`` decide what happens to a remainder.  There is no dataset graph to
rewrite in an eager framework, so the same policy is applied where the file list enters the input pipeline:

* :func:`shard_files` — the files of THIS process, given the parallel plan (one process per GPU: a process hosts one
  model replica slot; ranks that are not the first stage of a pipeline replica read nothing but must still iterate in
  lock-step, so they get their replica's files as well — the engine ignores their batches);
* :class:`ShardedFileDataset` — an ``IterableDataset`` over ``reader(file)`` samples of the sharded list, further split
  over DataLoader workers, with an epoch-seeded shuffle that is identical on the ranks of one replica.

With ``io.slicing = False`` (the default, as in the reference) every process sees the whole list.
"""
from __future__ import annotations

import random
from typing import Any, Callable, Iterable, Iterator, List, Optional, Sequence

import torch
from torch.utils.data import IterableDataset, get_worker_info

from easyparallellibrary_b200.utils.io_slicing import slice_files


def replica_layout(plan=None):
  """``(replica_index, num_replicas)`` of this process for input purposes."""
  from easyparallellibrary_b200.env import Env
  env = Env.get()
  if plan is None:
    plan = getattr(env, "parallel_plan", None)
  if plan is not None:
    pl = plan.placements.get(plan.stage_taskgraphs[0]) or next(iter(plan.placements.values()))
    return int(pl.replica), int(plan.num_replicas)
  cluster = env.cluster
  world = cluster.total_gpu_num if cluster is not None else 1
  rank = cluster.rank if cluster is not None and cluster.rank is not None else 0
  return int(rank), int(max(world, 1))


def shard_files(files: Sequence[Any], plan=None, config=None, replicas_per_worker: Optional[Sequence[int]] = None,
                worker_index: Optional[int] = None) -> List[Any]:
  """The part of ``files`` this process reads under the ``io.*`` settings of ``config`` (default: the active config)."""
  from easyparallellibrary_b200.env import Env
  cfg = config or Env.get().config
  files = list(files)
  if cfg is None or not cfg.io.slicing:
    return files
  if replicas_per_worker is None:
    idx, n = replica_layout(plan)
    replicas_per_worker, worker_index = [1] * n, idx
  return slice_files(files, replicas_per_worker, int(worker_index or 0), drop_last_files=cfg.io.drop_last_files,
                     unbalanced_io_slicing=cfg.io.unbalanced_io_slicing)


class ShardedFileDataset(IterableDataset):
  """``for sample in ShardedFileDataset(files, reader)``: samples of this process's files.

  ``reader(file) -> iterable of samples``.  ``shuffle`` permutes the (already sharded) file order per epoch with a seed that
  does not depend on the rank, so the replicas of one model keep seeing disjoint files."""

  def __init__(self, files: Sequence[Any], reader: Callable[[Any], Iterable[Any]], shuffle: bool = False, seed: int = 0,
               plan=None, config=None):
    super().__init__()
    self.all_files = list(files)
    self.files = shard_files(self.all_files, plan=plan, config=config)
    self.reader, self.shuffle, self.seed, self.epoch = reader, shuffle, seed, 0

  def set_epoch(self, epoch: int) -> None:
    self.epoch = int(epoch)

  def __len__(self) -> int:
    return len(self.files)

  def __iter__(self) -> Iterator[Any]:
    files = list(self.files)
    if self.shuffle:
      random.Random(self.seed + self.epoch).shuffle(files)
    info = get_worker_info()
    if info is not None:                       # DataLoader workers take every num_workers-th file of the shard
      files = files[info.id::info.num_workers]
    for f in files:
      for sample in self.reader(f):
        yield sample


def synthetic_token_files(num_files: int, samples_per_file: int, seq_len: int, vocab: int, seed: int = 0):
  """In-memory stand-in for a tokenised corpus split into files (there is no network or dataset in the sandbox):
  ``(files, reader)`` for :class:`ShardedFileDataset`; file ``i`` deterministically yields ``samples_per_file`` sequences."""
  files = list(range(num_files))

  def reader(i: int):
    g = torch.Generator().manual_seed(seed * 1000003 + int(i))
    for _ in range(samples_per_file):
      yield torch.randint(0, vocab, (seq_len,), generator=g)
  return files, reader
