"""Input-file sharding across workers (reference ``graph_editor.py:787-854``,
``fetch_slice_objects_proportion_to_local_num_replicas``; cases in ``tests/utils_test.py:191-339``).

Files are divided in proportion to each worker's number of local model replicas (gcd-normalised).
``drop_last_files`` trims the remainder so every replica sees the same count; ``unbalanced_io_slicing``
hands the remainder to the first workers; with neither, too few files are handled by giving every worker the
whole list rotated by its index (duplicate + shuffle fallback).
"""
from __future__ import annotations

from functools import reduce
from math import gcd
from typing import List, Sequence, TypeVar

T = TypeVar("T")


def slice_files(files: Sequence[T], replicas_per_worker: Sequence[int], worker_index: int, drop_last_files: bool = False,
                unbalanced_io_slicing: bool = False) -> List[T]:
  files = list(files)
  if not replicas_per_worker or sum(replicas_per_worker) <= 0:
    raise ValueError("replicas_per_worker must contain positive counts")
  g = reduce(gcd, [r for r in replicas_per_worker if r > 0])
  shares = [r // g for r in replicas_per_worker]
  unit = sum(shares)
  n = len(files)
  per_unit, rem = divmod(n, unit)
  if per_unit == 0 or (rem and not (drop_last_files or unbalanced_io_slicing)):
    if per_unit == 0 and drop_last_files:
      raise RuntimeError("Files number %d is less than the number of data slices %d." % (n, unit))
    if rem == 0 and per_unit > 0:
      pass
    else:
      # fallback: every worker reads all files, in a different order
      k = worker_index % max(n, 1)
      return files[k:] + files[:k]
  counts = [s * per_unit for s in shares]
  if rem and unbalanced_io_slicing:
    i = 0
    while rem > 0:
      take = min(shares[i % len(shares)], rem)
      counts[i % len(shares)] += take
      rem -= take
      i += 1
  start = sum(counts[:worker_index])
  return files[start:start + counts[worker_index]]
