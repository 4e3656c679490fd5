"""Contiguous partitioning helpers used by the auto-stage planner, ZeRO
ownership, grouped optimizer apply and checkpoint selection.

Behavioural parity with ``epl/parallel/partitioner.py``:
``partition_buckets`` (26-41) — greedy fill, ``None`` when more than
``num_stages`` buckets would be needed; ``partition_balance`` (44-69) — smallest
bucket bound that fits; ``find_repeated_blocks`` (79-121) — scopes at the same
depth with identical type histograms, ``>= min_dup`` repeats; ``partition_stages``
(124-164) — always exactly ``num_stages`` contiguous groups.

The search itself is different: instead of scanning every integer bound the
min-max bound is found by bisection on the prefix sums (O(n log W)), and
``partition_stages`` then repairs the group count, so million-parameter lists
partition in microseconds.
"""
from __future__ import annotations

from collections import Counter, OrderedDict, defaultdict
from typing import Any, List, Optional, Sequence, Tuple


def partition_buckets(weights: Sequence[float], bucket_size: float, num_stages: int) -> Optional[List[Tuple[int, float]]]:
  """Greedy left-to-right fill.  Returns ``[(start_index, total), ...]`` or ``None``."""
  out: List[List[float]] = [[0, 0]]
  for i, w in enumerate(weights):
    start, total = out[-1]
    if total + w > bucket_size:
      if total == 0:
        out[-1][1] = w
      else:
        out.append([i, w])
        if len(out) > num_stages:
          return None
    else:
      out[-1][1] = total + w
  return [(int(s), t) for s, t in out]


def _fits(weights: Sequence[float], bound: float, parts: int) -> bool:
  used, cur = 1, 0.0
  for w in weights:
    if w > bound:
      return False
    if cur + w > bound:
      used += 1
      cur = w
      if used > parts:
        return False
    else:
      cur += w
  return True


def min_max_bound(weights: Sequence[float], parts: int) -> float:
  """Smallest B such that ``weights`` splits into <= parts contiguous groups each <= B."""
  if not weights:
    return 0.0
  lo, hi = float(max(weights)), float(sum(weights))
  if all(float(w).is_integer() for w in weights):
    lo_i, hi_i = int(lo), int(hi)
    while lo_i < hi_i:
      mid = (lo_i + hi_i) // 2
      if _fits(weights, mid, parts):
        hi_i = mid
      else:
        lo_i = mid + 1
    return float(lo_i)
  for _ in range(60):
    mid = (lo + hi) / 2
    if _fits(weights, mid, parts):
      hi = mid
    else:
      lo = mid
  return hi


def partition_balance(items: Sequence[Any], weights: Sequence[float], num_stages: int) -> List[List[Any]]:
  """Balanced contiguous split into *at most* ``num_stages`` groups."""
  items = list(items)
  if num_stages <= 1 or not items:
    return [items]
  bound = min_max_bound(weights, num_stages)
  groups: List[List[Any]] = [[]]
  cur = 0.0
  for it, w in zip(items, weights):
    if groups[-1] and cur + w > bound:
      groups.append([])
      cur = 0.0
    groups[-1].append(it)
    cur += w
  return groups


def partition_stages(items: Sequence[Any], weights: Sequence[float], num_stages: int) -> List[List[Any]]:
  """Exactly ``num_stages`` contiguous groups (some empty iff ``len(items) < num_stages``)."""
  if num_stages <= 0:
    raise ValueError("partition_stages requires num_stages>=1, got {}".format(num_stages))
  items = list(items)
  if num_stages == 1:
    return [items]
  if len(items) <= num_stages:
    return [[it] for it in items] + [[] for _ in range(num_stages - len(items))]
  weights = [max(float(w), 1e-12) for w in weights]
  groups = partition_balance(items, weights, num_stages)
  # repair: the balanced split may use fewer groups; halve the heaviest multi-item group until exact
  pos = 0
  spans = []
  for g in groups:
    spans.append((pos, pos + len(g)))
    pos += len(g)
  while len(spans) < num_stages:
    best, best_w = -1, -1.0
    for i, (a, b) in enumerate(spans):
      if b - a > 1:
        w = sum(weights[a:b])
        if w > best_w:
          best, best_w = i, w
    a, b = spans[best]
    half, acc, cut = best_w / 2, 0.0, a + 1
    for j in range(a, b - 1):
      acc += weights[j]
      cut = j + 1
      if acc >= half:
        break
    spans[best:best + 1] = [(a, cut), (cut, b)]
  return [items[a:b] for a, b in spans]


def group_list(items: Sequence[Any], num_groups: int, weights: Optional[Sequence[float]] = None) -> List[List[Any]]:
  """Split into ``num_groups`` size-balanced contiguous groups (ZeRO / grouped apply)."""
  if weights is None:
    weights = [1.0] * len(items)
  return partition_stages(items, weights, num_groups)


# ------------------------------------------------------------------------------------------
# repeated blocks
# ------------------------------------------------------------------------------------------
def _histogram_key(nodes: Sequence[Any], min_types: int) -> Optional[str]:
  c = Counter(getattr(n, "type", type(n).__name__) for n in nodes)
  if len(c) < min_types:
    return None
  return repr(sorted(c.items()))


def find_repeated_blocks(nodes: Sequence[Any], max_depth: int = 20, min_dup: int = 4, min_ops: int = 2,
                         min_types: int = 2, sep: str = ".") -> List[List[Any]]:
  """Find repeated model blocks: groups of nodes under sibling scopes whose
  type histograms are identical.  Returns the blocks in execution order.

  Module-level IR nodes are ~50x coarser than TF ops, hence the smaller
  ``min_ops`` / ``min_types`` defaults (reference: 20 ops, >5 op types).
  """
  order = {id(n): i for i, n in enumerate(nodes)}
  depth_scopes: "OrderedDict[int, OrderedDict[str, List[Any]]]" = OrderedDict()
  real_max = max((len(n.name.split(sep)) for n in nodes), default=0)
  for depth in range(1, min(max_depth, real_max) + 1):
    table: "OrderedDict[str, List[Any]]" = OrderedDict()
    for n in nodes:
      parts = n.name.split(sep)
      if len(parts) < depth:
        continue
      table.setdefault(sep.join(parts[:depth]), []).append(n)
    depth_scopes[depth] = table
  found: List[List[Any]] = []
  covered: List[str] = []
  for depth, table in depth_scopes.items():
    similar = defaultdict(list)
    fresh = False
    for scope, members in table.items():
      if any(scope == c or scope.startswith(c + sep) for c in covered):
        continue
      fresh = True
      key = _histogram_key(members, min_types)
      if key:
        similar[key].append((scope, members))
    if not fresh:
      break
    for key, blocks in similar.items():
      if len(blocks) >= min_dup and len(blocks[0][1]) >= min_ops:
        for scope, members in blocks:
          found.append(members)
          covered.append(scope)
  found.sort(key=lambda b: order[id(b[0])])
  return found
