"""Stack of active strategies with the reference's nesting rules
(``epl/strategies/strategy_context.py:34-54``): a strategy may not nest inside
one of the same type, nothing nests inside ``split``, and ``split`` may not
nest inside ``replicate``."""
from __future__ import annotations

from typing import List, Optional

from easyparallellibrary_b200.strategies.base import ParallelStrategy, Replicate, Split


class StrategyContext(object):
  def __init__(self):
    self._stack: List[ParallelStrategy] = []
    self._default: Optional[ParallelStrategy] = None
    self._seen = 0
    self.update_flag = False     # the next captured node opens a new taskgraph

  # -- stack ------------------------------------------------------------------
  def push(self, strategy: ParallelStrategy) -> None:
    for active in self._stack:
      if type(active) is type(strategy):
        raise RuntimeError("Can't nest strategy %s inside a strategy of the same type." % type(strategy).__name__)
      if isinstance(active, Split):
        raise RuntimeError("Can't nest any strategy inside split.")
      if isinstance(active, Replicate) and isinstance(strategy, Split):
        raise RuntimeError("Can't nest split inside replicate.")
    self._register(strategy)
    self._stack.append(strategy)

  def pop(self, strategy: ParallelStrategy) -> None:
    if not self._stack or self._stack[-1] is not strategy:
      raise RuntimeError("Strategy scopes must be exited in reverse order of entry.")
    self._stack.pop()
    self.update_flag = True

  def _register(self, strategy: ParallelStrategy) -> None:
    strategy.index = self._seen
    self._seen += 1
    self.update_flag = True
    if isinstance(strategy, Split):
      from easyparallellibrary_b200.env import Env
      cluster = Env.get().cluster
      if cluster is not None and cluster.virtual_devices:
        strategy.devices = cluster.virtual_devices[0].all_devices

  # -- default strategy -----------------------------------------------------------
  @property
  def default_strategy(self) -> Optional[ParallelStrategy]:
    return self._default

  @default_strategy.setter
  def default_strategy(self, strategy: Optional[ParallelStrategy]) -> None:
    if strategy is not None:
      if not isinstance(strategy, Replicate):
        raise ValueError("Only replicate can be the default strategy.")
      strategy.is_default = True
      self._register(strategy)
    self._default = strategy

  def suspend_default(self):
    prev, self._default_suspended = self._default, True
    return prev

  def resume_default(self, prev) -> None:
    self._default_suspended = False

  # -- queries --------------------------------------------------------------------
  @property
  def state(self) -> List[ParallelStrategy]:
    if self._stack:
      return list(self._stack)
    if self._default is not None and not getattr(self, "_default_suspended", False):
      return [self._default]
    return []

  @property
  def current(self) -> Optional[ParallelStrategy]:
    st = self.state
    return st[-1] if st else None

  @property
  def split_strategy(self) -> Optional[Split]:
    for s in self._stack:
      if isinstance(s, Split):
        return s
    return None

  @property
  def replicate_strategy(self) -> Optional[Replicate]:
    for s in self.state:
      if isinstance(s, Replicate):
        return s
    return None

  @property
  def identity(self) -> int:
    return hash(tuple((type(s).__name__, s.identity, s.name) for s in self.state))

  def __len__(self) -> int:
    return len(self._stack)
