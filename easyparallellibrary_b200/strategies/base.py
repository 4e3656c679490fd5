"""Parallel strategy annotations.

``with epl.replicate(n):`` / ``with epl.split(n):`` tag everything built in the
scope (parameters, buffers, sub-modules — see ``ir/capture.py``) and everything
called in the scope (forward-time scopes) with a *taskgraph*.  Parity:
``epl/strategies/parallel_strategy.py:28-75`` (device_count + name, scope
identity from the user's call site, default strategy suspended while an
explicit scope is open, forbidden under ``auto.auto_parallel``).
"""
from __future__ import annotations

import sys
from typing import Optional


def _call_site(depth: int = 4) -> str:
  """A short call-stack fingerprint: re-entering the *same* ``with`` line maps
  back to the same taskgraph (weight sharing / module reuse)."""
  frames = []
  f = sys._getframe(2)
  while f is not None and len(frames) < depth:
    name = f.f_code.co_filename
    if "easyparallellibrary_b200/strategies" not in name.replace("\\", "/"):
      frames.append("%s:%d" % (name, f.f_lineno))
    f = f.f_back
  return "|".join(frames)


class ParallelStrategy(object):
  kind = "base"

  def __init__(self, device_count: Optional[int] = None, name: Optional[str] = None):
    if device_count is not None and (not isinstance(device_count, int) or device_count < 1):
      raise ValueError("device_count must be a positive int, got %r" % (device_count,))
    self.device_count = device_count
    self.name = name
    self.index = -1            # position among all strategies seen so far
    self.is_default = False
    self.identity = _call_site()
    self._suspended_default = None

  # -- context manager --------------------------------------------------------
  def __enter__(self):
    from easyparallellibrary_b200.env import Env
    env = Env.get()
    if env.config.auto.auto_parallel:
      raise RuntimeError("Strategy annotations are not allowed when auto.auto_parallel is enabled.")
    ctx = env.strategy_context
    self._suspended_default = ctx.suspend_default()
    ctx.push(self)
    return self

  def __exit__(self, exc_type, exc, tb):
    from easyparallellibrary_b200.env import Env
    ctx = Env.get().strategy_context
    ctx.pop(self)
    ctx.resume_default(self._suspended_default)
    self._suspended_default = None
    return False

  def __repr__(self):
    return "%s(device_count=%s, name=%s, index=%d)" % (type(self).__name__, self.device_count, self.name, self.index)


class Replicate(ParallelStrategy):
  """Data parallelism over replicas of ``device_count`` devices each.

  Several consecutive ``replicate`` scopes become pipeline stages when
  ``pipeline.num_micro_batch > 1`` (reference ``strategies/replicate.py:24-41``).
  """
  kind = "replicate"


class Split(ParallelStrategy):
  """Tensor / expert parallelism over ``device_count`` devices
  (reference ``strategies/split.py:24-51``)."""
  kind = "split"

  def __init__(self, device_count=None, name=None):
    super(Split, self).__init__(device_count, name)
    self.is_nested = False
    self.devices = None


def replicate(device_count: Optional[int] = None, name: Optional[str] = None) -> Replicate:
  return Replicate(device_count, name)


def split(device_count: Optional[int] = None, name: Optional[str] = None) -> Split:
  return Split(device_count, name)
