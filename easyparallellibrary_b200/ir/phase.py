"""Model phases (reference ``epl/ir/phase.py:21-53``).

In an eager runtime a phase is *where the engine currently is* in a step —
forward call, autograd backward, optimizer apply, checkpoint IO — not a tag
inferred from graph structure, so the context manager is all that is needed.
"""
from __future__ import annotations

import enum
import threading


class ModelPhase(enum.Enum):
  FORWARD = "forward"
  BACKWARD = "backward"
  APPLY = "apply"
  SAVE_AND_RESTORE = "save_and_restore"
  MICRO_BATCH_CLONE = "micro_batch_clone"   # kept for API familiarity; unused by the eager engine
  REPLICATED = "replicated"
  ADD_FUNCTION = "add_function"

  def __call__(self):
    return _PhaseScope(self)


_tls = threading.local()
_listeners = []          # callables (phase, "enter" | "exit"): profilers sample at phase boundaries (profiler/memory.py)


def add_phase_listener(fn) -> None:
  if fn not in _listeners:
    _listeners.append(fn)


def remove_phase_listener(fn) -> None:
  if fn in _listeners:
    _listeners.remove(fn)


def current_phase() -> ModelPhase:
  return getattr(_tls, "phase", ModelPhase.FORWARD)


class _PhaseScope(object):
  def __init__(self, phase: ModelPhase):
    self._phase = phase
    self._prev = None

  def __enter__(self):
    self._prev = current_phase()
    _tls.phase = self._phase
    for fn in _listeners:
      fn(self._phase, "enter")
    return self._phase

  def __exit__(self, *exc):
    for fn in _listeners:
      fn(self._phase, "exit")
    _tls.phase = self._prev
    return False


def phase_scope(phase: ModelPhase) -> _PhaseScope:
  return _PhaseScope(phase)
