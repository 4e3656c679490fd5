"""Step watchdog: failure *detection* for hung collectives.

The reference has none — a dead peer leaves every other worker blocked inside an NCCL kernel forever
(``SURVEY.md`` 5.3: no heartbeat, no collective timeout, ``NcclCommWrapper::Abort`` never called).  Here a hook arms a
timer around every training step; if a step does not finish within ``timeout_s`` the watchdog

1. dumps the Python stack of every thread (so the log shows *which* collective / transfer is stuck),
2. aborts the in-tree NCCL communicators (``epl_comm_abort``), which unblocks device-side waits, and
3. exits the process with a non-zero code so the launcher tears the job down and it can be restarted from the last
   checkpoint (the recovery story is the same as the reference's: resume from ``runtime/saver``).

``on_timeout`` can be replaced (the tests use a callback instead of exiting).
"""
from __future__ import annotations

import faulthandler
import os
import sys
import threading
import time
from typing import Callable, Optional


def _default_timeout_action(step: int, elapsed: float) -> None:  # pragma: no cover - terminates the process
  sys.stderr.write("[epl watchdog] step %d has been running for %.0f s: dumping stacks, aborting communicators\n" % (step, elapsed))
  faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
  try:
    from easyparallellibrary_b200.communicators import native
    for be in list(getattr(native, "_LIVE", [])):
      be.abort()
  except Exception:
    pass
  os._exit(70)


class StepWatchdog(object):
  """Trainer hook (``trainer.hooks.append(StepWatchdog(600))``)."""

  def __init__(self, timeout_s: float = 600.0, on_timeout: Optional[Callable[[int, float], None]] = None, poll_s: float = 0.5):
    self.timeout_s, self.poll_s = float(timeout_s), float(poll_s)
    self.on_timeout = on_timeout or _default_timeout_action
    self._started_at: Optional[float] = None
    self._step = 0
    self._fired_for = -1
    self._halt = threading.Event()
    self._thread = threading.Thread(target=self._run, name="epl-step-watchdog", daemon=True)
    self._thread.start()

  def _run(self) -> None:
    while not self._halt.wait(self.poll_s):
      t0 = self._started_at
      if t0 is not None and self._fired_for != self._step:
        elapsed = time.monotonic() - t0
        if elapsed > self.timeout_s:
          self._fired_for = self._step
          self.on_timeout(self._step, elapsed)

  def before_step(self, trainer) -> None:
    self._started_at = time.monotonic()

  def after_step(self, trainer, out) -> None:
    self._started_at = None
    self._step += 1

  def close(self) -> None:
    self._halt.set()
    self._thread.join(timeout=2)
