"""Communicator descriptions (reference ``epl/communicators/options.py``, ``collective_keys.py``).

A *logical* communicator is a name plus an ordered rank list; ``CommunicatorSpec`` carries the knobs that shape the
transport behind it.  ``CollectiveKeys`` hands out deterministic, collision-free keys (store keys for NCCL-id exchange,
signal-pad slots) — the role TF collective group/instance keys play in the reference.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple


@dataclass
class CommunicatorSpec:
  name: str
  ranks: List[int]
  max_splits: int = 5
  num_communicators: int = 2
  enable_fp16: bool = False
  fp16_scale: float = 128.0
  backend: str = "auto"            # auto | torch | native

  def build(self, device=None):
    from easyparallellibrary_b200.communicators.collective_communicator import CollectiveCommunicator
    return CollectiveCommunicator(self.name, self.ranks, self.max_splits, self.num_communicators, self.enable_fp16, self.fp16_scale,
                                  device=device, prefer_native=self.backend == "native")


def build_communicator(spec: CommunicatorSpec, device=None):
  return spec.build(device)


class CollectiveKeys(object):
  def __init__(self):
    self._next: Dict[Tuple[int, ...], int] = {}

  def key(self, ranks: Sequence[int], purpose: str = "") -> str:
    t = tuple(int(r) for r in ranks)
    n = self._next.get(t, 0)
    self._next[t] = n + 1
    return "epl/%s/%s/%d" % (purpose or "comm", "-".join(map(str, t)), n)
