"""Pipeline schedules as explicit per-stage instruction lists.

The reference expresses its three policies as control-dependency edges between
cloned micro-batch sub-graphs (``epl/strategies/scheduler.py:36-116``):

* ``PreferForward``  — GPipe: every forward, then every backward;
* ``PreferBackward`` — 1F1B: stage ``s`` of ``S`` keeps at most ``S - s``
  forwards in flight (``F(m)`` waits for ``B(m - (S - s))``);
* ``PreferBackwardOptimizer`` — 1F1B with one more forward in flight on every
  stage but the last (``F(m)`` waits for ``B(m - (S - s) - 1)``).

Here the same policies produce an ordered program per stage which the pipeline
engine interprets: compute instructions ``F(m)``/``B(m)`` interleaved with the
NCCL p2p instructions that move activations / activation-gradients.  Receives
are hoisted ``prefetch`` compute slots ahead of their consumer (double
buffering), so a transfer overlaps the previous micro-batch's math.  A
deterministic simulator checks any program set for deadlock and reports the
bubble fraction; the C++ runtime (``csrc/runtime.cpp``) implements the same
generator and simulator for the hot path and is cross-checked in the tests.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

from easyparallellibrary_b200.utils import constant

# opcodes (shared with csrc/runtime.cpp)
F, B, SEND_F, RECV_F, SEND_B, RECV_B, REDUCE, APPLY = range(8)
OP_NAMES = ("F", "B", "SEND_F", "RECV_F", "SEND_B", "RECV_B", "REDUCE", "APPLY")


@dataclass(frozen=True)
class Instr:
  op: int
  mb: int = -1

  def __repr__(self):
    return "%s(%d)" % (OP_NAMES[self.op], self.mb) if self.mb >= 0 else OP_NAMES[self.op]


def in_flight_cap(policy: str, stage: int, num_stages: int, num_micro_batch: int) -> int:
  policy = policy.lower()
  if policy == constant.SCHEDULE_PREFER_FORWARD:
    return num_micro_batch
  if policy == constant.SCHEDULE_PREFER_BACKWARD:
    return max(1, min(num_stages - stage, num_micro_batch))
  if policy == constant.SCHEDULE_PREFER_BACKWARD_OPT:
    extra = 0 if stage == num_stages - 1 else 1
    return max(1, min(num_stages - stage + extra, num_micro_batch))
  raise RuntimeError("Unknown scheduler {}, current supported schedulers are {}".format(
      policy, [constant.SCHEDULE_PREFER_FORWARD, constant.SCHEDULE_PREFER_BACKWARD, constant.SCHEDULE_PREFER_BACKWARD_OPT]))


def compute_order(policy: str, stage: int, num_stages: int, num_micro_batch: int) -> List[Instr]:
  """The F/B order of one stage."""
  cap = in_flight_cap(policy, stage, num_stages, num_micro_batch)
  f = b = 0
  out: List[Instr] = []
  while b < num_micro_batch:
    if f < num_micro_batch and f - b < cap:
      out.append(Instr(F, f))
      f += 1
    else:
      out.append(Instr(B, b))
      b += 1
  return out


def build_stage_program(policy: str, stage: int, num_stages: int, num_micro_batch: int, prefetch: int = 1,
                        with_reduce: bool = True) -> List[Instr]:
  """Full program of one stage: compute + p2p, receives hoisted by ``prefetch`` compute slots."""
  order = compute_order(policy, stage, num_stages, num_micro_batch)
  first, last = stage == 0, stage == num_stages - 1
  slots: List[List[Instr]] = []          # per compute slot: [recvs..., compute, sends...]
  for ins in order:
    pre, post = [], []
    if ins.op == F:
      if not first:
        pre.append(Instr(RECV_F, ins.mb))
      if not last:
        post.append(Instr(SEND_F, ins.mb))
    else:
      if not last:
        pre.append(Instr(RECV_B, ins.mb))
      if not first:
        post.append(Instr(SEND_B, ins.mb))
    slots.append([pre, ins, post])
  prog: List[Instr] = []
  posted = set()
  f_slot = {ins.mb: n for n, (_, ins, _) in enumerate(slots) if ins.op == F}
  for i, (pre, ins, post) in enumerate(slots):
    # post this slot's receives (if not already) plus those of the next `prefetch` slots that receive anything;
    # a gradient receive is never hoisted above the forward of its own micro-batch (its buffer mirrors that output)
    ahead, j = 0, i
    while j < len(slots) and ahead <= max(prefetch, 0):
      hoistable = j == i or slots[j][1].op == F or f_slot[slots[j][1].mb] < i
      if (slots[j][0] and hoistable) or j == i:
        for r in slots[j][0]:
          if (r.op, r.mb) not in posted:
            posted.add((r.op, r.mb))
            prog.append(r)
        if j > i:
          ahead += 1
      j += 1
    prog.append(ins)
    prog.extend(post)
  if with_reduce:
    prog.append(Instr(REDUCE))
  prog.append(Instr(APPLY))
  return prog


def build_programs(policy: str, num_stages: int, num_micro_batch: int, prefetch: int = 1) -> List[List[Instr]]:
  return [build_stage_program(policy, s, num_stages, num_micro_batch, prefetch) for s in range(num_stages)]


# ------------------------------------------------------------------------------------------
# simulator
# ------------------------------------------------------------------------------------------
@dataclass
class SimResult:
  ok: bool
  makespan: float
  bubble_fraction: float
  max_in_flight: List[int]
  timeline: List[List[Tuple[str, float, float]]]
  reason: str = ""


def simulate(programs: Sequence[Sequence[Instr]], t_fwd: float = 1.0, t_bwd: float = 2.0, t_p2p: float = 0.0) -> SimResult:
  """Event simulation with asynchronous receives (posted early, waited on by the
  consuming compute instruction) and asynchronous sends."""
  S = len(programs)
  pc = [0] * S
  clock = [0.0] * S
  sent: Dict[Tuple[int, int, int], float] = {}     # (op, dst_stage, mb) -> arrival time
  in_flight = [0] * S
  max_in_flight = [0] * S
  busy = [0.0] * S
  timeline: List[List[Tuple[str, float, float]]] = [[] for _ in range(S)]
  progress = True
  while progress:
    progress = False
    for s in range(S):
      while pc[s] < len(programs[s]):
        ins = programs[s][pc[s]]
        if ins.op in (RECV_F, RECV_B, REDUCE, APPLY):
          pc[s] += 1               # posting a receive never blocks
          progress = True
          continue
        if ins.op in (SEND_F, SEND_B):
          dst = s + 1 if ins.op == SEND_F else s - 1
          sent[(ins.op, dst, ins.mb)] = clock[s] + t_p2p
          pc[s] += 1
          progress = True
          continue
        # compute: needs its input to have arrived
        need = None
        if ins.op == F and s > 0:
          need = (SEND_F, s, ins.mb)
        if ins.op == B and s < S - 1:
          need = (SEND_B, s, ins.mb)
        if need is not None:
          if need not in sent:
            break
          # the matching receive must have been posted earlier in this program
          rop = RECV_F if ins.op == F else RECV_B
          if not any(p.op == rop and p.mb == ins.mb for p in programs[s][:pc[s]]):
            return SimResult(False, 0.0, 1.0, max_in_flight, timeline, "stage %d computes %r before posting its receive" % (s, ins))
          clock[s] = max(clock[s], sent[need])
        dur = t_fwd if ins.op == F else t_bwd
        timeline[s].append((repr(ins), clock[s], clock[s] + dur))
        clock[s] += dur
        busy[s] += dur
        in_flight[s] += 1 if ins.op == F else -1
        max_in_flight[s] = max(max_in_flight[s], in_flight[s])
        pc[s] += 1
        progress = True
  if any(pc[s] < len(programs[s]) for s in range(S)):
    stuck = [(s, programs[s][pc[s]]) for s in range(S) if pc[s] < len(programs[s])]
    return SimResult(False, 0.0, 1.0, max_in_flight, timeline, "deadlock at %r" % (stuck,))
  makespan = max(clock) if clock else 0.0
  total = makespan * S
  bubble = 1.0 - (sum(busy) / total) if total > 0 else 0.0
  return SimResult(True, makespan, bubble, max_in_flight, timeline)


def get_scheduler(name: str):
  """Name -> program builder (reference ``scheduler.get_scheduler`` 126-131)."""
  key = name.lower()
  in_flight_cap(key, 0, 1, 1)   # validates the name
  return lambda num_stages, num_micro_batch, prefetch=1: build_programs(key, num_stages, num_micro_batch, prefetch)
