"""Learning-rate schedules: callables ``global_step -> lr`` to pass as ``Trainer(..., lr=schedule)``.

The reference's examples build these from TF ops — ``tf.train.exponential_decay`` (``tests/multi_optimizer_test.py:59-62``),
BERT's linear warm-up + polynomial decay (``examples/bert/optimization.py:29-58``).  Here a schedule is a pure function of the
global step, evaluated on the host before each step; the fused / CUDA-graph paths read the value from device memory, so a changing
learning rate never re-captures or re-launches anything differently.
"""
from __future__ import annotations

import math
from typing import Callable

Schedule = Callable[[int], float]


def constant(lr: float) -> Schedule:
  return lambda step: float(lr)


def exponential_decay(lr: float, decay_steps: int, decay_rate: float, staircase: bool = False) -> Schedule:
  """``lr * decay_rate ** (step / decay_steps)`` (integer division with ``staircase``)."""
  def f(step: int) -> float:
    e = step // decay_steps if staircase else step / float(decay_steps)
    return float(lr) * decay_rate ** e
  return f


def polynomial_decay(lr: float, decay_steps: int, end_lr: float = 0.0, power: float = 1.0) -> Schedule:
  def f(step: int) -> float:
    t = min(step, decay_steps) / float(max(decay_steps, 1))
    return (float(lr) - end_lr) * (1.0 - t) ** power + end_lr
  return f


def warmup(schedule: Schedule, warmup_steps: int) -> Schedule:
  """Linear warm-up from 0 over ``warmup_steps`` steps towards ``schedule`` (BERT: ``global_step / warmup_steps * init_lr``)."""
  def f(step: int) -> float:
    if warmup_steps > 0 and step < warmup_steps:
      return schedule(step) * (step + 1) / float(warmup_steps)      # +1: the very first step already moves
    return schedule(step)
  return f


def warmup_linear_decay(lr: float, total_steps: int, warmup_steps: int = 0) -> Schedule:
  """BERT fine-tuning: linear warm-up, then linear decay to 0 at ``total_steps``."""
  return warmup(polynomial_decay(lr, total_steps, 0.0, 1.0), warmup_steps)


def warmup_cosine(lr: float, total_steps: int, warmup_steps: int = 0, min_lr: float = 0.0) -> Schedule:
  """GPT-style: linear warm-up, cosine decay to ``min_lr``."""
  def cos(step: int) -> float:
    t = min(max(step - warmup_steps, 0), max(total_steps - warmup_steps, 1)) / float(max(total_steps - warmup_steps, 1))
    return min_lr + 0.5 * (float(lr) - min_lr) * (1.0 + math.cos(math.pi * t))
  return warmup(cos, warmup_steps)
