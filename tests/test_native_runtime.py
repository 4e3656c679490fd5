"""The C++ runtime (csrc/runtime.cpp) against the Python reference implementations."""
import random

import pytest
import torch

from easyparallellibrary_b200.communicators.coalescing import plan_buckets
from easyparallellibrary_b200.parallel import partitioner, schedule
from easyparallellibrary_b200.runtime import native

pytestmark = pytest.mark.skipif(not native.available(), reason="native runtime not built")


@pytest.mark.parametrize("policy", ["PreferForward", "PreferBackward", "PreferBackwardOptimizer"])
def test_schedule_generator_and_simulator_match(policy):
  for S, M in ((1, 3), (2, 4), (3, 5), (4, 8), (8, 16), (4, 2)):
    for prefetch in (0, 1, 2):
      for s in range(S):
        a = native.schedule_stage(policy, s, S, M, prefetch)
        b = [(i.op, i.mb) for i in schedule.build_stage_program(policy, s, S, M, prefetch)]
        assert a == b, (policy, S, M, prefetch, s)
    ok, mk, bub, infl = native.schedule_simulate(policy, S, M, 1, 1.0, 2.0, 0.1)
    r = schedule.simulate(schedule.build_programs(policy, S, M, 1), 1.0, 2.0, 0.1)
    assert ok and r.ok and abs(mk - r.makespan) < 1e-9 and abs(bub - r.bubble_fraction) < 1e-9 and infl == r.max_in_flight


def test_bucket_planner_matches():
  rnd = random.Random(0)
  dts = [torch.float32, torch.bfloat16, torch.float16]
  for trial in range(200):
    n = rnd.randrange(1, 40)
    sizes = [rnd.choice([0, 4, 64, 1000, 1 << 20]) * rnd.randrange(1, 5) for _ in range(n)]
    kinds = [rnd.choice(dts[:rnd.randrange(1, 4)]) for _ in range(n)]
    for k in (1, 2, 5, 9):
      assert plan_buckets(sizes, kinds, k, use_native=True) == plan_buckets(sizes, kinds, k, use_native=False), (sizes, kinds, k)


def test_partitioner_matches_contract():
  rnd = random.Random(1)
  for trial in range(100):
    n = rnd.randrange(1, 200)
    w = [rnd.randrange(1, 1000) for _ in range(n)]
    parts = rnd.randrange(1, 20)
    starts = native.partition_stages(w, parts)
    assert len(starts) == parts + 1 and starts[0] == 0 and starts[-1] == n and starts == sorted(starts)
    py = partitioner.partition_stages(list(range(n)), w, parts)
    if n > parts:
      mx_native = max(sum(w[a:b]) for a, b in zip(starts, starts[1:]))
      mx_py = max(sum(w[i] for i in g) for g in py if g)
      assert mx_native <= 1.25 * mx_py + 1 and mx_py <= 1.25 * mx_native + 1
