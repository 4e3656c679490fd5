"""Partitioner, planner, bucket planner and pipeline schedules (reference utils_test.py:340-366,
planner_test.py:48-62, scheduler_test.py:72-130)."""
import random

import pytest

from easyparallellibrary_b200.communicators.coalescing import estimate_split_num_for_comm, plan_buckets
from easyparallellibrary_b200.ir.node import Node
from easyparallellibrary_b200.parallel import partitioner, schedule
from easyparallellibrary_b200.parallel.planner import AutoStageGenerator
from easyparallellibrary_b200.utils import constant


def test_partition_buckets():
  assert partitioner.partition_buckets([1] * 6, 2, 3) == [(0, 2), (2, 2), (4, 2)]
  assert partitioner.partition_buckets([1] * 6, 3, 3) == [(0, 3), (3, 3)]
  w = [1, 7, 4, 3, 1, 11]
  assert partitioner.partition_buckets(w, 1, 3) is None
  assert partitioner.partition_buckets(w, 5, 3) is None
  assert partitioner.partition_buckets(w, 11, 3) == [(0, 8), (2, 8), (5, 11)]


def test_partition_stages_exact_group_count():
  assert partitioner.partition_stages(["a"] * 10, [1] * 10, 2) == [["a"] * 5, ["a"] * 5]
  assert partitioner.partition_stages(["a"] * 10, [1] * 10, 5) == [["a"] * 2] * 5
  for n in range(1, 1024, 8):
    res = partitioner.partition_stages(["a"] * 10, [1] * 10, n)
    assert len(res) == n and [x for g in res for x in g] == ["a"] * 10
  rnd = random.Random(0)
  for n in range(1, 400, 7):
    data = [rnd.randrange(1, 100) for _ in range(512)]
    res = partitioner.partition_stages(data, data, n)
    assert len(res) == n and [x for g in res for x in g] == data
    if n <= 64:
      sums = [sum(g) for g in res]
      assert max(sums) <= 2.2 * sum(data) / n + 100
  res = partitioner.partition_stages(list("abcdef"), [1024, 1, 2, 3, 4, 5], 4)
  assert len(res) == 4 and [x for g in res for x in g] == list("abcdef")
  with pytest.raises(ValueError):
    partitioner.partition_stages([1], [1], 0)


def _transformer_nodes(layers=8):
  nodes = [Node("embed.wte", "Embedding", param_count=1000, flops=0)]
  for i in range(layers):
    for sub, typ, p in (("ln_1", "LayerNorm", 10), ("attn.qkv", "Linear", 300), ("attn.proj", "Linear", 100),
                        ("ln_2", "LayerNorm", 10), ("mlp.fc", "Linear", 400), ("mlp.proj", "Linear", 400)):
      nodes.append(Node("h.%d.%s" % (i, sub), typ, param_count=p, flops=2.0 * p * 64))
  nodes.append(Node("head.ln_f", "LayerNorm", param_count=10))
  nodes.append(Node("head.proj", "Linear", param_count=1000, flops=2.0 * 1000 * 64))
  return nodes


def test_find_repeated_blocks_and_auto_stages():
  nodes = _transformer_nodes(8)
  blocks = partitioner.find_repeated_blocks(nodes, min_dup=4)
  assert len(blocks) == 8 and all(len(b) == 6 for b in blocks)
  assert blocks[0][0].name == "h.0.ln_1" and blocks[-1][-1].name == "h.7.mlp.proj"
  for policy in (constant.STAGE_POLICY_HEURISTIC, constant.STAGE_POLICY_REPEATED_LAYERS, constant.STAGE_POLICY_BALANCE_OP_NUM):
    stages = AutoStageGenerator(policy, num_stages=4).search(nodes)
    assert len(stages) == 4 and [n for s in stages for n in s] == nodes
  stages = AutoStageGenerator(num_stages=4).search(nodes)
  # cuts fall on block boundaries
  for st in stages[1:]:
    assert st[0].name.endswith("ln_1")
  assert sorted(len([n for n in s if n.name.endswith("mlp.fc")]) for s in stages) == [2, 2, 2, 2]


def test_plan_buckets_policy():
  import torch
  sizes = [4 * n for n in (100, 200, 300, 400, 500, 600, 700, 800)]
  plan = plan_buckets(sizes, [torch.float32] * 8, 5)
  assert sorted(i for b in plan for i in b) == list(range(8)) and len(plan) <= 6
  assert all(b == sorted(b) for b in plan)
  # as many dtypes as splits -> one bucket per dtype
  plan = plan_buckets([4, 4, 2, 2], [torch.float32, torch.float32, torch.float16, torch.float16], 2)
  assert plan == [[0, 1], [2, 3]]
  assert plan_buckets([8], [torch.float32], 5) == [[0]]
  assert plan_buckets([], [], 5) == []
  a = torch.zeros(8, 1024, 1024, dtype=torch.int32)
  b = torch.zeros(7, 1024, 1024, dtype=torch.int32)
  assert estimate_split_num_for_comm([a]) == 1
  assert estimate_split_num_for_comm([a, b]) == 2
  assert estimate_split_num_for_comm([a, b, torch.zeros(1, dtype=torch.float32)]) == 3


@pytest.mark.parametrize("S,M", [(2, 4), (4, 6), (4, 8), (3, 2), (1, 4), (8, 16)])
def test_schedules_in_flight_and_deadlock_free(S, M):
  for policy in ("PreferForward", "PreferBackward", "PreferBackwardOptimizer"):
    progs = schedule.build_programs(policy, S, M)
    res = schedule.simulate(progs)
    assert res.ok, (policy, res.reason)
    for s in range(S):
      cap = schedule.in_flight_cap(policy, s, S, M)
      assert res.max_in_flight[s] == cap, (policy, s, res.max_in_flight, cap)
      fs = [i.mb for i in progs[s] if i.op == schedule.F]
      bs = [i.mb for i in progs[s] if i.op == schedule.B]
      assert fs == list(range(M)) and bs == list(range(M))
  gpipe = schedule.simulate(schedule.build_programs("PreferForward", S, M)).makespan
  onef = schedule.simulate(schedule.build_programs("PreferBackward", S, M)).makespan
  assert onef <= gpipe + 1e-9


def test_1f1b_reference_control_edges():
  """The edges the reference asserts (scheduler_test.py): 4 stages, 6 micro-batches, PreferBackward:
  on stage s, F(m) runs after B(m - (S - s))."""
  S, M = 4, 6
  progs = schedule.build_programs("PreferBackward", S, M)
  for s in range(S):
    order = [(i.op, i.mb) for i in progs[s] if i.op in (schedule.F, schedule.B)]
    pos = {k: n for n, k in enumerate(order)}
    for m in range(M):
      dep = m - (S - s)
      if dep >= 0:
        assert pos[(schedule.B, dep)] < pos[(schedule.F, m)]
      if m >= 1 and dep < 0:
        assert pos[(schedule.F, m - 1)] < pos[(schedule.F, m)]
  with pytest.raises(RuntimeError):
    schedule.get_scheduler("nope")
  assert len(schedule.get_scheduler("PreferBackwardOptimizer")(2, 4)) == 2


def test_receives_are_prefetched():
  progs = schedule.build_programs("PreferBackward", 2, 4, prefetch=1)
  last = progs[1]
  first_f = next(n for n, i in enumerate(last) if i.op == schedule.F)
  # RECV_F(0) and RECV_F(1) are both posted before F(0) runs
  posted = [(i.op, i.mb) for i in last[:first_f]]
  assert (schedule.RECV_F, 0) in posted and (schedule.RECV_F, 1) in posted


@pytest.mark.parametrize("family", ["gpt2", "bert"])
def test_scoped_models_split_their_blocks_evenly(family):
  """Blocks built between two ``set_default_strategy`` calls belong to the stage that was open when they were BUILT, not to
  the one open when the ``ModuleList`` holding them was attached (that put every block on the last stage)."""
  import easyparallellibrary_b200 as epl
  epl.init(epl.Config({"pipeline.num_micro_batch": 4}))
  if family == "gpt2":
    from easyparallellibrary_b200.models.gpt2 import GPT2, GPT2Config, lm_loss
    model = GPT2(GPT2Config.named("small", num_pipeline_stages=2, tie_embeddings=False))
    tr = epl.Trainer(model, "adamw", lr=1e-4, loss_fn=lm_loss).build()
  else:
    from easyparallellibrary_b200.models.bert import Bert, BertConfig, squad_loss
    model = Bert(BertConfig.named("base", num_pipeline_stages=2))
    tr = epl.Trainer(model, "adamw", lr=1e-4, loss_fn=squad_loss).build()
  layers = model.epl_sequential()
  assign = tr._assign_layers(layers, epl.Graph.get())
  n_blocks = len(layers) - 2
  # GPT-2 weighs the vocabulary projection (5.45 blocks' worth for "small") onto the last stage: 9 + 3 blocks; BERT: 6 + 6
  first = round((n_blocks + 50257 / (12.0 * 768)) / 2) if family == "gpt2" else n_blocks // 2
  assert assign.count(0) == 1 + first and assign.count(1) == 1 + n_blocks - first, assign
  epl.shutdown()


def test_stage_function_reads_the_loss_scale_from_a_tensor():
  """``parallel/pipeline.py::_StageFn`` (what a captured stage graph runs): loss scale x 1/M is a device scalar, so a dynamic loss
  scale changes the replayed program's result without re-capturing."""
  import torch
  from torch import nn
  from easyparallellibrary_b200.parallel.pipeline import _StageFn
  torch.manual_seed(0)
  lin = nn.Linear(4, 1)
  fn = _StageFn(lin, lambda y, t: ((y - t) ** 2).mean(), 8.0, device="cpu")
  x, t = torch.randn(3, 4), torch.randn(3, 1)
  plain = ((lin(x) - t) ** 2).mean()
  y = fn(x, t)
  assert torch.allclose(y, plain * 8.0)
  (g8,) = torch.autograd.grad(y, lin.weight)
  fn.set_factor(2.0)
  assert fn.factor == 2.0 and float(fn.factor_t) == 2.0
  y2 = fn(x, t)
  (g2,) = torch.autograd.grad(y2, lin.weight)
  assert torch.allclose(y2, plain * 2.0) and torch.allclose(g8, g2 * 4.0)
  assert list(p for p in fn.parameters()) == list(lin.parameters())           # the scalar is not a parameter of the stage
  mid = _StageFn(lin, None, 1.0)                                               # a stage without the loss passes its output through
  assert mid.factor_t is None and torch.equal(mid(x), lin(x))
