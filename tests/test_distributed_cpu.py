"""Multi-process CPU (gloo) tests: communicator verbs, DP/ZeRO/GA parity, pipeline parity, collections.
The reference has no CPU communication backend and only two true multi-process tests;
this tier is what BASELINE config #1 asks for."""
import numpy as np
import pytest
import torch
from torch import nn

from dist_utils import run_distributed


# ------------------------------------------------------------------------------------------- communicator
def _comm_worker(rank, world):
  import easyparallellibrary_b200 as epl
  from easyparallellibrary_b200.communicators import CollectiveCommunicator
  from easyparallellibrary_b200.communicators import functional as F
  from easyparallellibrary_b200.communicators.sparse import sparse_all_reduce
  epl.init()
  comm = CollectiveCommunicator("test", list(range(world)), max_splits=3, num_communicators=2)
  out = {}
  ts = [torch.full((n,), float(rank + 1)) for n in (3, 5, 7, 11)] + [torch.full((4,), rank + 1, dtype=torch.int64)]
  red = comm.batch_allreduce(ts, mean=False)
  out["allreduce"] = [t.tolist() for t in red]
  out["mean"] = comm.batch_allreduce([torch.full((2,), float(rank))], mean=True)[0].tolist()
  comm16 = CollectiveCommunicator("test16", list(range(world)), enable_fp16=True, fp16_scale=128)
  out["fp16"] = comm16.batch_allreduce([torch.full((4,), 0.5 * (rank + 1))], mean=False)[0].tolist()
  b = torch.full((6,), float(rank))
  comm.broadcast(b, root=1)
  out["broadcast"] = b.tolist()
  out["allgather"] = comm.allgather(torch.full((2, 2), float(rank))).tolist()
  out["reduce"] = comm.reduce(torch.ones(3) * (rank + 1), root=0).tolist()
  out["reduce_scatter"] = comm.reduce_scatter(torch.arange(2 * world, dtype=torch.float32) * (rank + 1)).tolist()
  out["alltoall"] = comm.alltoall(torch.arange(world * 2, dtype=torch.float32) + 100 * rank).tolist()
  g, counts = comm.allgatherv(torch.full((rank + 1, 2), float(rank)))
  out["allgatherv"] = (g.tolist(), counts.tolist())
  rows = torch.arange(3, dtype=torch.float32).unsqueeze(1) + 10 * rank          # 3 rows: send 1 to rank0, 2 to rank1
  recv, rc = comm.alltoallv(rows, torch.tensor([1, 2]))
  out["alltoallv"] = (recv.flatten().tolist(), rc.tolist())
  # autograd adjoints
  x = torch.ones(2, 3, requires_grad=True)
  F.all_gather(x * (rank + 1), comm).sum().backward()
  out["ag_grad"] = x.grad.tolist()
  y = torch.ones(2 * world, requires_grad=True)
  (F.reduce_scatter(y, comm) * (rank + 1)).sum().backward()
  out["rs_grad"] = y.grad.tolist()
  z = torch.ones(4, requires_grad=True)
  F.all_reduce(z * 2, comm).sum().backward()
  out["ar_grad"] = z.grad.tolist()
  w = torch.arange(2.0 * world, requires_grad=True)
  (F.all_to_all(w, comm) * (rank + 1)).sum().backward()
  out["a2a_grad"] = w.grad.tolist()
  sp = torch.sparse_coo_tensor(torch.tensor([[rank, 3]]), torch.ones(2, 2) * (rank + 1), (5, 2))
  out["sparse"] = sparse_all_reduce(comm, sp).to_dense().tolist()
  return out


def test_communicator_verbs_two_ranks():
  r0, r1 = run_distributed(_comm_worker, 2)
  assert r0["allreduce"] == [[3.0] * 3, [3.0] * 5, [3.0] * 7, [3.0] * 11, [3] * 4] == r1["allreduce"]
  assert r0["mean"] == [0.5, 0.5]
  assert r0["fp16"] == [1.5] * 4
  assert r0["broadcast"] == [1.0] * 6 == r1["broadcast"]
  assert r0["allgather"] == [[0.0, 0.0], [0.0, 0.0], [1.0, 1.0], [1.0, 1.0]]
  assert r0["reduce"] == [3.0, 3.0, 3.0]
  assert r0["reduce_scatter"] == [0.0, 3.0] and r1["reduce_scatter"] == [6.0, 9.0]
  assert r0["alltoall"] == [0.0, 1.0, 100.0, 101.0] and r1["alltoall"] == [2.0, 3.0, 102.0, 103.0]
  assert r0["allgatherv"][0] == [[0.0, 0.0], [1.0, 1.0], [1.0, 1.0]] and r0["allgatherv"][1] == [1, 2]
  assert r0["alltoallv"] == ([0.0, 10.0], [1, 1]) and r1["alltoallv"] == ([1.0, 2.0, 11.0, 12.0], [2, 2])
  assert r0["ag_grad"] == [[2.0] * 3] * 2 and r1["ag_grad"] == [[4.0] * 3] * 2       # adjoint = reduce-scatter
  assert r0["rs_grad"] == [1.0, 1.0, 2.0, 2.0]                                        # adjoint = all-gather
  assert r0["ar_grad"] == [4.0] * 4
  assert r0["a2a_grad"] == [1.0, 1.0, 2.0, 2.0]
  assert r0["sparse"] == [[1.0, 1.0], [2.0, 2.0], [0.0, 0.0], [3.0, 3.0], [0.0, 0.0]]


# ------------------------------------------------------------------------------------------- DP parity
def _mlp():
  torch.manual_seed(0)
  return nn.Sequential(nn.Linear(10, 16), nn.ReLU(), nn.Linear(16, 1))


def _train(rank, world, conf, steps=4, clip=None, opt="adamw"):
  import easyparallellibrary_b200 as epl
  epl.init(epl.Config(conf))
  with epl.replicate(device_count=1):
    model = _mlp()
  tr = epl.Trainer(model, opt, loss_fn=lambda o, y: ((o - y) ** 2).mean(), lr=1e-2, max_grad_norm=clip)
  torch.manual_seed(1)
  X, Y = torch.randn(steps, 8, 10), torch.randn(steps, 8, 1)
  losses = []
  for i in range(steps):
    x, y = X[i], Y[i]
    if tr.build().plan.num_replicas > 1:
      n = tr.plan.num_replicas
      r = tr.plan.placements[tr.plan.stage_taskgraphs[0]].replica
      x, y = x.chunk(n)[r], y.chunk(n)[r]
    out = tr.step(x, y)
    epl.add_to_collection  # noqa: B018
    losses.append(out.item())
  return losses, [p.detach().float().numpy().copy() for p in model.parameters()]


def _max_diff(a, b):
  return max(float(np.abs(x - y).max()) for x, y in zip(a, b))


@pytest.mark.parametrize("conf", [
    {}, {"zero.level": "v0"}, {"zero.level": "v1"}, {"communication.fp16": True, "communication.fp16_scale": 64},
    {"pipeline.num_micro_batch": 2}, {"optimizer.num_apply_group": 3}, {"communication.max_splits": 1},
])
def test_data_parallel_matches_single_process(conf):
  base = run_distributed(_train, 1, args=({k: v for k, v in conf.items() if k.startswith(("pipeline", "optimizer"))},))[0]
  dist_res = run_distributed(_train, 2, args=(conf,))
  tol = 2e-3 if conf.get("communication.fp16") else 1e-6
  for res in dist_res:
    assert _max_diff(base[1], res[1]) < tol
  assert _max_diff(dist_res[0][1], dist_res[1][1]) == 0.0       # replicas stay bit-identical


@pytest.mark.parametrize("conf", [{}, {"zero.level": "v1"}])
def test_wrapped_torch_optimizer_under_data_parallel(conf):
  """A torch optimizer class (here Adagrad) on the fp32 master shards: two ranks (optionally ZeRO-1 shards) == one process."""
  base = run_distributed(_train, 1, args=({}, 4, None, torch.optim.Adagrad))[0]
  two = run_distributed(_train, 2, args=(conf, 4, None, torch.optim.Adagrad))
  for res in two:
    assert _max_diff(base[1], res[1]) < 1e-6
  assert _max_diff(two[0][1], two[1][1]) == 0.0


def test_sgd_and_clipping_modes():
  for conf in ({}, {"communication.clip_after_allreduce": True}):
    base = run_distributed(_train, 1, args=({}, 3, 0.5, "sgd"))[0]
    two = run_distributed(_train, 2, args=(conf, 3, 0.5, "sgd"))
    if conf:      # reduce-then-clip == single process on the full batch
      assert _max_diff(base[1], two[0][1]) < 1e-6
    assert _max_diff(two[0][1], two[1][1]) < 1e-7


def test_gradient_accumulation_equals_big_batch():
  a = run_distributed(_train, 1, args=({},))[0]
  b = run_distributed(_train, 1, args=({"pipeline.num_micro_batch": 4},))[0]
  assert _max_diff(a[1], b[1]) < 1e-6


# ------------------------------------------------------------------------------------------- pipeline parity
_LAYER_FACTORIES = [lambda: nn.Sequential(nn.Linear(10, 32), nn.Tanh()), lambda: nn.Sequential(nn.Linear(32, 32), nn.Tanh()),
                    lambda: nn.Sequential(nn.Linear(32, 32), nn.Tanh()), lambda: nn.Linear(32, 1)]


def _train_pipe(rank, world, conf, stages, steps=3):
  import easyparallellibrary_b200 as epl
  epl.init(epl.Config(conf))
  torch.manual_seed(0)
  per = len(_LAYER_FACTORIES) // stages
  mods = []
  for s in range(stages):
    with epl.replicate(device_count=1, name="stage%d" % s):
      mods.append(nn.Sequential(*[f() for f in _LAYER_FACTORIES[s * per:(s + 1) * per]]))
  model = nn.Sequential(*mods)
  tr = epl.Trainer(model, "adamw", loss_fn=lambda o, y: ((o - y) ** 2).mean(), lr=1e-2).build()
  torch.manual_seed(1)
  X, Y = torch.randn(steps, 16, 10), torch.randn(steps, 16, 1)
  n = tr.plan.num_replicas
  rep = next(iter(tr.plan.placements.values())).replica
  losses = []
  for i in range(steps):
    x, y = X[i].chunk(n)[rep], Y[i].chunk(n)[rep]
    losses.append(tr.step(x, y).item())
  params = {}
  for s in tr.plan.local_stages:
    for k, v in tr.stage_modules[s].state_dict().items():
      params["%d.%s" % (s, k)] = v.float().numpy().copy()
  return losses, params, tr.plan.pipeline


@pytest.mark.parametrize("policy", ["PreferBackward", "PreferForward", "PreferBackwardOptimizer"])
def test_two_stage_pipeline_matches_single_process(policy):
  conf = {"pipeline.num_micro_batch": 4, "pipeline.strategy": policy}
  base = run_distributed(_train_pipe, 1, args=(conf, 2))[0]
  res = run_distributed(_train_pipe, 2, args=(conf, 2))
  assert res[0][2] and res[1][2] and not base[2]
  merged = {}
  for r in res:
    merged.update(r[1])
  assert set(merged) == set(base[1])
  assert max(float(np.abs(merged[k] - base[1][k]).max()) for k in merged) < 1e-6
  for a, b in zip(res[0][0], base[0]):
    assert abs(a - b) < 1e-5          # every stage reports the replica's loss


def test_pipeline_times_data_parallel_four_ranks():
  conf = {"pipeline.num_micro_batch": 2}
  base = run_distributed(_train_pipe, 1, args=(conf, 2))[0]
  res = run_distributed(_train_pipe, 4, args=(conf, 2))
  merged = {}
  for r in res:
    merged.update(r[1])
  assert max(float(np.abs(merged[k] - base[1][k]).max()) for k in merged) < 1e-6


# ------------------------------------------------------------------------------------------- collections
def _collections_worker(rank, world):
  import easyparallellibrary_b200 as epl
  epl.init(epl.Config({"pipeline.num_micro_batch": 2}))

  class Net(nn.Module):
    def __init__(self):
      super().__init__()
      self.fc = nn.Linear(4, 1)

    def forward(self, x, y):
      out = self.fc(x)
      loss = ((out - y) ** 2).mean()
      epl.add_to_collection(loss, epl.GraphKeys.GLOBAL_MEAN_OBJECTS)
      epl.add_to_collection(out.detach().flatten(), epl.GraphKeys.GLOBAL_CONCAT_OBJECTS)
      epl.add_to_collection(torch.tensor(float(x.shape[0])), epl.GraphKeys.GLOBAL_SUM_OBJECTS)
      epl.add_to_collection(torch.tensor(float(rank)), epl.GraphKeys.LOCAL_MEAN_OBJECTS)
      return loss

  with epl.replicate(1):
    net = Net()
  tr = epl.Trainer(net, "sgd", lr=0.0)
  x, y = torch.ones(4, 4) * (rank + 1), torch.zeros(4, 1)
  out = tr.step(x, y)
  c = out.collections
  return {k: [v.tolist() if v.dim() else float(v) for v in vals] for k, vals in c.items()}


def test_collections_merge_over_micro_batches_and_replicas():
  r0, r1 = run_distributed(_collections_worker, 2)
  assert r0[r"global_sum_objects"] == [8.0]                    # 2 micro-batches x 2 rows x 2 replicas
  assert len(r0["global_concat_objects"][0]) == 8
  assert r0["global_mean_objects"] == r1["global_mean_objects"]
  assert r0["local_mean_objects"] == [0.0] and r1["local_mean_objects"] == [1.0]


def test_zero3_matches_single_process():
  base = run_distributed(_train, 1, args=({},))[0]
  for conf in ({"zero.level": "v3"}, {"zero.level": "v3", "gradient_checkpoint.type": "auto", "offload.level": "v0"}):
    res = run_distributed(_train, 2, args=(conf,))
    for r in res:
      assert abs(r[0][-1] - res[0][0][-1]) < 10          # losses are per-replica; parameters below are the real check
    one = run_distributed(_train, 1, args=(conf,))[0]
    assert max(abs(a - b) for a, b in zip(one[0], base[0])) < 1e-5


def _zero3_params_worker(rank, world):
  import easyparallellibrary_b200 as epl
  epl.init(epl.Config({"zero.level": "v3"}))
  with epl.replicate(1):
    model = _mlp()
  tr = epl.Trainer(model, "adamw", loss_fn=lambda o, y: ((o - y) ** 2).mean(), lr=1e-2).build()
  torch.manual_seed(1)
  X, Y = torch.randn(4, 8, 10), torch.randn(4, 8, 1)
  for i in range(4):
    tr.step(X[i].chunk(world)[rank], Y[i].chunk(world)[rank])
  z = tr.zero3[0]
  resident = sum(p.numel() for p in model.parameters())
  z.gather_all()
  out = [p.detach().float().numpy().copy() for p in model.parameters()]
  z.release_all()
  return out, resident


def test_zero3_two_ranks_parameters_match_and_are_released():
  base = run_distributed(_train, 1, args=({},))[0]
  res = run_distributed(_zero3_params_worker, 2)
  for out, resident in res:
    assert resident == 0                                  # nothing materialised between steps
    assert _max_diff(base[1], out) < 1e-6


def _handle_exchange_worker(rank, world):
  import easyparallellibrary_b200 as epl
  from easyparallellibrary_b200.runtime.symmetric import _exchange_handles
  epl.init()                                       # joins the gloo process group (and with it the rendezvous store)
  ranks = [0, 1] if rank < 2 else [2, 3]
  outs = []
  for k in range(3 if rank < 2 else 1):          # subset [0,1] creates three buffers, subset [2,3] only one
    raw = bytes([rank, k]) * 32
    outs.append(_exchange_handles(raw, ranks, ranks.index(rank)))
  return [[list(b[:2]) for b in o] for o in outs]


def test_symmetric_handle_exchange_over_rank_subsets():
  """The IPC-handle exchange of NVLink symmetric buffers must work for strict subsets of the world that create different
  numbers of buffers (the data-parallel group of each pipeline stage, one of several tensor-parallel groups): it goes
  through the rendezvous store, not a WORLD collective."""
  res = run_distributed(_handle_exchange_worker, 4)
  assert res[0] == [[[0, k], [1, k]] for k in range(3)] and res[1] == res[0]
  assert res[2] == [[[2, 0], [3, 0]]] and res[3] == res[2]


def _zero3_fused_gather_worker(rank, world, fused):
  """ZeRO-3 on a stack of layers built from ops.linear.Linear; with ``fused`` the first GEMM weight of every layer defers its
  all-gather into the GEMM (K2 protocol), emulated here by all-gather + matmul behind the implementation hook."""
  import easyparallellibrary_b200 as epl
  from easyparallellibrary_b200.ops.linear import Linear
  from easyparallellibrary_b200.parallel import zero3
  calls = [0]

  def emulate(x2, w_shard, group, bias=None, gelu=False, out_w_full=None):
    calls[0] += 1
    full = group.comm.allgather(w_shard.contiguous())             # [N, K] rows in rank order
    out_w_full.view_as(full).copy_(full)
    pre = torch.nn.functional.linear(x2, full, bias)
    return (torch.nn.functional.gelu(pre, approximate="tanh"), pre, full) if gelu else (pre, None, full)

  zero3.GATHER_GEMM_IMPL = emulate if fused else None

  class Layer(nn.Module):
    def __init__(self, d, h, gelu):
      super().__init__()
      self.norm = nn.LayerNorm(d)
      self.up = Linear(d, h, gelu=gelu)                    # first GEMM: deferred gather (h % world == 0)
      self.down = Linear(h, d)

    def forward(self, x):
      return x + self.down(self.up(self.norm(x)))

  epl.init(epl.Config({"zero.level": "v3"}))
  torch.manual_seed(0)
  with epl.replicate(1):
    model = nn.Sequential(Layer(16, 32, True), Layer(16, 24, False), nn.Linear(16, 1))
  tr = epl.Trainer(model, "adamw", loss_fn=lambda o, y: ((o - y) ** 2).mean(), lr=1e-2).build()
  z = tr.zero3[0]
  deferred = sum(1 for u in z.units if u.deferred)
  torch.manual_seed(1)
  X, Y = torch.randn(4, 8, 16), torch.randn(4, 8, 1)
  losses = [float(tr.step(X[i].chunk(world)[rank], Y[i].chunk(world)[rank]).loss) for i in range(4)]
  z.gather_all()
  out = [p.detach().float().numpy().copy() for p in model.parameters()]
  z.release_all()
  zero3.GATHER_GEMM_IMPL = None
  return out, losses, calls[0], deferred


def test_zero3_deferred_weight_gather_matches_plain_zero3():
  """K2 integration: the weight of a layer's first GEMM is gathered inside that GEMM (parallel/zero3.py::PendingGather ->
  ops.linear._GatherLinearFn).  Same parameters and losses as ZeRO-3 with ordinary all-gathers, and the fused path really ran:
  2 deferred units x 4 steps forward (the backward re-gathers through the library path)."""
  plain = run_distributed(_zero3_fused_gather_worker, 2, args=(False,))
  fused = run_distributed(_zero3_fused_gather_worker, 2, args=(True,))
  assert plain[0][3] == 0 and fused[0][3] == 2 and plain[0][2] == 0 and fused[0][2] == 8
  for r in range(2):
    np.testing.assert_allclose(fused[r][1], plain[r][1], rtol=1e-5, atol=1e-6)
    for a, b in zip(fused[r][0], plain[r][0]):
      np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-6)


def _ckpt_pipeline_worker(rank, world, directory):
  import easyparallellibrary_b200 as epl
  import torch.distributed as dist
  from easyparallellibrary_b200.models.gpt2 import GPT2, GPT2Config, lm_loss
  from easyparallellibrary_b200.runtime.saver import load_checkpoint, save_checkpoint

  def build():
    epl.init(epl.Config({"pipeline.num_micro_batch": 2}))
    cfg = GPT2Config.named("tiny", num_pipeline_stages=2, tie_embeddings=False)
    torch.manual_seed(0)
    return epl.Trainer(GPT2(cfg), "adamw", lr=1e-3, loss_fn=lm_loss), cfg

  tr, cfg = build()
  g = torch.Generator().manual_seed(rank // 2)                 # ranks (0,1) and (2,3) are the two pipeline replicas
  toks = [torch.randint(0, cfg.vocab_size, (4, 16), generator=g) for _ in range(4)]
  for t in toks[:2]:
    tr.step(t, t)
  save_checkpoint(tr, directory)
  after = [float(tr.step(t, t).loss) for t in toks[2:]]
  tr2, _ = build()
  step = load_checkpoint(tr2, directory)
  resumed = [float(tr2.step(t, t).loss) for t in toks[2:]]
  dist.barrier()
  return step, after, resumed


def test_checkpoint_resume_under_pipeline_times_data_parallel(tmp_path):
  """Save after 2 steps, resume in a fresh Trainer, continue: identical losses (reference saver_test.py:123-297 asserts
  global-step continuity and tensor equality; here under 2 stages x 2 replicas, first replica of each stage writes)."""
  res = run_distributed(_ckpt_pipeline_worker, 4, args=(str(tmp_path / "ckpt"),), timeout=300)
  for step, after, resumed in res:
    assert step == 2
    np.testing.assert_allclose(resumed, after, rtol=0, atol=1e-6)


def _ckpt_zero_worker(rank, world, directory, zero):
  import easyparallellibrary_b200 as epl
  from easyparallellibrary_b200.models.gpt2 import GPT2, GPT2Config
  from easyparallellibrary_b200.runtime.saver import load_checkpoint, save_checkpoint

  def build():
    epl.init(epl.Config({"zero.level": zero}))
    cfg = GPT2Config.named("tiny")
    torch.manual_seed(0)
    with epl.replicate(1):
      model = GPT2(cfg)
    return epl.Trainer(model, "adamw", lr=1e-3), cfg

  tr, cfg = build()
  g = torch.Generator().manual_seed(rank)
  toks = [torch.randint(0, cfg.vocab_size, (2, 16), generator=g) for _ in range(4)]
  for t in toks[:2]:
    tr.step(t, t)
  save_checkpoint(tr, directory)
  after = [float(tr.step(t, t).loss) for t in toks[2:]]
  tr2, _ = build()
  step = load_checkpoint(tr2, directory)
  resumed = [float(tr2.step(t, t).loss) for t in toks[2:]]
  return step, after, resumed


@pytest.mark.parametrize("zero", ["v1", "v3"])
def test_checkpoint_resume_with_sharded_optimizer_state(tmp_path, zero):
  """With ZeRO every rank owns a shard of the optimizer state (v3: of the parameters too): each rank writes and restores its
  own shard, and training resumes bit-identically (the reference drops optimizer slots under ZeRO, hooks.py:340-344)."""
  res = run_distributed(_ckpt_zero_worker, 2, args=(str(tmp_path / "ckpt"), zero), timeout=300)
  for step, after, resumed in res:
    assert step == 2
    np.testing.assert_allclose(resumed, after, rtol=0, atol=1e-6)


def _amp_pipeline_worker(rank, world):
  import easyparallellibrary_b200 as epl
  epl.init(epl.Config({"pipeline.num_micro_batch": 2, "amp.level": "O1", "amp.loss_scale": "dynamic"}))
  torch.manual_seed(0)
  with epl.replicate(1, name="s0"):
    a = nn.Sequential(nn.Linear(10, 16), nn.ReLU())
  with epl.replicate(1, name="s1"):
    b = nn.Linear(16, 2)
  tr = epl.Trainer(nn.Sequential(a, b), "sgd", lr=1000.0, loss_fn=lambda y, t: nn.functional.cross_entropy(y.float(), t))
  g = torch.Generator().manual_seed(0)
  skipped = []
  for _ in range(6):
    x, y = torch.randn(8, 10, generator=g) * 50, torch.randint(0, 2, (8,), generator=g)
    skipped.append(bool(tr.step(x, y).skipped))
  ev = tr.eval_step(torch.randn(4, 10), torch.randint(0, 2, (4,)))            # fp32 input is cast like in step()
  return skipped, float(tr.scaler.loss_scale), ev is not None


def test_amp_overflow_is_skipped_on_every_pipeline_stage():
  """2-stage pipeline + fp16 dynamic loss scaling with a learning rate that overflows (reference amp_parallel_test.py /
  test_amp_parallel.sh): an overflow seen by one stage skips the update on ALL stages and the loss scales stay equal."""
  res = run_distributed(_amp_pipeline_worker, 2)
  assert res[0][0] == res[1][0] and any(res[0][0]) and res[0][1] == res[1][1]
  assert res[1][2] and not res[0][2]                                             # the loss lives on the last stage


@pytest.mark.parametrize("policy", ["PreferBackward", "PreferForward", "PreferBackwardOptimizer"])
def test_four_stage_pipeline_matches_single_process(policy):
  """Four stages x eight micro-batches: deeper than the 2-stage case, so middle stages both receive and send in each direction
  and the receive hoisting of every schedule is exercised."""
  conf = {"pipeline.num_micro_batch": 8, "pipeline.strategy": policy}
  base = run_distributed(_train_pipe, 1, args=(conf, 4))[0]
  res = run_distributed(_train_pipe, 4, args=(conf, 4), timeout=300)
  assert all(r[2] for r in res)
  merged = {}
  for r in res:
    merged.update(r[1])
  assert set(merged) == set(base[1])
  assert max(float(np.abs(merged[k] - base[1][k]).max()) for k in merged) < 1e-6
  for r in res:
    for a, b in zip(r[0], base[0]):
      assert abs(a - b) < 1e-5


# ------------------------------------------------------------------------------------------- train / evaluate loops
def _loop_pipe_worker(rank, world, stages):
  """``epl.train_and_evaluate`` under a pipeline: every stage takes part in every evaluation forward, only the last stage gets a
  result (loss with labels, model output without), and the evaluation barrier keeps the ranks together."""
  import easyparallellibrary_b200 as epl
  epl.init(epl.Config({"pipeline.num_micro_batch": 2}))
  torch.manual_seed(0)
  per = len(_LAYER_FACTORIES) // stages
  mods = []
  for s in range(stages):
    with epl.replicate(device_count=1, name="stage%d" % s):
      mods.append(nn.Sequential(*[f() for f in _LAYER_FACTORIES[s * per:(s + 1) * per]]))
  tr = epl.Trainer(nn.Sequential(*mods), "adamw", loss_fn=lambda o, y: ((o - y) ** 2).mean(), lr=1e-2).build()
  g = torch.Generator().manual_seed(7)
  data = [(torch.randn(8, 10, generator=g), torch.randn(8, 1, generator=g)) for _ in range(3)]
  hist = epl.train_and_evaluate(tr, data, data[:2], max_steps=4, eval_every=2)
  outs = epl.evaluate(tr, [x for x, _ in data[:2]], metric_fn=lambda out, batch: {"rows": out.shape[0]})
  return [(h["global_step"], h["batches"], h.get("loss")) for h in hist], outs, tr.plan.pipeline


def test_train_and_evaluate_under_a_pipeline():
  base = run_distributed(_loop_pipe_worker, 1, args=(2,))[0]
  res = run_distributed(_loop_pipe_worker, 2, args=(2,))
  assert res[0][2] and res[1][2] and not base[2]
  # rank 0 holds stage 0: it runs every evaluation forward but sees no result of its own; the merged metrics of the job (here: of
  # the last stage) are what one process reports, and every rank returns them
  for r in (0, 1):
    assert [h[:2] for h in res[r][0]] == [(2, 2), (4, 2)] and res[r][1] == {"rows": 8.0, "batches": 2}
    for a, b in zip(res[r][0], base[0]):
      assert abs(a[2] - b[2]) < 1e-5


# ------------------------------------------------------------------------------------------- elastic resume
def _elastic_worker(rank, world, directory, mode):
  """mode "save": train 2 steps on the global batch split over ``world`` ranks, checkpoint, then 2 more steps (the reference
  trajectory).  mode "load": restore in a job of another size and run the same 2 steps."""
  import easyparallellibrary_b200 as epl
  from easyparallellibrary_b200.models.gpt2 import GPT2, GPT2Config
  from easyparallellibrary_b200.runtime.saver import load_checkpoint, save_checkpoint
  epl.init(epl.Config({"zero.level": "v1"}))
  cfg = GPT2Config.named("tiny")
  torch.manual_seed(0)
  with epl.replicate(1):
    model = GPT2(cfg)
  tr = epl.Trainer(model, "adamw", lr=1e-3)
  g = torch.Generator().manual_seed(5)
  toks = [torch.randint(0, cfg.vocab_size, (4, 16), generator=g) for _ in range(4)]        # the GLOBAL batch of each step
  mine = lambda t: t.chunk(world)[rank]                                                     # noqa: E731
  if mode == "save":
    for t in toks[:2]:
      tr.step(mine(t), mine(t))
    save_checkpoint(tr, directory)
  else:
    assert load_checkpoint(tr, directory) == 2
  losses = [float(tr.step(mine(t), mine(t)).loss) for t in toks[2:]]
  params = torch.cat([p.detach().float().flatten() for p in model.parameters()]).numpy().copy()
  moments = float(sum(o.m.abs().sum() for o in tr.optimizers[0]))
  return losses, params, moments


@pytest.mark.parametrize("old,new", [(2, 1), (1, 2)])
def test_elastic_resume_reshards_the_optimizer_state(tmp_path, old, new):
  """A checkpoint written by ``old`` data-parallel ranks (ZeRO-1 shards, or one unsharded copy) resumes in a job of ``new`` ranks:
  the Adam moments are re-sliced, so the trajectory continues as if the job size had never changed."""
  d = str(tmp_path / "ckpt")
  ref = run_distributed(_elastic_worker, old, args=(d, "save"), timeout=300)
  got = run_distributed(_elastic_worker, new, args=(d, "load"), timeout=300)
  assert got[0][2] > 0                                              # moments were restored, not restarted at zero
  assert float(np.abs(got[0][1] - ref[0][1]).max()) < 2e-5          # same weights after two more steps (restarted moments: ~1e-3)
  if new == 1:                                                      # (per-rank losses differ when the batch is split differently)
    assert abs(sum(l[0] for l, _, _ in ref) / old - got[0][0][0]) < 1e-5
