"""epl.split tensor-parallel ops on 2 CPU ranks (gloo) against the unsharded computation.
Mirrors the intent of the reference's split_test.py / resnet_split example (replicate backbone + split head)."""
import numpy as np
import torch
from torch import nn

from dist_utils import run_distributed


def _split_head_worker(rank, world, steps=3):
  import easyparallellibrary_b200 as epl
  from easyparallellibrary_b200.ops import tensor_parallel as tp
  epl.init(epl.Config({"cluster.colocate_split_and_replicate": True}))
  torch.manual_seed(0)
  C, H = 11, 8                       # 11 classes over 2 shards -> 6 + 5 (remainder to shard 0)
  full_w = torch.randn(C, H) * 0.3
  full_b = torch.randn(C) * 0.1
  with epl.replicate(device_count=world):
    backbone = nn.Linear(6, H)
  with epl.split(device_count=world):
    head = tp.DistributedDense(H, C)
    with torch.no_grad():
      head.weight.copy_(full_w[head.start:head.end])
      head.bias.copy_(full_b[head.start:head.end])

  class Net(nn.Module):
    def __init__(self):
      super().__init__()
      self.backbone, self.head = backbone, head

    def forward(self, x, y):
      with epl.split(device_count=world):
        logits = self.head(torch.tanh(self.backbone(x)))
        loss = tp.distributed_sparse_softmax_cross_entropy_with_logits(y, logits)
        pred = tp.distributed_argmax(logits)
        acc = tp.distributed_equal(pred, y).float().mean()
      epl.add_to_collection(acc, epl.GraphKeys.LOCAL_MEAN_OBJECTS)
      return loss

  net = Net()
  tr = epl.Trainer(net, "sgd", lr=0.1).build()
  assert tr.has_split and len(tr.group_keys) == 2
  torch.manual_seed(1)
  X, Y = torch.randn(steps, 8, 6), torch.randint(0, C, (steps, 8))
  losses, accs = [], []
  for i in range(steps):
    x, y = X[i].chunk(world)[rank], Y[i].chunk(world)[rank]
    out = tr.step(x, y)
    losses.append(out.item())
    accs.append(float(out.collections[epl.GraphKeys.LOCAL_MEAN_OBJECTS][0]))
  return (losses, accs, backbone.weight.detach().numpy().copy(), head.weight.detach().numpy().copy(), (head.start, head.end),
          full_w.numpy(), full_b.numpy())


def test_replicate_backbone_split_head_matches_unsharded():
  res = run_distributed(_split_head_worker, 2)
  full_w, full_b = res[0][5], res[0][6]
  # unsharded reference in this process
  torch.manual_seed(0)
  _ = torch.randn(11, 8), torch.randn(11)
  backbone = nn.Linear(6, 8)
  W = nn.Parameter(torch.tensor(full_w))
  Bb = nn.Parameter(torch.tensor(full_b))
  opt = torch.optim.SGD(list(backbone.parameters()) + [W, Bb], lr=0.1)
  torch.manual_seed(1)
  X, Y = torch.randn(3, 8, 6), torch.randint(0, 11, (3, 8))
  ref_losses = []
  for i in range(3):
    logits = torch.nn.functional.linear(torch.tanh(backbone(X[i])), W, Bb)
    loss = torch.nn.functional.cross_entropy(logits, Y[i])
    opt.zero_grad()
    loss.backward()
    opt.step()
    ref_losses.append(loss.item())
  for r in res:
    assert np.allclose(r[0], ref_losses, atol=1e-5), (r[0], ref_losses)
    assert np.allclose(r[2], backbone.weight.detach().numpy(), atol=1e-5)
    s, e = r[4]
    assert np.allclose(r[3], W.detach().numpy()[s:e], atol=1e-5)
  assert res[0][4] == (0, 6) and res[1][4] == (6, 11)
  assert res[0][1] == res[1][1]          # accuracy over the gathered batch agrees on both shards


def _megatron_worker(rank, world):
  import easyparallellibrary_b200 as epl
  from easyparallellibrary_b200.ops import tensor_parallel as tp
  epl.init(epl.Config({"cluster.colocate_split_and_replicate": True}))
  d, T = 8, 12
  torch.manual_seed(0)
  W1, b1, W2, b2 = torch.randn(4 * d, d) * 0.2, torch.randn(4 * d) * 0.1, torch.randn(d, 4 * d) * 0.2, torch.randn(d) * 0.1
  with epl.split(device_count=world):
    fc1 = tp.ColumnParallelLinear(d, 4 * d, gelu=True)
    fc2 = tp.RowParallelLinear(4 * d, d)
    n = 4 * d // world
    with torch.no_grad():
      fc1.weight.copy_(W1[rank * n:(rank + 1) * n]); fc1.bias.copy_(b1[rank * n:(rank + 1) * n])
      fc2.weight.copy_(W2[:, rank * n:(rank + 1) * n]); fc2.bias.copy_(b2)
  torch.manual_seed(1)
  x = torch.randn(T, d)
  xs = x.chunk(world)[rank].clone().requires_grad_()
  y = fc2(fc1(xs))
  y.pow(2).sum().backward()
  return (y.detach().numpy(), xs.grad.numpy(), fc1.weight.grad.numpy(), fc2.weight.grad.numpy(), fc2.bias.grad.numpy())


def test_column_row_parallel_pair_matches_unsharded():
  res = run_distributed(_megatron_worker, 2)
  d, T = 8, 12
  torch.manual_seed(0)
  W1, b1, W2, b2 = [t.requires_grad_() for t in (torch.randn(4 * d, d) * 0.2, torch.randn(4 * d) * 0.1,
                                                   torch.randn(d, 4 * d) * 0.2, torch.randn(d) * 0.1)]
  torch.manual_seed(1)
  x = torch.randn(T, d, requires_grad=True)
  F = torch.nn.functional
  y = F.linear(F.gelu(F.linear(x, W1, b1), approximate="tanh"), W2, b2)
  y.pow(2).sum().backward()
  n = 4 * d // 2
  for r, out in enumerate(res):
    rows = slice(r * T // 2, (r + 1) * T // 2)
    assert np.allclose(out[0], y.detach().numpy()[rows], atol=1e-5)
    assert np.allclose(out[1], x.grad.numpy()[rows], atol=1e-5)
    assert np.allclose(out[2], W1.grad.numpy()[r * n:(r + 1) * n], atol=1e-5)
    assert np.allclose(out[3], W2.grad.numpy()[:, r * n:(r + 1) * n], atol=1e-5)
  # the replicated bias sees only the local token shard: summing over the split group gives the true gradient
  assert np.allclose(res[0][4] + res[1][4], b2.grad.numpy(), atol=1e-5)


def test_shard_helpers_and_init():
  import easyparallellibrary_b200 as epl
  from easyparallellibrary_b200.ops import tensor_parallel as tp
  assert tp.shard_sizes(10, 4) == [4, 2, 2, 2]
  assert tp.shard_sizes(10, 4, remainder_to_first=False) == [3, 3, 2, 2]
  assert tp.shard_range(11, 2, 1) == (6, 11)
  epl.init(init_process_group=False)
  t = torch.empty(1000, 50)
  tp.distributed_glorot_uniform_(t, 400, 2000)
  assert t.abs().max().item() <= (6.0 / 2400) ** 0.5 + 1e-6
  with epl.split(1):
    w = tp.add_weight((7, 3))
    assert w.shape == (7, 3) and w.epl_tp_shard == (0, 0, 7, 7)


def _moe_worker(rank, world):
  """Expert-parallel MoE layer over split(world): forward value and gradients vs the same layer with all experts local."""
  import easyparallellibrary_b200 as epl
  from easyparallellibrary_b200.ops import moe
  from easyparallellibrary_b200.ops import tensor_parallel as tp
  epl.init(epl.Config({"cluster.colocate_split_and_replicate": True}))
  E, M, F = 4, 8, 16
  torch.manual_seed(3)
  gate_w = torch.randn(M, E) * 0.5
  wi_full, wo_full = torch.randn(E, M, F) * 0.3, torch.randn(E, F, M) * 0.3
  with epl.split(device_count=world):
    layer = moe.MoEFFN(M, F, E, capacity_factor=2.0, gating="top2")
  lo = rank * (E // world)
  with torch.no_grad():
    layer.gate.w.copy_(gate_w.view_as(layer.gate.w))
    layer.wi.copy_(wi_full[lo:lo + E // world])
    layer.wo.copy_(wo_full[lo:lo + E // world])
  # every rank routes its own token group; the unsharded reference processes both groups
  torch.manual_seed(10)
  x_all = torch.randn(world, 2, 6, M)                           # [rank][G, S, M]
  x = x_all[rank].clone().requires_grad_()
  y = layer(x)
  (y.square().sum() + layer.aux_loss).backward()
  return (y.detach().numpy(), x.grad.numpy(), layer.wi.grad.numpy(), layer.wo.grad.numpy(), layer.gate.w.grad.numpy(),
          gate_w.numpy(), wi_full.numpy(), wo_full.numpy(), x_all.numpy())


def test_expert_parallel_moe_matches_unsharded():
  """C9 / examples/moe: dispatch all-to-all -> local experts -> combine all-to-all equals the single-device layer."""
  import easyparallellibrary_b200 as epl
  from easyparallellibrary_b200.ops import moe
  res = run_distributed(_moe_worker, 2)
  gate_w, wi_full, wo_full, x_all = (torch.tensor(a) for a in res[0][5:9])
  epl.init(epl.Config({"cluster.colocate_split_and_replicate": True}), init_process_group=False)
  with epl.split(device_count=1):
    ref = moe.MoEFFN(8, 16, 4, capacity_factor=2.0, gating="top2")
  with torch.no_grad():
    ref.gate.w.copy_(gate_w.view_as(ref.gate.w)); ref.wi.copy_(wi_full); ref.wo.copy_(wo_full)
  wi_g, wo_g = torch.zeros_like(wi_full), torch.zeros_like(wo_full)
  for rank in range(2):
    ref.zero_grad()
    x = x_all[rank].clone().requires_grad_()
    y = ref(x)
    (y.square().sum() + ref.aux_loss).backward()
    np.testing.assert_allclose(res[rank][0], y.detach().numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(res[rank][1], x.grad.numpy(), rtol=1e-4, atol=1e-5)
    wi_g += ref.wi.grad; wo_g += ref.wo.grad
  # expert-weight gradients: rank r owns experts [2r, 2r+2) and accumulates the tokens routed from BOTH ranks
  got_wi = np.concatenate([res[0][2], res[1][2]], 0)
  got_wo = np.concatenate([res[0][3], res[1][3]], 0)
  np.testing.assert_allclose(got_wi, wi_g.numpy(), rtol=1e-4, atol=1e-5)
  np.testing.assert_allclose(got_wo, wo_g.numpy(), rtol=1e-4, atol=1e-5)


def _tp2_dp2_worker(rank, world):
  """split(2) on 4 ranks = two tensor-parallel groups [0,1] and [2,3] that are data-parallel replicas of each other."""
  import easyparallellibrary_b200 as epl
  from easyparallellibrary_b200.models.bert import Bert, BertConfig
  epl.init(epl.Config({"cluster.colocate_split_and_replicate": True}))
  epl.set_default_strategy(epl.replicate(device_count=1))
  torch.manual_seed(0)
  model = Bert(BertConfig.named("tiny", tensor_parallel=2))
  tr = epl.Trainer(model, "sgd", lr=0.1)
  g = torch.Generator().manual_seed(rank // 2)        # one batch per TP group, different between the groups
  losses = []
  for _ in range(3):
    ids = torch.randint(0, 1000, (4, 16), generator=g)
    s, e = torch.randint(0, 16, (4,), generator=g), torch.randint(0, 16, (4,), generator=g)
    losses.append(float(tr.step(ids, s, e).loss))
  flat = torch.cat([p.detach().float().flatten() for p in model.parameters()])
  return flat.numpy(), losses


def test_two_tensor_parallel_groups_are_data_parallel_replicas():
  """TP x DP hybrid: every rank creates both TP process groups collectively (a lazy per-group new_group deadlocks), sharded
  weights are reduced across the groups (ranks 0/2 and 1/3 stay bit-identical), replicated weights across all four."""
  res = run_distributed(_tp2_dp2_worker, 4)
  np.testing.assert_array_equal(res[0][0], res[2][0])
  np.testing.assert_array_equal(res[1][0], res[3][0])
  assert res[0][1] == res[1][1] and res[2][1] == res[3][1] and res[0][1] != res[2][1]   # the loss is per TP group
  assert not np.array_equal(res[0][0], res[1][0])                                        # different shards of the sharded weights
