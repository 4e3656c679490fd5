"""Tensor-parallel (``epl.split``) op library.

Reference (``epl/ops/*``): ``Replica2Split`` bridging all-gather
(``bridging_layer.py:46-58``), column-parallel ``distributed_dense`` with the
remainder columns on shard 0 (``distributed_dense.py:99-137``), distributed
softmax cross-entropy (``distributed_losses.py:58-151``: four collectives),
``distributed_argmax`` / ``distributed_equal`` (``distributed_ops.py:58-148``),
fan-in/out-correct sharded Glorot init (``initializers.py:62-72``) and dim-0
sharding of any weight created under ``split`` (``hooks.py:667-707``).

B200 re-design:

* the cross-entropy is **one** pass over the local logits shard that emits
  (max, sum-exp, target-logit) per row, **one** tiny all-gather of ``[rows, 3]``
  and one pass that writes the gradient in place (``csrc/xent.cu`` modes 1/2);
* Megatron-style ``ColumnParallelLinear`` / ``RowParallelLinear`` pairs are
  provided on top of the reference's "giant classifier" pattern, with
  token-sharded activations between TP regions so the two collectives are an
  all-gather feeding a GEMM and a GEMM feeding a reduce-scatter — the shapes the
  fused tcgen05 + NVLink kernels (``ops/tp_fused.py``) implement.
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import torch
from torch import nn

from easyparallellibrary_b200.communicators import functional as CF
from easyparallellibrary_b200.communicators.collective_communicator import get_or_create
from easyparallellibrary_b200.env import Env


# ------------------------------------------------------------------------------------------------
# group discovery
# ------------------------------------------------------------------------------------------------
class TPGroup(object):
  def __init__(self, rank: int, size: int, ranks: List[int]):
    self.rank, self.size, self.ranks = rank, size, ranks
    self._comm = None

  @property
  def comm(self):
    if self._comm is None:
      # dist.new_group is collective over the WORLD and order sensitive.  With several tensor-parallel groups (split(n) on
      # k*n GPUs = k groups that are data-parallel replicas of each other) every rank must create every group, in the same
      # order, before it picks its own — otherwise ranks [0,1] and [2,3] call new_group with different lists and deadlock.
      from easyparallellibrary_b200.communicators.backend import register_groups
      import torch.distributed as dist
      if dist.is_available() and dist.is_initialized() and self.size > 1:
        world = dist.get_world_size()
        if world > self.size and world % self.size == 0:
          copies = 1                                   # the "simple" communicator kind owns a single transport
          register_groups([list(range(b, b + self.size)) for b in range(0, world, self.size)], copies=copies)
      self._comm = get_or_create("TENSOR_PARALLEL", self.ranks, kind="simple")
    return self._comm


_SINGLE = TPGroup(0, 1, [0])


def current_tp_group(strategy=None) -> TPGroup:
  """The tensor-parallel group of the active ``split`` scope (contiguous rank block containing this rank)."""
  env = Env.get()
  strategy = strategy or (env.strategy_context.split_strategy if env.strategy_context else None)
  if strategy is None:
    raise RuntimeError("Got none split strategy from context.")
  cached = getattr(strategy, "_tp_group", None)
  if cached is not None:
    return cached
  cluster = env.cluster
  world = cluster.total_gpu_num if cluster is not None else 1
  me = cluster.rank if cluster is not None and cluster.rank is not None else 0
  n = strategy.device_count or world
  if n > world or world % n:
    if world == 1:
      strategy._tp_group = _SINGLE
      return _SINGLE
    raise RuntimeError("split(device_count=%d) does not fit %d GPUs" % (n, world))
  base = (me // n) * n
  grp = TPGroup(me - base, n, list(range(base, base + n)))
  strategy._tp_group = grp
  return grp


def shard_sizes(total: int, n: int, remainder_to_first: bool = True) -> List[int]:
  q, r = divmod(total, n)
  if remainder_to_first:
    return [q + r] + [q] * (n - 1)
  return [q + (1 if i < r else 0) for i in range(n)]


def shard_range(total: int, n: int, rank: int, remainder_to_first: bool = True) -> Tuple[int, int]:
  sizes = shard_sizes(total, n, remainder_to_first)
  start = sum(sizes[:rank])
  return start, start + sizes[rank]


# ------------------------------------------------------------------------------------------------
# initialisers
# ------------------------------------------------------------------------------------------------
def distributed_glorot_uniform_(shard: torch.Tensor, global_fan_in: int, global_fan_out: int) -> torch.Tensor:
  """Glorot limits from the *global* fan-in/out, so a sharded layer is initialised like the unsharded one."""
  limit = math.sqrt(6.0 / (global_fan_in + global_fan_out))
  return nn.init.uniform_(shard, -limit, limit)


def add_weight(shape, init=None, dtype=None, device=None, group: Optional[TPGroup] = None) -> nn.Parameter:
  """Create a weight under ``split``: dim 0 is sharded, remainder rows go to the lowest ranks."""
  group = group or current_tp_group()
  lo, hi = shard_range(shape[0], group.size, group.rank, remainder_to_first=False)
  p = nn.Parameter(torch.empty((hi - lo,) + tuple(shape[1:]), dtype=dtype, device=device))
  if init is not None:
    init(p.data)
  else:
    fan_in = shape[1] if len(shape) > 1 else shape[0]
    distributed_glorot_uniform_(p.data, fan_in, shape[0])
  p.epl_tp_shard = (0, lo, hi, shape[0])
  return p


# ------------------------------------------------------------------------------------------------
# bridging
# ------------------------------------------------------------------------------------------------
class Replica2Split(nn.Module):
  """Batch-sharded (replica) tensor -> full batch on every split device.  Forward all-gather on dim 0, backward
  reduce-scatter (the autodiff pair of the reference, ``nccl_ops.py:53-62``)."""

  def __init__(self, name: str = "Replica2Split", group: Optional[TPGroup] = None):
    super().__init__()
    self.name, self.group = name, group

  def forward(self, x: torch.Tensor) -> torch.Tensor:
    g = self.group or current_tp_group()
    if g.size == 1:
      return x
    if x.is_floating_point() and x.requires_grad:
      return CF.all_gather(x.contiguous(), g.comm)
    return g.comm.allgather(x.contiguous())


# ------------------------------------------------------------------------------------------------
# column-parallel dense (the reference's distributed_dense)
# ------------------------------------------------------------------------------------------------
_ACT = {None: None, "relu": torch.relu, "gelu": lambda t: torch.nn.functional.gelu(t, approximate="tanh"), "tanh": torch.tanh}


class DistributedDense(nn.Module):
  """``units`` output columns sharded over the split devices; input is gathered from the replicas first."""

  def __init__(self, in_features: int, units: int, activation=None, use_bias: bool = True, gather_input: bool = True,
               group: Optional[TPGroup] = None):
    super().__init__()
    g = group or current_tp_group()
    self.group, self.units, self.in_features, self.gather_input = g, units, in_features, gather_input
    self.start, self.end = shard_range(units, g.size, g.rank, remainder_to_first=True)
    self.weight = nn.Parameter(torch.empty(self.end - self.start, in_features))
    distributed_glorot_uniform_(self.weight.data, in_features, units)
    self.bias = nn.Parameter(torch.zeros(self.end - self.start)) if use_bias else None
    self.activation = _ACT[activation] if not callable(activation) else activation
    self.bridge = Replica2Split("Replica2Split_AllGather_dense", g)
    Env.get().parallel_information.setdefault("INFO_KEY_START_DIM", {})[id(self)] = self.start

  def forward(self, x: torch.Tensor) -> torch.Tensor:
    from easyparallellibrary_b200.ops.linear import linear
    from easyparallellibrary_b200.ops import tp_fused
    g = self.group
    if (self.gather_input and g.size > 1 and x.dim() == 2 and x.is_contiguous() and tp_fused._fused_ok(x, g)
        and self.weight.dtype == x.dtype and x.shape[1] % 8 == 0 and self.weight.shape[0] % 8 == 0):
      # the bridge's all-gather and the GEMM are ONE kernel (K3a: copy CTAs pull the peers' batch shards over NVLink while the
      # tcgen05 tiles whose rows have landed are multiplied); its backward is the fused GEMM -> reduce-scatter (K3b)
      y = tp_fused.all_gather_linear(x, self.weight, self.bias, g)
    else:
      if self.gather_input:
        x = self.bridge(x)
      y = linear(x, self.weight, self.bias)
    y.epl_shard_start = self.start
    return self.activation(y) if self.activation is not None else y


def distributed_dense(x: torch.Tensor, layer: DistributedDense) -> torch.Tensor:
  return layer(x)


# ------------------------------------------------------------------------------------------------
# vocabulary / class parallel softmax cross-entropy (K4)
# ------------------------------------------------------------------------------------------------
class _VocabParallelXent(torch.autograd.Function):
  @staticmethod
  def forward(ctx, logits, labels, start, comm):
    rows, V = logits.shape
    if logits.is_cuda and logits.stride(1) == 1 and V * logits.element_size() <= 200 * 1024 and logits.stride(0) % (16 // logits.element_size()) == 0:
      from easyparallellibrary_b200.ops import _lib
      lib = _lib.require()
      stats = torch.empty(rows, 3, dtype=torch.float32, device=logits.device)
      rc = lib.epl_xent(logits.data_ptr(), labels.data_ptr(), None, None, stats.data_ptr(), None, rows, V, logits.stride(0),
                        1.0, -100, int(start), 1, _lib.dtype_code(logits.dtype), _lib.stream())
      _lib.check(rc, "xent_stats")
      fused = True
    else:
      lf = logits.float()
      mx = lf.max(1).values
      local = labels - start
      has = (local >= 0) & (local < V)
      tgt = torch.where(has, lf.gather(1, local.clamp(0, V - 1).unsqueeze(1)).squeeze(1), torch.zeros_like(mx))
      stats = torch.stack([mx, (lf - mx.unsqueeze(1)).exp().sum(1), tgt], 1)
      fused = False
    allstats = comm.allgather(stats.unsqueeze(0)) if comm.size > 1 else stats.unsqueeze(0)     # [N, rows, 3] — the one collective
    gmax = allstats[:, :, 0].max(0).values
    gsum = (allstats[:, :, 1] * (allstats[:, :, 0] - gmax).exp()).sum(0)
    target = allstats[:, :, 2].sum(0)
    loss = gsum.log() + gmax - target
    ctx.save_for_backward(logits, labels, torch.stack([gmax, gsum], 1).contiguous())
    ctx.start, ctx.fused = start, fused
    return loss

  @staticmethod
  def backward(ctx, gloss):
    logits, labels, gstats = ctx.saved_tensors
    rows, V = logits.shape
    if ctx.fused:
      from easyparallellibrary_b200.ops import _lib
      lib = _lib.require()
      rc = lib.epl_xent(logits.data_ptr(), labels.data_ptr(), None, logits.data_ptr(), None, gstats.data_ptr(), rows, V,
                        logits.stride(0), 1.0, -100, int(ctx.start), 2, _lib.dtype_code(logits.dtype), _lib.stream())
      _lib.check(rc, "xent_grad")
      grad = logits * gloss.unsqueeze(1).to(logits.dtype)
    else:
      p = (logits.float() - gstats[:, :1]).exp() / gstats[:, 1:2]
      local = labels - ctx.start
      has = (local >= 0) & (local < V)
      onehot = torch.zeros_like(p)
      onehot[has, local[has]] = 1.0
      grad = ((p - onehot) * gloss.unsqueeze(1)).to(logits.dtype)
    return grad, None, None, None


def distributed_sparse_softmax_cross_entropy_with_logits(labels: torch.Tensor, logits: torch.Tensor, weights=1.0,
                                                         group: Optional[TPGroup] = None, start: Optional[int] = None,
                                                         gather_labels: bool = True, reduction: str = "mean"):
  """``logits`` is this device's class shard ``[rows, C_local]``; ``labels`` are global class ids (batch-sharded
  like the layer input unless ``gather_labels=False``).  Returns the loss averaged over the full (gathered) batch —
  identical on every split device, like the reference's final ``LOSS_REDUCE`` all-reduce."""
  if labels is None:
    raise ValueError("Labels must not be None.")
  if logits is None:
    raise ValueError("Logits must not be None.")
  g = group or current_tp_group()
  if start is None:
    start = getattr(logits, "epl_shard_start", None)
    if start is None:
      sizes = g.comm.allgather(torch.tensor([logits.shape[1]], device=logits.device)) if g.size > 1 else torch.tensor([logits.shape[1]])
      start = int(sizes[:g.rank].sum())
  if gather_labels and g.size > 1:
    labels = g.comm.allgather(labels.contiguous())
  labels = labels.reshape(-1).to(torch.int64)
  loss = _VocabParallelXent.apply(logits.reshape(-1, logits.shape[-1]), labels, start, g.comm)
  if torch.is_tensor(weights) or weights != 1.0:
    loss = loss * weights
  if reduction == "mean":
    return loss.mean()
  if reduction == "sum":
    return loss.sum()
  return loss


def distributed_argmax(logits: torch.Tensor, axis: int = 1, group: Optional[TPGroup] = None, start: Optional[int] = None):
  """Global arg-max over class-sharded logits: one all-gather of ``[rows, 2]`` (value, global index)."""
  g = group or current_tp_group()
  if start is None:
    start = getattr(logits, "epl_shard_start", 0)
  val, idx = logits.float().max(axis)
  if g.size == 1:
    return idx + start
  packed = torch.stack([val, (idx + start).float()], -1)
  allp = g.comm.allgather(packed.unsqueeze(0))                     # [N, rows, 2]
  win = allp[..., 0].argmax(0)
  return allp[..., 1].gather(0, win.unsqueeze(0)).squeeze(0).to(torch.int64)


def distributed_equal(predictions: torch.Tensor, labels: torch.Tensor, group: Optional[TPGroup] = None,
                      gather_labels: bool = True) -> torch.Tensor:
  if labels is None:
    raise ValueError("Labels must not be None.")
  if predictions is None:
    raise ValueError("Logits must not be None.")
  g = group or current_tp_group()
  if gather_labels and g.size > 1:
    labels = g.comm.allgather(labels.contiguous())
  return predictions.to(labels.dtype) == labels


# ------------------------------------------------------------------------------------------------
# Megatron-style pairs with token-sharded activations (sequence parallel form)
# ------------------------------------------------------------------------------------------------
class ColumnParallelLinear(nn.Module):
  """``y[:, shard] = gather(x) @ W[shard]^T``.  Input: token shard ``[T/N, in]``; output: ``[T, out/N]``."""

  def __init__(self, in_features: int, out_features: int, bias: bool = True, gelu: bool = False, group: Optional[TPGroup] = None,
               init_std: Optional[float] = None):
    super().__init__()
    g = group or current_tp_group()
    if out_features % g.size:
      raise ValueError("out_features must be divisible by the split size")
    self.group, self.gelu = g, gelu
    self.weight = nn.Parameter(torch.empty(out_features // g.size, in_features))
    if init_std is not None:
      nn.init.normal_(self.weight, std=init_std)
    else:
      distributed_glorot_uniform_(self.weight.data, in_features, out_features)
    self.bias = nn.Parameter(torch.zeros(out_features // g.size)) if bias else None

  def forward(self, x_shard: torch.Tensor) -> torch.Tensor:
    from easyparallellibrary_b200.ops import tp_fused
    return tp_fused.all_gather_linear(x_shard, self.weight, self.bias, self.group, gelu=self.gelu)


class RowParallelLinear(nn.Module):
  """``y = reduce_scatter(x[:, shard] @ W[:, shard]^T) + b``.  Input ``[T, in/N]``; output token shard ``[T/N, out]``."""

  def __init__(self, in_features: int, out_features: int, bias: bool = True, group: Optional[TPGroup] = None,
               init_std: Optional[float] = None):
    super().__init__()
    g = group or current_tp_group()
    if in_features % g.size:
      raise ValueError("in_features must be divisible by the split size")
    self.group = g
    self.weight = nn.Parameter(torch.empty(out_features, in_features // g.size))
    if init_std is not None:
      nn.init.normal_(self.weight, std=init_std)
    else:
      distributed_glorot_uniform_(self.weight.data, in_features, out_features)
    self.bias = nn.Parameter(torch.zeros(out_features)) if bias else None
    if self.bias is not None:
      self.bias.epl_tp_replicated = True          # identical on every split device: gradient is all-reduced over the TP group

  def forward(self, x: torch.Tensor) -> torch.Tensor:
    from easyparallellibrary_b200.ops import tp_fused
    y = tp_fused.linear_reduce_scatter(x, self.weight, self.group)
    return y + self.bias if self.bias is not None else y
