"""ZeRO stage 3 (parameter partitioning) — a B200 extension; the reference stops at v1
(``epl/config.py:132-137``: "v2 partitions weights, gradients and optimizer states. Now v0 and v1 are supported").

Every *unit* (a layer of the stage: one element of the model's layer sequence, or a direct child) keeps only a
1/N flat shard of its weights, gradients, fp32 master weights and Adam moments.  Around a unit's forward, and
again around its backward, the full weights are materialised by an all-gather into a transient buffer and the
parameters are re-pointed at views of it; when the last gradient of the unit has been produced the full gradient
is reduce-scattered into the shard and both transient buffers are dropped.  HBM per rank is therefore
``16 B x params / N`` of persistent state plus two layers' worth of transients.

The gather that feeds a layer whose first op is a GEMM can run inside that GEMM (``ops/tp_kernels.py:
ag_weight_gemm`` — the same copy-CTA + flag mechanism as the all-gather->GEMM kernel with the roles of A and B
swapped), so the NVLink transfer hides behind the tiles whose weights are already local.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
from torch import nn

import os

from easyparallellibrary_b200.parallel.flat import ALIGN_ELEMS
from easyparallellibrary_b200.runtime.optimizer import FlatOptimizer

# K2 integration.  A unit that owns exactly one weight matrix — the weight of the first GEMM of a layer — can DEFER its forward
# all-gather: the parameter is pointed at the (still empty) transient buffer and tagged ``epl_pending_gather``; ops.linear
# resolves the tag by running the fused weight-gather GEMM (``ops/tp_kernels.ag_weight_gemm``), which fills the buffer over
# NVLink while it multiplies, or — whenever the fused kernel does not apply (CPU, residual epilogue, unsupported dtype) — by a
# plain all-gather.  ``GATHER_GEMM_IMPL`` is the implementation hook: ``None`` = pick the kernel on eligible GPUs; the CPU
# tests install an emulation to exercise the protocol.  Off by default (``zero.fused_gather`` /
# ``EPL_ZERO3_FUSED_GATHER=1``) until the integrated path has run on hardware.
FUSE_FIRST_GEMM = os.environ.get("EPL_ZERO3_FUSED_GATHER", "0") == "1"
GATHER_GEMM_IMPL = None


class PendingGather(object):
  """Tag on a parameter whose ZeRO-3 all-gather has been deferred into its first consumer."""

  def __init__(self, unit: "Zero3Unit"):
    self.unit = unit

  def _impl(self, x: torch.Tensor):
    if GATHER_GEMM_IMPL is not None:
      return GATHER_GEMM_IMPL
    u = self.unit
    if x.is_cuda and x.dtype == torch.bfloat16 and u.dtype == torch.bfloat16 and u.tp_group is not None:
      from easyparallellibrary_b200.ops import tp_kernels
      if tp_kernels.available(u.tp_group):
        return tp_kernels.ag_weight_gemm
    return None

  def can_fuse(self, x: torch.Tensor) -> bool:
    return self._impl(x) is not None

  def materialize(self) -> None:
    u = self.unit
    u.comm.primary.all_gather_into(u.full, u.shard_param)
    u.params[0].epl_pending_gather = None

  def gemm(self, x2: torch.Tensor, bias, gelu: bool):
    """``(y, pre)`` with ``y = x2 @ W^T (+bias)(gelu)``; W is gathered into the unit's transient buffer on the way."""
    u = self.unit
    n, k = u.shapes[0]
    y, pre, _ = self._impl(x2)(x2, u.shard_param.view(n // u.comm.size, k), u.tp_group, bias, gelu, out_w_full=u.full[:n * k])
    u.params[0].epl_pending_gather = None
    return y, pre


class Zero3Unit(object):
  def __init__(self, index: int, module: nn.Module, params: List[nn.Parameter], comm, opt_kind, hyper, no_decay, device, offload: bool,
               deferred: bool = False, tp_group=None):
    self.index, self.module, self.params, self.comm = index, module, params, comm
    self.deferred, self.tp_group = deferred, tp_group          # deferred: a single [N, K] weight, N % W == 0 (row shards)
    self.dtype = params[0].dtype
    self.device = device
    W = comm.size
    quantum = W * ALIGN_ELEMS
    self.offsets, off = [], 0
    for p in params:
      self.offsets.append(off)
      off += (p.numel() + ALIGN_ELEMS - 1) // ALIGN_ELEMS * ALIGN_ELEMS
    self.numel = (off + quantum - 1) // quantum * quantum
    self.shard_numel = self.numel // W
    lo = comm.rank * self.shard_numel
    full = torch.zeros(self.numel, dtype=self.dtype, device=device)
    for p, o in zip(params, self.offsets):
      full[o:o + p.numel()].copy_(p.data.reshape(-1))
    if W > 1:
      comm.primary.broadcast(full, 0)
    self.shard_param = full[lo:lo + self.shard_numel].clone()
    self.shard_grad = torch.zeros(self.shard_numel, dtype=self.dtype, device=device)
    mask = None
    if any(no_decay(p) for p in params):
      m = torch.ones(self.numel, dtype=torch.float32, device=device)
      for p, o in zip(params, self.offsets):
        if no_decay(p):
          m[o:o + p.numel()] = 0
      mask = m[lo:lo + self.shard_numel].clone()
    master = self.shard_param.to(torch.float32)
    if offload:
      from easyparallellibrary_b200.runtime.offload import OffloadedOptimizer
      self.opt = OffloadedOptimizer(opt_kind, hyper, master, mask, device)
    else:
      self.opt = FlatOptimizer(opt_kind, hyper, master, mask)
    self.full: Optional[torch.Tensor] = None
    self.ready = 0
    self.shapes = [p.shape for p in params]
    del full
    self.release()

  # -- materialise / release -------------------------------------------------------------------------
  def gather(self, defer: bool = False) -> None:
    if self.full is not None:
      pend = getattr(self.params[0], "epl_pending_gather", None) if self.deferred else None
      if pend is not None and not defer:
        pend.materialize()                       # e.g. the backward pass needs the weight although forward never consumed it
      return
    self.full = torch.empty(self.numel, dtype=self.dtype, device=self.device)
    if not (defer and self.deferred and self.comm.size > 1):
      self.comm.primary.all_gather_into(self.full, self.shard_param)
    for p, o, shp in zip(self.params, self.offsets, self.shapes):
      p.data = self.full[o:o + shp.numel()].view(shp)
    if defer and self.deferred and self.comm.size > 1:
      self.params[0].epl_pending_gather = PendingGather(self)

  def release(self) -> None:
    if self.deferred and getattr(self.params[0], "epl_pending_gather", None) is not None:
      self.params[0].epl_pending_gather = None
      if self.full is not None:
        raise RuntimeError("ZeRO-3: the deferred weight of unit %d was never consumed through ops.linear during the forward pass; "
                           "mark the module with epl_no_fused_gather = True" % self.index)
    self.full = None
    for p in self.params:
      p.data = torch.empty(0, dtype=self.dtype, device=self.device)

  def reduce_grads(self) -> None:
    """Full gradients -> this rank's shard (accumulated over micro-batches)."""
    flat = torch.zeros(self.numel, dtype=self.dtype, device=self.device)
    for p, o in zip(self.params, self.offsets):
      if p.grad is not None:
        flat[o:o + p.numel()].copy_(p.grad.reshape(-1))
        p.grad = None
    if self.comm.size > 1:
      out = torch.empty(self.shard_numel, dtype=self.dtype, device=self.device)
      self.comm.primary.reduce_scatter_into(out, flat, "sum")
    else:
      out = flat
    self.shard_grad.add_(out)
    self.ready = 0


class Zero3Engine(object):
  def __init__(self, trainer, stage: int, units: List[nn.Module], comm):
    self.trainer, self.stage, self.comm = trainer, stage, comm
    cfg = trainer.config
    seen = set()
    self.units: List[Zero3Unit] = []
    self._unit_of: Dict[int, Zero3Unit] = {}
    self.tp_group = None
    if comm.size > 1:
      from easyparallellibrary_b200.ops.tensor_parallel import TPGroup
      self.tp_group = TPGroup(comm.rank, comm.size, list(comm.ranks))     # the view of the DP group the fused kernel expects
      self.tp_group._comm = comm
    for m in units:
      ps = [p for p in m.parameters() if p.requires_grad and id(p) not in seen]
      for p in ps:
        seen.add(id(p))
      if not ps:
        continue
      first = self._first_gemm_weight(m, ps) if (cfg.zero.fused_gather or FUSE_FIRST_GEMM or GATHER_GEMM_IMPL is not None) else None
      groups = [([first], True), ([p for p in ps if p is not first], False)] if first is not None else [(ps, False)]
      for plist, deferred in groups:
        if not plist:
          continue
        u = Zero3Unit(len(self.units), m, plist, comm, trainer.opt_kind, trainer.hyper, trainer.no_decay, trainer.device,
                      cfg.offload.level == "v0", deferred=deferred, tp_group=self.tp_group)
        self.units.append(u)
        for p in plist:
          self._unit_of[id(p)] = u
          p.register_post_accumulate_grad_hook(self._on_grad)
    # a module may also use weights owned by an earlier unit (tied embeddings): it gathers every owner it needs
    for m in units:
      owners = []
      for p in m.parameters():
        u = self._unit_of.get(id(p))
        if u is not None and u not in owners:
          owners.append(u)
      if not owners:
        continue
      m.register_forward_pre_hook(self._pre_forward(owners))
      m.register_forward_hook(self._post_forward(owners))
      m.register_full_backward_pre_hook(self._pre_backward(owners))

  def _first_gemm_weight(self, m: nn.Module, ps: List[nn.Parameter]) -> Optional[nn.Parameter]:
    """The weight of the first ``ops.linear.Linear`` of the layer, if its rows shard evenly (K2's layout)."""
    from easyparallellibrary_b200.ops.linear import Linear
    W = self.comm.size
    if W <= 1 or getattr(m, "epl_no_fused_gather", False):
      return None
    mine = {id(p) for p in ps}
    for sub in m.modules():
      if isinstance(sub, Linear) and id(sub.weight) in mine:
        n, k = sub.weight.shape
        ok = n % W == 0 and k % 8 == 0 and (n // W) * k % ALIGN_ELEMS == 0 and not getattr(sub, "epl_no_fused_gather", False)
        return sub.weight if ok else None
    return None

  def _pre_forward(self, owners):
    def hook(mod, args):
      for u in owners:
        u.gather(defer=True)
    return hook

  def _post_forward(self, owners):
    def hook(mod, args, out):
      for u in owners:
        u.release()
    return hook

  def _pre_backward(self, owners):
    def hook(mod, grad_out):
      for u in owners:
        u.gather()
    return hook

  def _on_grad(self, p) -> None:
    u = self._unit_of[id(p)]
    u.ready += 1
    if u.ready == len(u.params):
      u.reduce_grads()
      u.release()

  # -- trainer interface ----------------------------------------------------------------------------------
  def zero_grad(self) -> None:
    for u in self.units:
      u.shard_grad.zero_()
      u.ready = 0

  def finish_backward(self) -> None:
    """Units whose parameters did not all receive a gradient (unused branches) are flushed here."""
    for u in self.units:
      if u.ready or any(p.grad is not None for p in u.params):
        u.reduce_grads()
        u.release()

  def grad_sq_norm(self) -> torch.Tensor:
    sq = torch.zeros(1, device=self.trainer.device, dtype=torch.float32)
    for u in self.units:
      sq += u.shard_grad.float().pow(2).sum()
    if self.comm.size > 1:
      self.comm.primary.all_reduce(sq, "sum")
    return sq

  def has_non_finite(self) -> torch.Tensor:
    bad = torch.zeros(1, device=self.trainer.device)
    for u in self.units:
      bad += (~torch.isfinite(u.shard_grad)).any().float()
    return bad

  def apply(self, scale: float) -> None:
    for u in self.units:
      u.opt.step(u.shard_grad, u.shard_param, scale)

  def state_dict(self):
    return [dict(u.opt.state_dict(), shard_param=u.shard_param) for u in self.units]

  def load_state_dict(self, sds) -> None:
    for u, sd in zip(self.units, sds):
      u.opt.load_state_dict(sd)
      u.shard_param.copy_(sd["shard_param"])

  def gather_all(self) -> None:
    """Materialise every unit (evaluation / checkpoint export)."""
    for u in self.units:
      u.gather()

  def release_all(self) -> None:
    for u in self.units:
      u.release()

  def persistent_bytes(self) -> int:
    n = 0
    for u in self.units:
      n += u.shard_numel * (2 * u.shard_param.element_size())
      for t in (u.opt.master, u.opt.m, u.opt.v):
        if t is not None and t.device.type != "cpu":
          n += t.numel() * t.element_size()
    return n
