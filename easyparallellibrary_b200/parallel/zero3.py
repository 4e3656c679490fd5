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
import warnings

# the backward PRE-hooks below also sit on the first unit, whose inputs (token ids) need no gradient; torch then warns that a
# full backward hook "is firing when gradients are computed with respect to module outputs" — exactly what is wanted here
warnings.filterwarnings("ignore", message="Full backward hook is firing when gradients are computed with respect to module outputs")

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
    u.comm.primary.all_gather_into(u.full, u._device_shard())
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
    # offload.level=v0 (reference graph_editor.py:727-751, "variables on the host, read just in time"): the weight shard lives
    # in pinned host memory; every gather starts with an H2D copy of the shard (prefetched one unit ahead on the copy stream)
    self.offload = bool(offload)
    self.shard_param = full[lo:lo + self.shard_numel].clone()
    if self.offload:
      host = torch.empty(self.shard_numel, dtype=self.dtype, pin_memory=device.type == "cuda")
      host.copy_(self.shard_param)
      self.shard_host = host
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
    self.full_grad: Optional[torch.Tensor] = None
    self.work = None                  # in-flight asynchronous all-gather into self.full
    self.bound = False                # parameters point into self.full
    self.ready = 0
    self.shapes = [p.shape for p in params]
    del full
    if self.offload:
      self.shard_param = None         # device copy exists only between a gather's H2D and its all-gather
    self.release()

  # -- materialise / release -------------------------------------------------------------------------
  def _device_shard(self, stream_ctx=None) -> torch.Tensor:
    if not self.offload:
      return self.shard_param
    dev = torch.empty(self.shard_numel, dtype=self.dtype, device=self.device)
    dev.copy_(self.shard_host, non_blocking=True)
    return dev

  def gather_async(self, copy_stream=None) -> None:
    """Start the all-gather of this unit's weights (no-op if already started); ``gather()`` completes it.  The engine calls
    this one unit ahead of the compute (forward: next unit, backward: previous unit) so the transfer hides behind it."""
    if self.full is not None or self.comm.size <= 1:
      return
    self.full = torch.empty(self.numel, dtype=self.dtype, device=self.device)
    be = self.comm.pool.backends[-1]               # second transport of the pool when there is one: not behind the reduce-scatters
    if copy_stream is not None:
      copy_stream.wait_stream(torch.cuda.current_stream())     # the buffer's previous owner (allocator reuse) is done with it
      with torch.cuda.stream(copy_stream):
        shard = self._device_shard()
        self.work = be.all_gather_into(self.full, shard, async_op=True)
        shard.record_stream(copy_stream) if shard.is_cuda and self.offload else None
    else:
      self.work = be.all_gather_into(self.full, self._device_shard(), async_op=True)

  def gather(self, defer: bool = False) -> None:
    if self.full is not None and self.bound:
      pend = getattr(self.params[0], "epl_pending_gather", None) if self.deferred else None
      if pend is not None and not defer:
        pend.materialize()                       # e.g. the backward pass needs the weight although forward never consumed it
      return
    deferred_now = defer and self.deferred and self.comm.size > 1 and not self.offload
    if self.full is None:
      self.full = torch.empty(self.numel, dtype=self.dtype, device=self.device)
      if not deferred_now:
        self.comm.primary.all_gather_into(self.full, self._device_shard())
    elif self.work is not None:
      self.work.wait()                           # prefetched: the current stream waits for the gather, the host does not
    self.work = None
    for p, o, shp in zip(self.params, self.offsets, self.shapes):
      p.data = self.full[o:o + shp.numel()].view(shp)
    self.bound = True
    if deferred_now:
      self.params[0].epl_pending_gather = PendingGather(self)

  def release(self) -> None:
    if self.deferred and getattr(self.params[0], "epl_pending_gather", None) is not None:
      self.params[0].epl_pending_gather = None
      if self.full is not None:
        raise RuntimeError("ZeRO-3: the deferred weight of unit %d was never consumed through ops.linear during the forward pass; "
                           "mark the module with epl_no_fused_gather = True" % self.index)
    if self.work is not None:                     # a prefetched gather nobody consumed: let it finish before the buffer goes
      self.work.wait()
      self.work = None
    self.full = None
    self.bound = False
    for p in self.params:
      p.data = torch.empty(0, dtype=self.dtype, device=self.device)

  def bind_grads(self) -> None:
    """Before the unit's backward: ``.grad`` of every parameter becomes a view of one flat buffer, so autograd accumulates
    straight into the reduce-scatter's input (no per-parameter copy afterwards)."""
    if self.full_grad is None:
      self.full_grad = torch.zeros(self.numel, dtype=self.dtype, device=self.device)
    for p, o, shp in zip(self.params, self.offsets, self.shapes):
      if p.grad is None:
        p.grad = self.full_grad[o:o + shp.numel()].view(shp)

  def reduce_grads(self) -> None:
    """Full gradients -> this rank's shard (accumulated over micro-batches)."""
    flat = self.full_grad if self.full_grad is not None else torch.zeros(self.numel, dtype=self.dtype, device=self.device)
    for p, o in zip(self.params, self.offsets):
      if p.grad is not None:
        if p.grad.data_ptr() != flat[o:].data_ptr():          # autograd replaced the view (or bind_grads never ran): fold it in
          flat[o:o + p.numel()].copy_(p.grad.reshape(-1))
        p.grad = None
    self.full_grad = None
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
    self.prefetch = os.environ.get("EPL_ZERO3_PREFETCH", "1") != "0" and comm.size > 1
    self.copy_stream = torch.cuda.Stream(device=trainer.device) if (trainer.device.type == "cuda" and cfg.offload.level == "v0") else None
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

  def _prefetch(self, index: int) -> None:
    if 0 <= index < len(self.units) and self.prefetch:
      u = self.units[index]
      if not (u.deferred and not u.offload):      # a deferred (K2) unit gathers inside its GEMM
        u.gather_async(self.copy_stream)

  def _pre_forward(self, owners):
    def hook(mod, args):
      for u in owners:
        u.gather(defer=True)
      if torch.is_grad_enabled() or True:
        self._prefetch(max(u.index for u in owners) + 1)
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
        if u.module is mod:
          u.bind_grads()
      self._prefetch(min(u.index for u in owners) - 1)
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

  def grad_sq_norm(self, local: bool = False, divisor: int = 1) -> torch.Tensor:
    """Squared norm of the (reduce-scattered, i.e. already summed over replicas) gradient.  ZeRO-3 never holds a
    replica-local gradient — units are reduced as soon as their backward finishes — so for clip-then-reduce (``local``)
    the engine passes ``divisor`` = number of replicas under ``mean``: the norm of the averaged gradient stands in for the
    per-replica norm."""
    sq = torch.zeros(1, device=self.trainer.device, dtype=torch.float32)
    for u in self.units:
      sq += u.shard_grad.float().pow(2).sum()
    if self.comm.size > 1:
      self.comm.primary.all_reduce(sq, "sum")
    return sq / float(divisor * divisor)

  def scale_grads(self, c: float) -> None:
    for u in self.units:
      u.shard_grad.mul_(c)

  def has_non_finite(self) -> torch.Tensor:
    bad = torch.zeros(1, device=self.trainer.device)
    for u in self.units:
      bad += (~torch.isfinite(u.shard_grad)).any().float()
    return bad

  def apply(self, scale: float) -> None:
    for u in self.units:
      if u.offload:                               # new weights are produced on the device, then parked on the host again
        out = torch.empty(u.shard_numel, dtype=u.dtype, device=u.device)
        u.opt.step(u.shard_grad, out, scale)
        u.shard_host.copy_(out, non_blocking=True)
      else:
        u.opt.step(u.shard_grad, u.shard_param, scale)

  def state_dict(self):
    return [dict(u.opt.state_dict(), shard_param=(u.shard_host if u.offload else u.shard_param)) for u in self.units]

  def load_state_dict(self, sds) -> None:
    for u, sd in zip(self.units, sds):
      u.opt.load_state_dict(sd)
      (u.shard_host if u.offload else u.shard_param).copy_(sd["shard_param"])

  def gather_all(self) -> None:
    """Materialise every unit (evaluation / checkpoint export)."""
    for u in self.units:
      u.gather()

  def release_all(self) -> None:
    for u in self.units:
      u.release()

  def persistent_bytes(self) -> int:
    """Device bytes this rank holds between steps: weight shard (unless offloaded) + gradient shard + optimizer state
    (unless offloaded) + decay mask."""
    n = 0
    for u in self.units:
      es = u.shard_grad.element_size()
      n += u.shard_numel * es * (1 if u.offload else 2)
      for t in (u.opt.master, u.opt.m, u.opt.v, getattr(u.opt, "decay_mask", None), getattr(u.opt, "device_mask", None)):
        if t is not None and t.device.type != "cpu":
          n += t.numel() * t.element_size()
    return n
