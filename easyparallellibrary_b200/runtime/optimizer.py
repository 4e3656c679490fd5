"""Flat-buffer optimizers.

The reference ships a BERT-style AdamW built from ~12 unfused TF ops per
variable (``epl/ops/adam_weight_decay_optimizer.py:117-153``).  On B200 the
update is HBM-bound (16-20 B/param), so it is one fused kernel over a *flat
fp32 shard* that also un-scales the gradient, checks it for inf/nan, applies
the update to the fp32 master copy and writes the low-precision model weight —
``ops/csrc/optim.cu``.  The torch composition below is the CPU path and the
fp32 numerics reference for the kernel tests.

Update rule (matches the reference: decoupled weight decay, *no* bias
correction when ``bias_correction=False`` as in BERT's AdamWeightDecay;
``torch.optim.AdamW`` semantics when ``True``)::

    m = b1*m + (1-b1)*g ;  v = b2*v + (1-b2)*g*g
    u = (m/c1) / (sqrt(v/c2) + eps) + wd*p ;  p -= lr*u
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch


@dataclass
class AdamHyper:
  lr: float = 1e-3
  beta1: float = 0.9
  beta2: float = 0.999
  eps: float = 1e-8
  weight_decay: float = 0.0
  bias_correction: bool = True


@dataclass
class SGDHyper:
  lr: float = 1e-2
  momentum: float = 0.0
  weight_decay: float = 0.0


@dataclass
class TorchHyper:
  """Any ``torch.optim.Optimizer`` class, applied to the fp32 master shards (library speed: no fused kernel, no CUDA graph)."""
  lr: float = 1e-3
  weight_decay: float = 0.0
  factory: Optional[type] = None
  kwargs: Optional[dict] = None


def adamw_reference(master: torch.Tensor, grad: torch.Tensor, m: torch.Tensor, v: torch.Tensor, step: int,
                    h: AdamHyper, grad_scale: float = 1.0, decay_mask: Optional[torch.Tensor] = None,
                    model_out: Optional[torch.Tensor] = None) -> None:
  """In place on ``master``/``m``/``v``; optionally writes ``model_out`` (low precision copy)."""
  g = grad.to(torch.float32)
  if grad_scale != 1.0:
    g = g * grad_scale
  m.mul_(h.beta1).add_(g, alpha=1 - h.beta1)
  v.mul_(h.beta2).addcmul_(g, g, value=1 - h.beta2)
  if h.bias_correction:
    c1 = 1 - h.beta1 ** step
    c2 = 1 - h.beta2 ** step
  else:
    c1 = c2 = 1.0
  upd = (m / c1) / ((v / c2).sqrt() + h.eps)
  if h.weight_decay:
    wd = h.weight_decay if decay_mask is None else h.weight_decay * decay_mask
    upd = upd + wd * master
  master.add_(upd, alpha=-h.lr)
  if model_out is not None:
    model_out.copy_(master)


def sgd_reference(master, grad, mom, h: SGDHyper, grad_scale: float = 1.0, model_out=None) -> None:
  g = grad.to(torch.float32)
  if grad_scale != 1.0:
    g = g * grad_scale
  if h.weight_decay:
    g = g + h.weight_decay * master
  if h.momentum:
    mom.mul_(h.momentum).add_(g)
    g = mom
  master.add_(g, alpha=-h.lr)
  if model_out is not None:
    model_out.copy_(master)


class FlatOptimizer(object):
  """Optimizer state for one flat shard ``[lo, hi)`` of a flat parameter buffer."""

  def __init__(self, kind: str, hyper, master_shard: torch.Tensor, decay_mask: Optional[torch.Tensor] = None,
               state_device: Optional[torch.device] = None):
    self.kind = kind.lower()
    self.hyper = hyper
    self.master = master_shard                     # fp32
    self.decay_mask = decay_mask                   # fp32 0/1 per element or None
    dev = state_device or master_shard.device
    self.state_device = dev
    if self.kind in ("adam", "adamw"):
      self.m = torch.zeros_like(master_shard, device=dev)
      self.v = torch.zeros_like(master_shard, device=dev)
    elif self.kind == "sgd":
      self.m = torch.zeros_like(master_shard, device=dev) if hyper.momentum else None
      self.v = None
    elif self.kind == "torch":
      # a wrapped torch optimizer sees the shard as a few fp32 parameters: one per run of equal weight-decay mask values (the
      # mask is constant over each model parameter, and model parameters are contiguous ranges of the flat buffer), in two groups
      self.m = self.v = None
      self._segments = []                          # (lo, hi, nn.Parameter viewing master[lo:hi])
      groups = {True: [], False: []}
      n = master_shard.numel()
      if decay_mask is None or n == 0:
        bounds = [0, n]
        decays = [True]
      else:
        change = (torch.nonzero(decay_mask[1:] != decay_mask[:-1]).flatten() + 1).tolist()
        bounds = [0] + change + [n]
        decays = [bool(decay_mask[b] != 0) for b in bounds[:-1]]
      for (a, b), d in zip(zip(bounds[:-1], bounds[1:]), decays):
        if b > a:
          p = torch.nn.Parameter(master_shard[a:b], requires_grad=True)      # shares the master's storage: updated in place
          self._segments.append((a, b, p))
          groups[d].append(p)
      kw = dict(hyper.kwargs or {})
      param_groups = []
      if groups[True]:
        param_groups.append({"params": groups[True]})
      if groups[False]:
        param_groups.append(dict({"params": groups[False]}, **({"weight_decay": 0.0} if "weight_decay" in kw or hyper.weight_decay else {})))
      if hyper.weight_decay:
        kw["weight_decay"] = hyper.weight_decay
      self.opt = hyper.factory(param_groups or [{"params": [torch.nn.Parameter(master_shard[:0])]}], lr=hyper.lr, **kw)
    else:
      raise ValueError("unknown optimizer %r (adam | adamw | sgd | a torch.optim.Optimizer class)" % kind)
    self.step_count = 0
    self.dyn: Optional[torch.Tensor] = None        # device {lr, inv_c1, inv_c2, grad scale}: set by the engine in CUDA-graph mode

  def step(self, grad_shard: torch.Tensor, model_shard: Optional[torch.Tensor] = None, grad_scale: float = 1.0,
           lo: int = 0, hi: Optional[int] = None, count_step: bool = True) -> None:
    """Apply to elements ``[lo, hi)`` of the shard (grouped apply calls this per group)."""
    if count_step:
      self.step_count += 1
    hi = self.master.numel() if hi is None else hi
    if hi <= lo:
      return
    sl = slice(lo, hi)
    out = model_shard[sl] if model_shard is not None and model_shard.data_ptr() != self.master.data_ptr() else None
    mask = self.decay_mask[sl] if self.decay_mask is not None else None
    if self.kind == "torch":
      if lo != 0 or hi != self.master.numel():
        raise NotImplementedError("optimizer.num_apply_group > 1 needs one of the built-in optimizers (adam | adamw | sgd)")
      g = grad_shard.to(torch.float32)
      if grad_scale != 1.0:
        g = g * grad_scale
      for a, b, p in self._segments:
        p.grad = g[a:b]
      for group in self.opt.param_groups:
        group["lr"] = self.hyper.lr                 # schedules / trainer.lr act on the shared hyper object
      self.opt.step()
      for _, _, p in self._segments:
        p.grad = None
      if out is not None:
        out.copy_(self.master[sl])
      return
    if self.master.is_cuda:
      from easyparallellibrary_b200.ops import fused_optim
      if self.kind == "sgd":
        fused_optim.sgd_step(self.master[sl], grad_shard[sl], None if self.m is None else self.m[sl], self.hyper,
                             grad_scale, out)
      else:
        fused_optim.adamw_step(self.master[sl], grad_shard[sl], self.m[sl], self.v[sl], self.step_count, self.hyper,
                               grad_scale, mask, out, dyn=self.dyn)
      return
    if self.kind == "sgd":
      sgd_reference(self.master[sl], grad_shard[sl], None if self.m is None else self.m[sl], self.hyper, grad_scale, out)
    else:
      adamw_reference(self.master[sl], grad_shard[sl], self.m[sl], self.v[sl], self.step_count, self.hyper,
                      grad_scale, mask, out)

  def state_dict(self):
    sd = {"kind": self.kind, "step": self.step_count, "m": self.m, "v": self.v, "master": self.master}
    if self.kind == "torch":
      # flat "t<param index>.<state name>" entries (tensors / numbers) so the sharded saver can store them like m / v
      for idx, st in self.opt.state_dict()["state"].items():
        for name, val in st.items():
          sd["t%d.%s" % (idx, name)] = val
    return sd

  def load_state_dict(self, sd) -> None:
    self.step_count = int(sd["step"])
    self.master.copy_(sd["master"])
    if self.kind == "torch":
      state = {}
      for k, v in sd.items():
        if k[:1] == "t" and "." in k and k[1:k.index(".")].isdigit():
          state.setdefault(int(k[1:k.index(".")]), {})[k[k.index(".") + 1:]] = v
      if state:
        osd = self.opt.state_dict()
        osd["state"] = state
        self.opt.load_state_dict(osd)
    if self.m is not None and sd.get("m") is not None:
      self.m.copy_(sd["m"])
    if self.v is not None and sd.get("v") is not None:
      self.v.copy_(sd["v"])


def make_hyper(kind: str, **kw):
  kind = kind.lower()
  if kind == "torch":
    kw = dict(kw)
    factory = kw.pop("factory")
    return TorchHyper(lr=kw.pop("lr", 1e-3), weight_decay=kw.pop("weight_decay", 0.0), factory=factory, kwargs=kw)
  if kind == "sgd":
    return SGDHyper(lr=kw.get("lr", 1e-2), momentum=kw.get("momentum", 0.0), weight_decay=kw.get("weight_decay", 0.0))
  betas = kw.get("betas", (kw.get("beta1", 0.9), kw.get("beta2", 0.999)))
  wd = kw.get("weight_decay", 0.01 if kind == "adamw" else 0.0)
  return AdamHyper(lr=kw.get("lr", 1e-3), beta1=betas[0], beta2=betas[1], eps=kw.get("eps", 1e-8),
                   weight_decay=wd, bias_correction=kw.get("bias_correction", True))
