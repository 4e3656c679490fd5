"""Mixture-of-experts building blocks with expert parallelism over ``epl.split``.

Reference: ``examples/moe/moe_ffn.py`` — ``Top2Gating`` (127-269) / ``SwitchGating`` (270-369) with capacity factor
and auxiliary load-balancing loss, ``MoEFFN`` (370-464) whose expert weights are created under ``epl.split``
(dim 0 = experts, ``hooks.py:667-707``) and whose three einsums trigger an all-to-all *dispatch* before the first
and an all-to-all *combine* before the third (``hooks.py:758-794``).

Here the layer is explicit: gating -> dispatch einsum -> all-to-all -> expert FFN (batched GEMMs over the local
experts) -> all-to-all -> combine einsum.  On B200 the all-to-all is one hand-written kernel that stores each
segment straight into the destination rank's symmetric receive buffer over NVLink
(``csrc/symm.cu: alltoall_p2p_kernel``); NCCL send/recv is the CPU/baseline path.  The all-to-all is never
recomputed by gradient checkpointing (``epl_collective``), like the reference (``constant.py:97``).
"""
from __future__ import annotations

import ctypes
from typing import Dict

import torch
from torch import nn

from easyparallellibrary_b200.communicators import functional as CF

# The peer-store all-to-all kernel has not been exercised on hardware yet (tools/mgpu_check.py moe is the check); until it
# has, NCCL grouped send/recv is the default transport and EPL_MOE_P2P=1 opts in.
import os as _os
USE_P2P_KERNEL = _os.environ.get("EPL_MOE_P2P", "0") == "1"
_A2A_WS: Dict[int, "_A2AWorkspace"] = {}


class _A2AWorkspace(object):
  def __init__(self, group, device):
    from easyparallellibrary_b200.runtime.symmetric import SignalPad
    self.group, self.device = group, device
    self.pg = getattr(group.comm.primary, "group", None)
    self.pad = SignalPad(1, group.ranks, device, group=self.pg)
    self.sync = torch.zeros(4, dtype=torch.int32, device=device)
    self.recv = None
    self.epoch = 0

  def recv_buffer(self, nbytes: int):
    from easyparallellibrary_b200.runtime.symmetric import SymmetricBuffer
    if self.recv is None or self.recv.nbytes < nbytes:
      torch.cuda.synchronize(self.device)
      self.recv = SymmetricBuffer(max(nbytes, 1 << 20), self.group.ranks, self.device, group=self.pg)
    return self.recv


def _p2p_all_to_all(t: torch.Tensor, group) -> torch.Tensor:
  from easyparallellibrary_b200.ops import _lib
  from easyparallellibrary_b200.runtime.symmetric import _sym_lib
  lib = _sym_lib()
  if not hasattr(lib, "_a2a_ready"):
    lib.epl_alltoall_p2p.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_uint, ctypes.c_int, ctypes.c_void_p]
    lib._a2a_ready = True
  ws = _A2A_WS.get(id(group))
  if ws is None:
    ws = _A2A_WS[id(group)] = _A2AWorkspace(group, t.device)
  t = t.contiguous()
  nbytes = t.numel() * t.element_size()
  seg = nbytes // group.size
  buf = ws.recv_buffer(nbytes)
  ws.epoch += 1
  rc = lib.epl_alltoall_p2p(t.data_ptr(), buf.peer_table(0), ws.pad.slot_table(0), ws.sync.data_ptr(), seg, group.rank, group.size,
                            ws.epoch, 64, _lib.stream())
  _lib.check(rc, "alltoall_p2p")
  return buf.tensor(t.dtype, t.numel()).view(t.shape).clone()


class _ExpertAllToAll(torch.autograd.Function):
  @staticmethod
  def forward(ctx, t, group):
    ctx.group = group
    return expert_all_to_all_raw(t, group)

  @staticmethod
  def backward(ctx, g):
    return expert_all_to_all_raw(g.contiguous(), ctx.group), None


def expert_all_to_all_raw(t: torch.Tensor, group) -> torch.Tensor:
  if group.size == 1:
    return t
  nbytes = t.numel() * t.element_size()
  if USE_P2P_KERNEL and t.is_cuda and group.size <= 8 and (nbytes // group.size) % 16 == 0:
    return _p2p_all_to_all(t, group)
  return group.comm.alltoall(t.contiguous())


def expert_all_to_all(t: torch.Tensor, group) -> torch.Tensor:
  """dim 0 is split into ``group.size`` equal segments; segment j goes to rank j (differentiable)."""
  if group.size == 1:
    return t
  return _ExpertAllToAll.apply(t, group)


# --------------------------------------------------------------------------------------------------------
# gating
# --------------------------------------------------------------------------------------------------------
def _capacity(tokens: int, experts: int, factor: float, minimum: int = 4) -> int:
  return max(int(tokens * factor / experts), minimum)


class SwitchGating(nn.Module):
  """Top-1 routing.  Returns (dispatch [G,S,E,C] bool-ish, combine [G,S,E,C], aux_loss)."""

  def __init__(self, d_model: int, num_experts: int, capacity_factor: float = 1.25):
    super().__init__()
    self.w = nn.Parameter(torch.empty(d_model, num_experts))
    nn.init.normal_(self.w, std=d_model ** -0.5)
    self.E, self.cf = num_experts, capacity_factor

  def forward(self, x):                                   # [G, S, M]
    G, S, _ = x.shape
    C = _capacity(S, self.E, self.cf)
    gates = torch.softmax((x.float() @ self.w.float()), -1)          # [G,S,E]
    idx = gates.argmax(-1)
    mask = torch.nn.functional.one_hot(idx, self.E).float()
    density, density_proxy = mask.mean(1), gates.mean(1)
    aux = (density * density_proxy).mean() * self.E * self.E
    pos = torch.cumsum(mask, 1) * mask - mask                         # position of each token inside its expert
    mask = mask * (pos < C)
    gate = (gates * mask).sum(-1, keepdim=True)
    pos_oh = torch.nn.functional.one_hot(pos.sum(-1).long().clamp(max=C - 1), C).float()
    combine = gate.unsqueeze(-1) * mask.unsqueeze(-1) * pos_oh.unsqueeze(2)   # [G,S,E,C]
    return (combine > 0).to(x.dtype), combine.to(x.dtype), aux


class Top2Gating(nn.Module):
  """Top-2 routing with capacity (second choice taken after first choices are placed)."""

  def __init__(self, d_model: int, num_experts: int, capacity_factor: float = 1.25):
    super().__init__()
    self.w = nn.Parameter(torch.empty(d_model, num_experts))
    nn.init.normal_(self.w, std=d_model ** -0.5)
    self.E, self.cf = num_experts, capacity_factor

  def forward(self, x):
    G, S, _ = x.shape
    C = _capacity(2 * S, self.E, self.cf)
    gates = torch.softmax((x.float() @ self.w.float()), -1)
    i1 = gates.argmax(-1)
    m1 = torch.nn.functional.one_hot(i1, self.E).float()
    g2 = gates * (1 - m1)
    i2 = g2.argmax(-1)
    m2 = torch.nn.functional.one_hot(i2, self.E).float()
    aux = (m1.mean(1) * gates.mean(1)).mean() * self.E * self.E
    p1 = torch.cumsum(m1, 1) * m1 - m1
    m1 = m1 * (p1 < C)
    used = m1.sum(1, keepdim=True)
    p2 = (torch.cumsum(m2, 1) - m2 + used) * m2
    m2 = m2 * (p2 < C)
    w1, w2 = (gates * m1).sum(-1), (gates * m2).sum(-1)
    denom = (w1 + w2).clamp(min=1e-9)
    w1, w2 = w1 / denom, w2 / denom
    oh = lambda p: torch.nn.functional.one_hot(p.sum(-1).long().clamp(max=C - 1), C).float()
    combine = (w1[..., None, None] * m1.unsqueeze(-1) * oh(p1 * m1).unsqueeze(2) +
               w2[..., None, None] * m2.unsqueeze(-1) * oh(p2 * m2).unsqueeze(2))
    return (combine > 0).to(x.dtype), combine.to(x.dtype), aux


class MoEFFN(nn.Module):
  """Expert-parallel feed-forward layer.  Must be built inside ``with epl.split(n):`` (n = expert-parallel degree)."""
  epl_collective = True          # contains all-to-all: gradient checkpointing never recomputes it

  def __init__(self, d_model: int, d_ff: int, num_experts: int, capacity_factor: float = 1.25, gating: str = "top2", group=None):
    super().__init__()
    from easyparallellibrary_b200.ops import tensor_parallel as tp
    self.group = group or tp.current_tp_group()
    if num_experts % self.group.size:
      raise ValueError("num_experts must be divisible by the split size")
    self.E, self.E_local = num_experts, num_experts // self.group.size
    self.gate = (Top2Gating if gating == "top2" else SwitchGating)(d_model, num_experts, capacity_factor)
    self.gate.w.epl_tp_replicated = True
    # expert weights: dim 0 = experts, sharded over the split devices (add_weight semantics)
    self.wi = tp.add_weight((num_experts, d_model, d_ff), group=self.group, init=lambda t: nn.init.normal_(t, std=d_model ** -0.5))
    self.wo = tp.add_weight((num_experts, d_ff, d_model), group=self.group, init=lambda t: nn.init.normal_(t, std=d_ff ** -0.5))
    self.aux_loss = None

  def forward(self, x):                                   # [G, S, M]
    G, S, M = x.shape
    dispatch, combine, aux = self.gate(x)
    self.aux_loss = aux
    C = dispatch.shape[-1]
    routed = torch.einsum("gsec,gsm->egcm", dispatch, x)                 # einsum 1: dispatch
    n = self.group.size
    if n > 1:
      routed = expert_all_to_all(routed.reshape(n, self.E_local, G, C, M), self.group)       # [src, E_local, G, C, M]
      routed = routed.permute(1, 0, 2, 3, 4).reshape(self.E_local, n * G, C, M)
    h = torch.relu(torch.einsum("egcm,emh->egch", routed, self.wi.to(x.dtype)))                # einsum 2
    out = torch.einsum("egch,ehm->egcm", h, self.wo.to(x.dtype))                               # einsum 3 (expert side)
    if n > 1:
      out = out.reshape(self.E_local, n, G, C, M).permute(1, 0, 2, 3, 4).contiguous()
      out = expert_all_to_all(out, self.group).reshape(self.E, G, C, M)
    return torch.einsum("gsec,egcm->gsm", combine, out)                  # combine
