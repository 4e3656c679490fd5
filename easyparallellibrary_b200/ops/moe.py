"""Mixture-of-experts building blocks with expert parallelism over ``epl.split``.

Reference: ``examples/moe/moe_ffn.py`` — ``Top2Gating`` (127-269) / ``SwitchGating`` (270-369) with capacity factor
and auxiliary load-balancing loss, ``MoEFFN`` (370-464) whose expert weights are created under ``epl.split``
(dim 0 = experts, ``hooks.py:667-707``) and whose three einsums trigger an all-to-all *dispatch* before the first
and an all-to-all *combine* before the third (``hooks.py:758-794``).

Here the layer is explicit and index based: gating produces, per token, the (expert, slot) of each of its choices — the
reference's dense one-hot dispatch/combine tensors ``[G,S,E,C]`` and their ``O(S*E*C*M)`` einsums are never built; dispatch
is a row gather, combine a weighted row gather (``O(S*k*M)``).  On B200:

* **dispatch + all-to-all are one kernel** (``csrc/symm.cu: alltoall_gather_p2p_kernel``, K5b): every (expert, slot) row is
  read from the local token matrix and stored straight into the owning rank's symmetric receive buffer over NVLink, already
  in the ``[local expert, source rank x slots, M]`` layout the expert GEMMs consume;
* **expert FFNs run on the tcgen05 GEMM** (one launch per local expert and projection, all three of forward / dX / dW
  without transposes);
* the combine all-to-all is the peer-store kernel ``alltoall_p2p_kernel`` (K5).

NCCL all-to-all + einsum experts is the CPU / baseline path (``USE_P2P_KERNEL = False``).  The all-to-all is never
recomputed by gradient checkpointing (``epl_collective``), like the reference (``constant.py:97``).
"""
from __future__ import annotations

import ctypes
from typing import Dict

import torch
from torch import nn

from easyparallellibrary_b200.communicators import functional as CF

# peer-store all-to-all kernels (K5 / K5b) on GPUs; EPL_MOE_P2P=0 selects NCCL (the baseline arm of tools/mgpu_check.py moe)
import os as _os
USE_P2P_KERNEL = _os.environ.get("EPL_MOE_P2P", "1") == "1"
USE_EXPERT_GEMM = _os.environ.get("EPL_MOE_GEMM", "1") == "1"
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
  rc = lib.epl_alltoall_p2p(t.data_ptr(), buf.peer_table(0), ws.pad.slot_table(0), ws.sync.data_ptr(), seg, group.rank, group.size,
                            0, 64, _lib.stream())                     # epoch 0: the kernels keep it on the device (graph-replayable)
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
# fused dispatch + all-to-all (K5b)
# --------------------------------------------------------------------------------------------------------
def _p2p_dispatch(x2: torch.Tensor, index: torch.Tensor, E: int, slots: int, group) -> torch.Tensor:
  """``x2`` [tokens, M], ``index`` [E * slots] int32 (row of x2 or -1) -> [E_local, world * slots, M] on every rank."""
  from easyparallellibrary_b200.ops import _lib
  from easyparallellibrary_b200.runtime.symmetric import _sym_lib
  lib = _sym_lib()
  if not hasattr(lib, "_a2ag_ready"):
    lib.epl_alltoall_gather_p2p.argtypes = ([ctypes.c_void_p] * 5 + [ctypes.c_int64] + [ctypes.c_int] * 4 + [ctypes.c_uint, ctypes.c_int,
                                                                                                    ctypes.c_void_p])
    lib._a2ag_ready = True
  ws = _A2A_WS.get(id(group))
  if ws is None:
    ws = _A2A_WS[id(group)] = _A2AWorkspace(group, x2.device)
  M = x2.shape[1]
  e_local = E // group.size
  nbytes = group.size * e_local * slots * M * x2.element_size()
  buf = ws.recv_buffer(nbytes)
  rc = lib.epl_alltoall_gather_p2p(x2.data_ptr(), index.data_ptr(), buf.peer_table(0), ws.pad.slot_table(0), ws.sync.data_ptr(),
                                   M * x2.element_size(), E, slots, group.rank, group.size, 0, 96, _lib.stream())
  _lib.check(rc, "alltoall_gather_p2p")
  return buf.tensor(x2.dtype, group.size * e_local * slots * M).view(e_local, group.size * slots, M).clone()


class _Dispatch(torch.autograd.Function):
  """tokens [T, M] + slot table -> expert inputs [E_local, world * slots, M] (gather + all-to-all); the adjoint is the
  all-to-all back + a scatter-add of the slot gradients onto their tokens."""

  @staticmethod
  def forward(ctx, x2, index, E, slots, group):
    ctx.group, ctx.E, ctx.slots, ctx.T = group, E, slots, x2.shape[0]
    ctx.save_for_backward(index)
    n, e_local, M = group.size, E // group.size, x2.shape[1]
    if (USE_P2P_KERNEL and x2.is_cuda and n > 1 and n <= 8 and (M * x2.element_size()) % 16 == 0 and x2.is_contiguous()
        and x2.dtype in (torch.bfloat16, torch.float16)):
      return _p2p_dispatch(x2, index, E, slots, group)
    pad = torch.cat([x2, x2.new_zeros(1, M)], 0)
    routed = pad[index.long()]                                           # [E * slots, M]; index -1 -> the zero row
    if n > 1:
      routed = expert_all_to_all_raw(routed.view(n, e_local * slots * M), group).view(n, e_local, slots, M)
      routed = routed.permute(1, 0, 2, 3).reshape(e_local, n * slots, M)
    else:
      routed = routed.view(E, slots, M)
    return routed

  @staticmethod
  def backward(ctx, g):
    (index,) = ctx.saved_tensors
    group, E, slots = ctx.group, ctx.E, ctx.slots
    n, e_local, M = group.size, E // group.size, g.shape[-1]
    g = g.contiguous()
    if n > 1:
      g = g.view(e_local, n, slots, M).permute(1, 0, 2, 3).contiguous()
      g = expert_all_to_all_raw(g.view(n, -1), group)
    g = g.reshape(E * slots, M)
    dx = g.new_zeros(ctx.T + 1, M)
    dx.index_add_(0, torch.where(index < 0, torch.full_like(index, ctx.T), index).long(), g)
    return dx[:ctx.T], None, None, None, None


# --------------------------------------------------------------------------------------------------------
# expert GEMMs on the tcgen05 kernel
# --------------------------------------------------------------------------------------------------------
class _ExpertGemm(torch.autograd.Function):
  """``y[e] = x[e] @ w[e]`` for the local experts: x [E, T, K], w [E, K, N].  One tcgen05 GEMM launch per expert; forward
  (B MN-major), dX (B K-major) and dW (both operands MN-major) all read the tensors as they are."""

  @staticmethod
  def forward(ctx, x, w):
    from easyparallellibrary_b200.ops.linear import gemm
    ctx.save_for_backward(x, w)
    y = x.new_empty(x.shape[0], x.shape[1], w.shape[2])
    for e in range(x.shape[0]):
      gemm(x[e], w[e], b_mn_major=True, out=y[e])
    return y

  @staticmethod
  def backward(ctx, dy):
    from easyparallellibrary_b200.ops.linear import gemm
    x, w = ctx.saved_tensors
    dy = dy.contiguous()
    dx, dw = torch.empty_like(x), torch.empty_like(w)
    for e in range(x.shape[0]):
      gemm(dy[e], w[e], out=dx[e])                                       # [T, N] x ([K, N] read as N_out=K, contraction N)
      gemm(x[e], dy[e], a_mn_major=True, b_mn_major=True, out=dw[e])     # x^T dy -> [K, N]
    return dx, dw


def expert_matmul(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
  """[E, T, K] x [E, K, N] -> [E, T, N]."""
  from easyparallellibrary_b200.ops import _lib
  ok = (USE_EXPERT_GEMM and x.is_cuda and _lib.available() and x.dtype in (torch.bfloat16, torch.float16) and w.dtype == x.dtype
        and x.shape[2] % 8 == 0 and w.shape[2] % 8 == 0 and x.is_contiguous() and w.is_contiguous() and x.shape[1] >= 8)
  if ok:
    return _ExpertGemm.apply(x, w)
  return torch.einsum("etk,ekn->etn", x, w.to(x.dtype))


# --------------------------------------------------------------------------------------------------------
# gating
# --------------------------------------------------------------------------------------------------------
def _capacity(tokens: int, experts: int, factor: float, minimum: int = 4) -> int:
  return max(int(tokens * factor / experts), minimum)


class _Gating(nn.Module):
  """Shared routing logic.  ``route(x)`` -> ``(slots [G,S,k] int64 in [0, E*C) or -1, weights [G,S,k], C, aux)``: choice j of
  token s goes to expert ``slot // C`` at position ``slot % C`` (capacity overflow: -1, weight 0).  ``forward`` builds the
  reference's dense ``(dispatch, combine, aux)`` triple from it (API parity with examples/moe/moe_ffn.py; not used by MoEFFN)."""
  k = 1

  def __init__(self, d_model: int, num_experts: int, capacity_factor: float = 1.25):
    super().__init__()
    self.w = nn.Parameter(torch.empty(d_model, num_experts))
    nn.init.normal_(self.w, std=d_model ** -0.5)
    self.E, self.cf = num_experts, capacity_factor

  def route(self, x):
    raise NotImplementedError

  def forward(self, x):
    slots, weights, C, aux = self.route(x)
    G, S, _ = x.shape
    combine = x.new_zeros(G, S, self.E * C, dtype=torch.float32)
    for j in range(slots.shape[-1]):
      sj, wj = slots[..., j], weights[..., j]
      combine.scatter_add_(2, sj.clamp(min=0).unsqueeze(-1), (wj * (sj >= 0)).unsqueeze(-1).float())
    combine = combine.view(G, S, self.E, C)
    return (combine > 0).to(x.dtype), combine.to(x.dtype), aux


class SwitchGating(_Gating):
  """Top-1 routing with capacity and the load-balancing auxiliary loss."""
  k = 1

  def route(self, x):                                     # [G, S, M]
    G, S, _ = x.shape
    C = _capacity(S, self.E, self.cf)
    gates = torch.softmax((x.float() @ self.w.float()), -1)          # [G,S,E]
    idx = gates.argmax(-1)
    mask = torch.nn.functional.one_hot(idx, self.E).float()
    aux = (mask.mean(1) * gates.mean(1)).mean() * self.E * self.E
    pos = ((torch.cumsum(mask, 1) - 1) * mask).sum(-1).long()         # position of each token inside its expert
    keep = pos < C
    gate = gates.gather(-1, idx.unsqueeze(-1)).squeeze(-1)
    slots = torch.where(keep, idx * C + pos, torch.full_like(idx, -1)).unsqueeze(-1)
    return slots, (gate * keep).unsqueeze(-1), C, aux


class Top2Gating(_Gating):
  """Top-2 routing with capacity (second choices are placed after all first choices)."""
  k = 2

  def route(self, x):
    G, S, _ = x.shape
    C = _capacity(2 * S, self.E, self.cf)
    gates = torch.softmax((x.float() @ self.w.float()), -1)
    i1 = gates.argmax(-1)
    m1 = torch.nn.functional.one_hot(i1, self.E).float()
    i2 = (gates * (1 - m1)).argmax(-1)
    m2 = torch.nn.functional.one_hot(i2, self.E).float()
    aux = (m1.mean(1) * gates.mean(1)).mean() * self.E * self.E
    p1 = ((torch.cumsum(m1, 1) - 1) * m1).sum(-1).long()
    k1 = p1 < C
    used = (m1 * k1.unsqueeze(-1)).sum(1, keepdim=True)               # first choices actually placed per expert
    p2 = ((torch.cumsum(m2, 1) - 1 + used) * m2).sum(-1).long()
    k2 = p2 < C
    w1 = gates.gather(-1, i1.unsqueeze(-1)).squeeze(-1) * k1
    w2 = gates.gather(-1, i2.unsqueeze(-1)).squeeze(-1) * k2
    denom = (w1 + w2).clamp(min=1e-9)
    s1 = torch.where(k1, i1 * C + p1, torch.full_like(i1, -1))
    s2 = torch.where(k2, i2 * C + p2, torch.full_like(i2, -1))
    return torch.stack([s1, s2], -1), torch.stack([w1 / denom, w2 / denom], -1), C, aux


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
    slots, weights, C, aux = self.gate.route(x)            # [G,S,k]
    self.aux_loss = aux
    n, E = self.group.size, self.E
    k = slots.shape[-1]
    # slot table: index[e, g, c] = row of the flattened token matrix routed there (or -1); slot = e*C + c inside group g
    g_ids = torch.arange(G, device=x.device).view(G, 1, 1).expand(G, S, k)
    rows = (g_ids * S + torch.arange(S, device=x.device).view(1, S, 1)).reshape(-1)
    flat = slots.reshape(-1)
    e_idx, c_idx = torch.div(flat.clamp(min=0), C, rounding_mode="floor"), flat.clamp(min=0) % C
    dest = (e_idx * G + g_ids.reshape(-1)) * C + c_idx                 # position in [E, G, C]
    index = torch.full((E * G * C + 1,), -1, dtype=torch.int32, device=x.device)
    index.scatter_(0, torch.where(flat >= 0, dest, torch.full_like(dest, E * G * C)), rows.to(torch.int32))
    index = index[:E * G * C].contiguous()
    x2 = x.reshape(G * S, M)
    routed = _Dispatch.apply(x2 if x2.is_contiguous() else x2.contiguous(), index, E, G * C, self.group)    # [E_local, n*G*C, M]
    h = torch.relu(expert_matmul(routed, self.wi.to(x.dtype)))
    out = expert_matmul(h, self.wo.to(x.dtype))                        # [E_local, n*G*C, M]
    if n > 1:
      out = out.view(self.E_local, n, G * C * M)
      out = out.permute(1, 0, 2).contiguous() if self.E_local > 1 else out.view(n, G * C * M)
      out = expert_all_to_all(out, self.group)                         # [src rank, E_local * G*C*M] = experts in global order
    out = out.reshape(E * G * C, M)
    # combine: y[token] = sum_j weight_j * out[slot_j]
    pad = torch.cat([out, out.new_zeros(1, M)], 0)
    gathered = pad[torch.where(flat >= 0, dest, torch.full_like(dest, E * G * C))].view(G, S, k, M)
    return (gathered * weights.to(x.dtype).unsqueeze(-1)).sum(2)
