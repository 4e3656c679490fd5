"""Pipeline executor: interprets a stage program (``parallel/schedule.py``) on one rank.

Reference mechanics (``graph_editor.py:397-421`` micro-batch clones + ``scheduler.py`` control
edges + implicit TF ``_Send/_Recv``) become: a loop over instructions, micro-batch activations kept
in a dict, and explicit NCCL point-to-point transfers.  Transfers in the two directions use two
independent process groups (hence independent NCCL streams), receives are posted ``prefetch``
slots early into per-micro-batch buffers (double buffering) and sends are asynchronous, so under
1F1B a transfer overlaps the neighbouring micro-batch's compute instead of sitting on the
critical path.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from easyparallellibrary_b200.ir.graph import Graph
from easyparallellibrary_b200.ir.phase import ModelPhase, phase_scope
from easyparallellibrary_b200.parallel import schedule as S

_META_LEN = 10
_DTYPES = [torch.float32, torch.bfloat16, torch.float16, torch.int64, torch.int32]


class PipelineExecutor(object):
  def __init__(self, trainer):
    self.tr = trainer
    plan = trainer.plan
    if len(plan.local_stages) != 1:
      raise RuntimeError("pipeline execution expects exactly one stage per rank (got %s)" % plan.local_stages)
    self.stage = plan.local_stages[0]
    self.num_stages = plan.num_stages
    self.M = plan.num_micro_batch
    self.first = self.stage == 0
    self.last = self.stage == self.num_stages - 1
    self.prev = plan.prev_rank(self.stage)
    self.next = plan.next_rank(self.stage)
    self.module = trainer.stage_modules[self.stage]
    policy = trainer.config.pipeline.strategy
    self.program = S.build_stage_program(policy, self.stage, self.num_stages, self.M, prefetch=1)
    try:      # the native runtime generates the same program; use it when built (cross-checked in the tests)
      from easyparallellibrary_b200.runtime import native
      if native.available():
        self.program = [S.Instr(op, mb) for op, mb in native.schedule_stage(policy, self.stage, self.num_stages, self.M, 1)]
    except Exception:  # pragma: no cover
      pass
    # two process groups -> activations and activation-gradients travel on independent NCCL streams
    self.pg_fwd = dist.new_group()
    self.pg_bwd = dist.new_group()
    # one group per model replica (all stages of that replica) for cross-stage reductions
    self.replica_group = None
    for r in range(plan.num_replicas):
      ranks = [plan.stage_ranks[s][r][0] for s in range(self.num_stages)]
      g = dist.new_group(ranks)
      if plan.rank in ranks:
        self.replica_group = g
        self.replica_ranks = ranks
    self._shape_fwd: Optional[Tuple[torch.Size, torch.dtype]] = None
    self.device = trainer.device

  # ------------------------------------------------------------------ metadata handshake (first step only)
  def _send_meta(self, t: torch.Tensor, dst: int, group) -> None:
    meta = torch.zeros(_META_LEN, dtype=torch.int64, device=self.device)
    meta[0] = t.dim()
    meta[1] = _DTYPES.index(t.dtype)
    for i, s in enumerate(t.shape):
      meta[2 + i] = s
    dist.send(meta, dst, group=group)

  def _recv_meta(self, src: int, group) -> Tuple[torch.Size, torch.dtype]:
    meta = torch.zeros(_META_LEN, dtype=torch.int64, device=self.device)
    dist.recv(meta, src, group=group)
    meta = meta.cpu().tolist()
    return torch.Size(meta[2:2 + meta[0]]), _DTYPES[meta[1]]

  # ------------------------------------------------------------------ one training step
  def run(self, micro: List[Tuple[Any, ...]], mean: bool):
    tr, graph = self.tr, Graph.get()
    inputs: Dict[int, torch.Tensor] = {}
    outputs: Dict[int, torch.Tensor] = {}
    recv_f: Dict[int, Tuple[torch.Tensor, Any]] = {}
    recv_b: Dict[int, Tuple[torch.Tensor, Any]] = {}
    sends: List[Any] = []
    losses: List[torch.Tensor] = []
    collected = []
    tr._first_micro_batch, tr._last_micro_batch = True, False
    n_back = 0
    debug = bool(int(__import__("os").environ.get("EPL_PIPE_DEBUG", "0")))
    for ins in self.program:
      if debug:
        print("[pipe rank %d stage %d] %r" % (tr.plan.rank, self.stage, ins), flush=True)
      if ins.op == S.RECV_F:
        if self._shape_fwd is None:
          self._shape_fwd = self._recv_meta(self.prev, self.pg_fwd)
        buf = torch.empty(self._shape_fwd[0], dtype=self._shape_fwd[1], device=self.device)
        recv_f[ins.mb] = (buf, dist.irecv(buf, self.prev, group=self.pg_fwd))
      elif ins.op == S.RECV_B:
        out = outputs[ins.mb]
        buf = torch.empty_like(out)
        recv_b[ins.mb] = (buf, dist.irecv(buf, self.next, group=self.pg_bwd))
      elif ins.op == S.F:
        mb = micro[ins.mb]
        if self.first:
          x = mb[0]
        else:
          x, req = recv_f.pop(ins.mb)
          req.wait()
          if x.is_floating_point():
            x.requires_grad_()
        with phase_scope(ModelPhase.FORWARD):
          y = self.module(x)
          if self.last:
            loss = tr.loss_fn(y, *mb[1:]) if tr.loss_fn is not None else y
            losses.append(loss.detach())
            y = tr.scaler.scale(loss)
            if mean and self.M > 1:
              y = y / self.M
        collected.append(graph.pop_collections())
        inputs[ins.mb], outputs[ins.mb] = x, y
      elif ins.op == S.SEND_F:
        y = outputs[ins.mb]
        if not getattr(self, "_meta_sent", False):
          self._send_meta(y, self.next, self.pg_fwd)
          self._meta_sent = True
        sends.append(dist.isend(y.detach(), self.next, group=self.pg_fwd))
      elif ins.op == S.B:
        n_back += 1
        tr._last_micro_batch = n_back == self.M
        y = outputs.pop(ins.mb)
        with phase_scope(ModelPhase.BACKWARD):
          if self.last:
            y.backward()
          else:
            g, req = recv_b.pop(ins.mb)
            req.wait()
            torch.autograd.backward(y, grad_tensors=g)
        tr._first_micro_batch = False
      elif ins.op == S.SEND_B:
        x = inputs.pop(ins.mb)
        sends.append(dist.isend(x.grad, self.prev, group=self.pg_bwd))
      # REDUCE / APPLY are executed by the trainer after the program
    for w in sends:
      w.wait()
    inputs.clear()
    # the loss lives on the last stage; share its mean with every stage of the replica
    stat = torch.zeros(1, device=self.device, dtype=torch.float32)
    if self.last and losses:
      stat[0] = torch.stack([l.float() for l in losses]).mean() if mean else torch.stack([l.float() for l in losses]).sum()
    dist.all_reduce(stat, group=self.replica_group)
    return [stat[0]], collected

  @torch.no_grad()
  def forward_only(self, batch: Tuple[Any, ...]):
    """Evaluation: stages run back to back, no micro-batching."""
    if self.first:
      x = batch[0]
    else:
      if self._shape_fwd is None:
        self._shape_fwd = self._recv_meta(self.prev, self.pg_fwd)
      x = torch.empty((batch[0].shape[0],) + tuple(self._shape_fwd[0][1:]), dtype=self._shape_fwd[1], device=self.device)
      dist.recv(x, self.prev, group=self.pg_fwd)
    y = self.module(x)
    if not self.last:
      if not getattr(self, "_meta_sent", False):
        self._send_meta(y, self.next, self.pg_fwd)
        self._meta_sent = True
      dist.send(y, self.next, group=self.pg_fwd)
      return None
    return self.tr.loss_fn(y, *batch[1:]) if self.tr.loss_fn is not None else y

  def all_reduce_over_stages(self, t: torch.Tensor) -> torch.Tensor:
    dist.all_reduce(t, group=self.replica_group)
    return t
