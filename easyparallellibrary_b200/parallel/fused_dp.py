"""Data-parallel hot path on B200: fused reduce-scatter + AdamW + all-gather over NVLink peer memory.

One kernel launch per gradient bucket (``csrc/symm.cu: fused_rs_adam_ag_kernel``) replaces
``ncclAllReduce`` + divide + unfused optimizer (reference ``graph_editor.py:670-725``,
``adam_weight_decay_optimizer.py:117-153``).  Optimizer state is an equal flat shard per rank, i.e.
ZeRO-v1 memory falls out for free; weights and gradients live in symmetric memory so the kernel
reads peers' gradient shards and writes peers' weight shards directly.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional, Tuple

import torch

from easyparallellibrary_b200.ops import _lib
from easyparallellibrary_b200.runtime.symmetric import SignalPad, _sym_lib


class FusedDataParallel(object):
  @staticmethod
  def eligible(trainer, comm) -> bool:
    cfg = trainer.config
    return (trainer.device.type == "cuda" and not trainer.baseline and cfg.communication.fused_kernels
            and 1 < comm.size <= 8 and cfg.offload.level == ""
            # clipping: only the reference's default clip-then-reduce (local norm, applied to the bucket before the kernel runs)
            and (trainer.max_grad_norm is None or not cfg.communication.clip_after_allreduce)
            and cfg.zero.level in ("", "v0", "v1", "v2") and trainer.opt_kind in ("adam", "adamw")
            and cfg.optimizer.num_apply_group == 1
            and trainer.compute_dtype in (torch.bfloat16, torch.float16) and not isinstance(
                trainer.scaler, __import__("easyparallellibrary_b200.runtime.amp", fromlist=["DynamicLossScale"]).DynamicLossScale))

  def __init__(self, trainer):
    self.trainer = trainer
    self.lib = _sym_lib()
    self.pads: Dict[int, SignalPad] = {}
    self.local_sync: Dict[int, torch.Tensor] = {}
    self.epochs: Dict[Tuple[int, int], int] = {}
    self.side = torch.cuda.Stream(device=trainer.device, priority=-1)
    self.blocks = 148                 # after backward: the whole GPU
    # During backward the bucket kernel can run on a side stream with a few CTAs.  It needs most of an SM's registers, so its
    # CTAs displace CTAs of the persistent GEMMs rather than sharing SMs with them; EPL_FUSED_OVERLAP=0 runs every bucket
    # after backward on the whole GPU instead (no interference, fully exposed), EPL_FUSED_OVERLAP_BLOCKS sizes the overlap.
    self.overlap_blocks = int(os.environ.get("EPL_FUSED_OVERLAP_BLOCKS", "32"))
    self.launched = set()
    self.overlap = os.environ.get("EPL_FUSED_OVERLAP", "0") != "0"   # measured on 2 x B200: 117.7 ms/step off vs 122.8 ms on (GPT-2-XL)

  @classmethod
  def maybe_create(cls, trainer) -> Optional["FusedDataParallel"]:
    if not getattr(trainer, "_symm_buffers", None):
      return None
    self = cls(trainer)
    for s in trainer.group_keys:
      comm = trainer.dp_comms[s]
      flat = trainer.flats[s]
      if comm.size <= 1 or (s, "grad", torch.bfloat16) not in trainer._symm_buffers and (s, "grad", torch.float16) not in trainer._symm_buffers:
        continue
      self.pads[s] = SignalPad(len(flat.buckets), comm.ranks, trainer.device, group=getattr(comm.primary, "group", None))
      self.local_sync[s] = torch.zeros(2 * len(flat.buckets), dtype=torch.int32, device=trainer.device)
    return self

  def launch_bucket_async(self, s: int, bi: int, mean: bool) -> None:
    """Called from the gradient hook when the last gradient of a bucket has been produced: the fused kernel runs on
    a side stream while backward continues.  Safe because its first action is a cross-GPU barrier: no rank's weights
    are overwritten before every rank has finished the backward of the layers in this bucket."""
    if (s, bi) in self.launched or s not in self.pads:
      return
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream())
    self.side.wait_event(ev)
    with torch.cuda.stream(self.side):
      self.launch_bucket(s, bi, mean, self.overlap_blocks)
    self.launched.add((s, bi))

  def launch_bucket(self, s: int, bi: int, mean: bool, blocks: int = 0) -> None:
    tr = self.trainer
    comm, flat = tr.dp_comms[s], tr.flats[s]
    b, opt = flat.buckets[bi], tr.optimizers[s][bi]
    gbuf, pbuf = tr._symm_buffers[(s, "grad", b.dtype)], tr._symm_buffers[(s, "param", b.dtype)]
    es = b.flat_grad.element_size()
    lo, hi = b.shard_range(comm.rank, comm.size)
    key = (s, bi)
    self.epochs[key] = self.epochs.get(key, 0) + 1
    opt.step_count += 1
    h = opt.hyper
    if h.bias_correction:
      inv_c1, inv_c2 = 1.0 / (1.0 - h.beta1 ** opt.step_count), 1.0 / (1.0 - h.beta2 ** opt.step_count)
    else:
      inv_c1 = inv_c2 = 1.0
    scale = tr.scaler.inv_scale / (comm.size if mean else 1)
    sync = self.local_sync[s][2 * bi:2 * bi + 2]
    rc = self.lib.epl_fused_rs_adam_ag(
        gbuf.peer_table(b.start * es), pbuf.peer_table(b.start * es), self.pads[s].slot_table(bi), sync.data_ptr(),
        opt.master.data_ptr(), opt.m.data_ptr(), opt.v.data_ptr(), _lib.ptr(opt.decay_mask), lo, hi - lo, comm.rank,
        comm.size, self.epochs[key], _lib.dtype_code(b.dtype), h.lr, h.beta1, h.beta2, h.eps, h.weight_decay, scale,
        inv_c1, inv_c2, blocks or self.blocks, _lib.stream())
    _lib.check(rc, "fused_rs_adam_ag")

  def reduce_and_apply(self, mean: bool):
    tr = self.trainer
    for s in tr.group_keys:
      if s not in self.pads:
        tr._apply_group_library(s, mean)
        continue
      for bi in range(len(tr.flats[s].buckets) - 1, -1, -1):
        if (s, bi) not in self.launched:
          self.launch_bucket(s, bi, mean and not tr.has_split)
    if self.launched:
      torch.cuda.current_stream().wait_stream(self.side)
      self.launched.clear()
    return False, None
