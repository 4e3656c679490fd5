"""Data-parallel hot path on B200: fused reduce-scatter + AdamW + all-gather over NVLink peer memory (K1).

One kernel launch per gradient bucket (``csrc/symm.cu: fused_rs_adam_ag_v2_kernel``) replaces
``ncclAllReduce`` + divide + unfused optimizer (reference ``graph_editor.py:670-725``,
``adam_weight_decay_optimizer.py:117-153``).  Optimizer state is an equal flat shard per rank, i.e.
ZeRO-v1 memory falls out for free; weights and gradients live in symmetric memory so the kernel
reads peers' gradient shards and writes peers' weight shards directly.

Overlap with backward (the reference overlaps its bucketed all-reduces with backward through its
communicator pool, ``communication_pool.py:84-105``): as soon as the last gradient of a bucket has been
produced the bucket's kernel is launched on a side stream with a handful of CTAs (TPC-aligned pairs).
It moves its bytes with bulk async copies through a shared-memory ring, so a few SMs sustain hundreds
of GB/s; the backward GEMMs keep running on the remaining SMs — their tile scheduler is an atomic
counter (``csrc/gemm_tcgen05.cu``), so CTA pairs that cannot be placed while the bucket kernel holds
SMs cost nothing.  Only the bucket whose gradients complete last (bucket 0: the embeddings) runs after
backward, on the whole GPU.
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
    from easyparallellibrary_b200.runtime.amp import DynamicLossScale
    return (trainer.device.type == "cuda" and not trainer.baseline and cfg.communication.fused_kernels
            and 1 < comm.size <= 8 and cfg.offload.level == ""
            # clipping: only the reference's default clip-then-reduce (local norm, applied to the bucket before the kernel runs)
            and (trainer.max_grad_norm is None or not cfg.communication.clip_after_allreduce)
            and cfg.zero.level in ("", "v0", "v1", "v2") and trainer.opt_kind in ("adam", "adamw")
            and cfg.optimizer.num_apply_group == 1
            and trainer.compute_dtype in (torch.bfloat16, torch.float16) and not isinstance(trainer.scaler, DynamicLossScale))

  def __init__(self, trainer):
    self.trainer = trainer
    self.lib = _sym_lib()
    self.pads: Dict[int, SignalPad] = {}
    self.local_sync: Dict[int, torch.Tensor] = {}
    self.dyn: Dict[int, torch.Tensor] = {}            # per group: device {lr, 1/(1-b1^t), 1/(1-b2^t), grad scale}
    self.epochs: Dict[Tuple[int, int], int] = {}      # v1 kernel only (v2 keeps its epoch on the device)
    self.side = torch.cuda.Stream(device=trainer.device, priority=-1)
    self.blocks = 148                                 # after backward: the whole GPU
    self.kernel = os.environ.get("EPL_K1", "v2")      # "v1": register-path kernel of round 1; "nvls": multimem.ld_reduce / multimem.st
    self.overlap_blocks = int(os.environ.get("EPL_FUSED_OVERLAP_BLOCKS", "8"))
    self.overlap = os.environ.get("EPL_FUSED_OVERLAP", "1") != "0" and self.kernel in ("v2", "nvls")
    # Each rank's AdamW shard is 1/W of the model, so the bucket kernels' work per rank shrinks with W while their cost to the
    # backward pass (8 SMs taken from every kernel that runs meanwhile + L2 traffic) does not.  Measured, GPT-2-XL, step ms
    # overlapped vs after backward: W=2 149.9 vs 117.5 (eager), W=4 122.6 vs 120.7 (CUDA graph; exposed 5.1 vs 10.5 ms but
    # forward+backward 7 ms slower).  At W=8 the per-rank work halves again and ~12 ms would be exposed otherwise: overlap
    # from W = 8 on (EPL_FUSED_OVERLAP_MIN_WORLD overrides).
    self.overlap_min_world = int(os.environ.get("EPL_FUSED_OVERLAP_MIN_WORLD", "8"))
    self.reserve_sms = os.environ.get("EPL_FUSED_RESERVE", "1") != "0"    # shrink the GEMM grids by the bucket kernel's SMs while it may run
    self.launched = set()
    self._prepared = False
    self._reserved = False

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
      self.local_sync[s] = torch.zeros(4 * len(flat.buckets), dtype=torch.int32, device=trainer.device)
      self.dyn[s] = torch.zeros(4, dtype=torch.float32, device=trainer.device)
    return self

  # ------------------------------------------------------------------ per step
  def begin_step(self, mean: bool) -> None:
    """Host -> device: the step's learning rate, bias corrections and gradient scale (everything else the kernels
    need is launch-invariant).  Called once per step before the first bucket can become ready."""
    tr = self.trainer
    for s in self.pads:
      comm = tr.dp_comms[s]
      opts = [o for b, o in zip(tr.flats[s].buckets, tr.optimizers[s]) if self.handles(s, b)]
      if not opts:
        continue
      for o in opts:
        o.step_count += 1
      h, t = opts[0].hyper, opts[0].step_count
      if h.bias_correction:
        inv_c1, inv_c2 = 1.0 / (1.0 - h.beta1 ** t), 1.0 / (1.0 - h.beta2 ** t)
      else:
        inv_c1 = inv_c2 = 1.0
      scale = tr.scaler.inv_scale / (tr.mean_divisor(s) if mean else 1)
      self.dyn[s].copy_(torch.tensor([h.lr, inv_c1, inv_c2, scale], dtype=torch.float32), non_blocking=True)
    self._prepared = True

  def handles(self, s: int, b) -> bool:
    """Buckets of 16-bit parameters live in symmetric memory and go through the fused kernel; fp32 buckets of the same group
    (e.g. BatchNorm parameters kept in fp32 under bf16 AMP) take the library path."""
    return s in self.pads and (s, "grad", b.dtype) in self.trainer._symm_buffers

  def reset_step(self) -> None:
    """Start of a step: forget per-step launch state (also after a step that was abandoned half-way, e.g. a failed capture)."""
    self.launched.clear()
    if self._reserved:
      from easyparallellibrary_b200.ops import linear as L
      L._NUM_SMS = 0
      self._reserved = False

  def launch_bucket_async(self, s: int, bi: int) -> None:
    """Called from the gradient hook when the last gradient of a bucket has been produced: the fused kernel runs on
    a side stream while backward continues.  Safe because its first action is a cross-GPU barrier: no rank's weights
    are overwritten before every rank has finished the backward of the layers in this bucket."""
    if (s, bi) in self.launched or s not in self.pads or bi == 0:       # bucket 0 completes last: whole GPU, after backward
      return
    if not self.handles(s, self.trainer.flats[s].buckets[bi]):
      return
    if self.trainer.dp_comms[s].size < self.overlap_min_world:
      return
    if not self._reserved and self.reserve_sms:
      # leave the bucket kernel's TPCs out of the GEMM grids until the step's reduce phase is over (the GEMM would cope —
      # its tile scheduler is dynamic — but CTA pairs that start late only to find no work left lengthen each GEMM's tail)
      from easyparallellibrary_b200.ops import linear as L
      L._NUM_SMS = 148 - (self.overlap_blocks + 1) // 2 * 2
      self._reserved = True
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream())
    self.side.wait_event(ev)
    with torch.cuda.stream(self.side):
      self.launch_bucket(s, bi, self.overlap_blocks)
    self.launched.add((s, bi))

  def launch_bucket(self, s: int, bi: int, blocks: int = 0) -> None:
    tr = self.trainer
    comm, flat = tr.dp_comms[s], tr.flats[s]
    b, opt = flat.buckets[bi], tr.optimizers[s][bi]
    gbuf, pbuf = tr._symm_buffers[(s, "grad", b.dtype)], tr._symm_buffers[(s, "param", b.dtype)]
    es = b.flat_grad.element_size()
    lo, hi = b.shard_range(comm.rank, comm.size)
    h = opt.hyper
    sync = self.local_sync[s][4 * bi:4 * bi + 4]
    if self.kernel == "v1":
      key = (s, bi)
      self.epochs[key] = self.epochs.get(key, 0) + 1
      d = self.dyn[s].tolist()                          # (host sync: measurement aid only)
      rc = self.lib.epl_fused_rs_adam_ag(
          gbuf.peer_table(b.start * es), pbuf.peer_table(b.start * es), self.pads[s].slot_table(bi), sync.data_ptr(),
          opt.master.data_ptr(), opt.m.data_ptr(), opt.v.data_ptr(), _lib.ptr(opt.decay_mask), lo, hi - lo, comm.rank,
          comm.size, self.epochs[key], _lib.dtype_code(b.dtype), d[0], h.beta1, h.beta2, h.eps, h.weight_decay, d[3],
          d[1], d[2], self.blocks, _lib.stream())
    elif self.kernel == "nvls" and getattr(gbuf, "multicast_ptr", 0) and b.dtype == torch.bfloat16:
      if not hasattr(self.lib, "_k1nvls_ready"):
        self.lib.epl_fused_nvls_adam.argtypes = ([ctypes.c_void_p] * 8 + [ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_uint,
                                                  ctypes.c_void_p] + [ctypes.c_float] * 4 + [ctypes.c_int, ctypes.c_void_p])
        self.lib._k1nvls_ready = True
      rc = self.lib.epl_fused_nvls_adam(
          gbuf.multicast_ptr + b.start * es, pbuf.multicast_ptr + b.start * es, self.pads[s].slot_table(bi), sync.data_ptr(),
          opt.master.data_ptr(), opt.m.data_ptr(), opt.v.data_ptr(), _lib.ptr(opt.decay_mask), lo, hi - lo, comm.rank, comm.size, 0,
          self.dyn[s].data_ptr(), h.beta1, h.beta2, h.eps, h.weight_decay, (blocks * 4) if blocks else 0, _lib.stream())
    else:
      rc = self.lib.epl_fused_rs_adam_ag_v2(
          gbuf.peer_table(b.start * es), pbuf.peer_table(b.start * es), self.pads[s].slot_table(bi), sync.data_ptr(),
          opt.master.data_ptr(), opt.m.data_ptr(), opt.v.data_ptr(), _lib.ptr(opt.decay_mask), lo, hi - lo, comm.rank,
          comm.size, 0, _lib.dtype_code(b.dtype), self.dyn[s].data_ptr(), h.beta1, h.beta2, h.eps, h.weight_decay,
          blocks or self.blocks, _lib.stream())
    _lib.check(rc, "fused_rs_adam_ag")

  def reduce_and_apply(self, mean: bool):
    tr = self.trainer
    if not self._prepared:
      self.begin_step(mean)
    if self._reserved:
      from easyparallellibrary_b200.ops import linear as L
      L._NUM_SMS = 0
      self._reserved = False
    if self.launched:                                   # finish the overlapped buckets first: the tail bucket gets the whole GPU
      torch.cuda.current_stream().wait_stream(self.side)
    for s in tr.group_keys:
      if s not in self.pads:
        tr._apply_group_library(s, mean)
        continue
      library = []
      for bi in range(len(tr.flats[s].buckets) - 1, -1, -1):
        if not self.handles(s, tr.flats[s].buckets[bi]):
          library.append(bi)
        elif (s, bi) not in self.launched:
          self.launch_bucket(s, bi)
      if library:
        tr._apply_group_library(s, mean, only=library)
    self.launched.clear()
    self._prepared = False
    return False, None
