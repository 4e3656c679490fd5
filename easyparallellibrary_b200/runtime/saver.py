"""Checkpoint utilities.

Reference (``epl/runtime/saver.py``, ``hooks.py:531-590``): only the first constructor rank writes (every
worker writes its own shard when a ``split`` taskgraph exists); restore on the first constructor then
broadcast; ``ShardingLoader`` (46-128) warm-starts from another checkpoint with a name map and per-tensor
slices; ``MemoryEfficientBuilder`` (145-207) saves in bounded-memory buckets (>= 50 MB, sequential, via CPU).
"""
from __future__ import annotations

import json
import os
import re
from typing import Any, Dict, List, Optional, Tuple

import torch

BUCKET_BYTES = 50 << 20


def _rank() -> int:
  import torch.distributed as dist
  return dist.get_rank() if dist.is_initialized() else 0


class MemoryEfficientBuilder(object):
  """Sharded save: tensors are staged on the host bucket by bucket (each >= ``bucket_bytes``) and written one
  file per bucket, so peak extra host memory is one bucket, never the whole model."""

  def __init__(self, directory: str, bucket_bytes: int = BUCKET_BYTES):
    self.dir, self.bucket_bytes = directory, bucket_bytes

  def save(self, state: Dict[str, torch.Tensor], prefix: str = "model", extra: Optional[Dict[str, Any]] = None) -> List[str]:
    os.makedirs(self.dir, exist_ok=True)
    index, files, cur, cur_bytes = {}, [], {}, 0

    def flush():
      nonlocal cur, cur_bytes
      if not cur:
        return
      name = "%s-%05d.pt" % (prefix, len(files))
      torch.save(cur, os.path.join(self.dir, name))
      files.append(name)
      cur, cur_bytes = {}, 0

    for k, v in state.items():
      t = v.detach().to("cpu") if isinstance(v, torch.Tensor) else v
      cur[k] = t
      index[k] = len(files)
      cur_bytes += t.numel() * t.element_size() if isinstance(t, torch.Tensor) else 0
      if cur_bytes >= self.bucket_bytes:
        flush()
    flush()
    with open(os.path.join(self.dir, prefix + ".index.json"), "w") as f:
      json.dump({"files": files, "index": {k: files[i] for k, i in index.items()}, "extra": extra or {}}, f)
    return files

  def load(self, prefix: str = "model") -> Tuple[Dict[str, torch.Tensor], Dict[str, Any]]:
    meta = json.load(open(os.path.join(self.dir, prefix + ".index.json")))
    out = {}
    for name in meta["files"]:
      out.update(torch.load(os.path.join(self.dir, name), map_location="cpu"))
    return out, meta.get("extra", {})


class ShardingLoader(object):
  """Initialise a model from a checkpoint of a *differently shaped* run.

  ``assign_map``: ``{checkpoint_name_regex: model_name_template}`` (``\\1`` style groups) or a callable;
  ``sharding_info``: ``{model_name: (begin, size)}`` per dimension slices — e.g. load this rank's slice of an
  unsharded tensor into a tensor-parallel shard.  Parameters tagged ``epl_tp_shard`` get their slice automatically.
  """

  def __init__(self, checkpoint: Dict[str, torch.Tensor], assign_map=None, sharding_info: Optional[Dict[str, Tuple]] = None):
    self.ckpt, self.assign_map, self.sharding_info = checkpoint, assign_map, sharding_info or {}

  def _target_name(self, ckpt_name: str) -> Optional[str]:
    if self.assign_map is None:
      return ckpt_name
    if callable(self.assign_map):
      return self.assign_map(ckpt_name)
    for pat, tmpl in self.assign_map.items():
      if re.fullmatch(pat, ckpt_name):
        return re.sub(pat, tmpl, ckpt_name)
    return None

  def load_into(self, module: torch.nn.Module, strict: bool = False) -> List[str]:
    params = dict(module.named_parameters())
    params.update(dict(module.named_buffers()))
    loaded = []
    for ck, tensor in self.ckpt.items():
      name = self._target_name(ck)
      if name is None or name not in params:
        continue
      dst = params[name]
      src = tensor
      info = self.sharding_info.get(name)
      if info is None and hasattr(dst, "epl_tp_shard") and tuple(src.shape) != tuple(dst.shape):
        dim, lo, hi, _total = dst.epl_tp_shard
        info = tuple((lo, hi - lo) if d == dim else (0, src.shape[d]) for d in range(src.dim()))
      if info is not None:
        if isinstance(info[0], int):
          info = (info,)
        for d, (begin, size) in enumerate(info):
          src = src.narrow(d, begin, size)
      if tuple(src.shape) != tuple(dst.shape):
        if strict:
          raise ValueError("shape mismatch for %s: checkpoint %s vs model %s" % (name, tuple(src.shape), tuple(dst.shape)))
        continue
      with torch.no_grad():
        dst.copy_(src.to(dst.dtype))
      loaded.append(name)
    if strict:
      missing = set(params) - set(loaded)
      if missing:
        raise KeyError("checkpoint misses %s" % sorted(missing))
    return loaded


def _optimizer_state(trainer) -> Dict[str, torch.Tensor]:
  opt: Dict[str, torch.Tensor] = {}
  for s in trainer.group_keys:
    for i, o in enumerate(trainer.optimizers[s]):
      for k, v in o.state_dict().items():
        if isinstance(v, torch.Tensor):
          opt["g%d.b%d.%s" % (s, i, k)] = v
        elif isinstance(v, (int, float)):
          opt["g%d.b%d.%s" % (s, i, k)] = torch.tensor(v)
  for s, z in getattr(trainer, "zero3", {}).items():
    for i, sd in enumerate(z.state_dict()):
      for k, v in sd.items():
        if isinstance(v, torch.Tensor):
          opt["z%d.u%d.%s" % (s, i, k)] = v
        elif isinstance(v, (int, float)):
          opt["z%d.u%d.%s" % (s, i, k)] = torch.tensor(v)
  return opt


def _optimizer_is_sharded(trainer) -> bool:
  """ZeRO v0 / v1 / v2 and the fused data-parallel path keep a 1/N shard of the optimizer state per rank; ZeRO-3 additionally
  shards the parameters.  Such state cannot be written by the first replica alone."""
  return any(trainer._sharded.values()) or bool(getattr(trainer, "zero3", {}))


def save_checkpoint(trainer, directory: str, bucket_bytes: int = BUCKET_BYTES) -> Optional[List[str]]:
  """Model weights: the first replica of every pipeline stage writes its stage (with split taskgraphs every rank writes its
  shards).  Optimizer state: the first replica, or — when it is sharded over the data-parallel ranks (ZeRO, fused path) —
  every rank writes its own shard under ``rank<r>/`` (the reference skips optimizer slots under ZeRO, ``hooks.py:340-344``,
  so a resumed run silently restarts them; here they resume exactly)."""
  trainer.build()
  per_rank_dirs = trainer.plan.num_stages > 1 or trainer.has_split
  must_write = trainer.is_first_replica or getattr(trainer, "has_split", False)
  sharded = _optimizer_is_sharded(trainer)
  files = None
  zero3 = getattr(trainer, "zero3", {})
  for z in zero3.values():                      # ZeRO-3 keeps parameters released between steps: materialise them for the export
    z.gather_all()
  try:
    if must_write:
      sub = os.path.join(directory, "rank%d" % _rank()) if per_rank_dirs else directory
      builder = MemoryEfficientBuilder(sub, bucket_bytes)
      state = {}
      for s in trainer.plan.local_stages:
        for k, v in trainer.stage_modules[s].state_dict().items():
          state["stage%d.%s" % (s, k)] = v
      files = builder.save(state, "model", {"global_step": trainer.global_step, "loss_scale": trainer.scaler.loss_scale,
                                            "loss_scale_good_steps": int(getattr(trainer.scaler, "good_steps", 0))})
      if not sharded:
        builder.save(_optimizer_state(trainer), "optim")
  finally:
    for z in zero3.values():
      z.release_all()
  if sharded:
    MemoryEfficientBuilder(os.path.join(directory, "rank%d" % _rank()), bucket_bytes).save(
        _optimizer_state(trainer), "optim", {"global_step": trainer.global_step})
  import torch.distributed as dist
  if dist.is_initialized():
    dist.barrier()
  return files


def _reshard_optimizer_state(trainer, directory: str, restored: set) -> None:
  """Elastic resume: the checkpoint was written with another data-parallel degree (or without sharding), so no file holds this
  rank's shard.  A flat bucket's optimizer state is the concatenation of the old ranks' shards in rank order — the parameters sit
  at the same (world-independent) offsets, only the tail padding differs — so it is rebuilt from every ``rank<r>/optim`` (or the
  unsharded ``optim``) and re-sliced for the new layout.  The reference cannot do this (its ZeRO shards are never merged,
  ``hooks.py:546-547``).  Pipelines, split taskgraphs and ZeRO-3 units keep the restart-from-weights fallback."""
  todo = [(s, i) for s in trainer.group_keys for i in range(len(trainer.optimizers[s])) if (s, i) not in restored]
  if not todo or trainer.plan.num_stages > 1 or getattr(trainer, "has_split", False) or getattr(trainer, "zero3", {}):
    return
  parts = []
  if os.path.exists(os.path.join(directory, "optim.index.json")):
    parts = [MemoryEfficientBuilder(directory).load("optim")[0]]            # written unsharded: one full copy
  else:
    ranks = sorted(int(d[4:]) for d in os.listdir(directory) if d.startswith("rank") and d[4:].isdigit()
                   and os.path.exists(os.path.join(directory, d, "optim.index.json")))
    if ranks != list(range(len(ranks))):
      return                                                                  # a rank directory is missing: cannot rebuild
    parts = [MemoryEfficientBuilder(os.path.join(directory, "rank%d" % r)).load("optim")[0] for r in ranks]
  if not parts:
    return
  from easyparallellibrary_b200.utils.logging import get_logger
  for s, i in todo:
    b, o = trainer.flats[s].buckets[i], trainer.optimizers[s][i]
    pre = "g%d.b%d." % (s, i)
    if o.kind == "torch" or not all(pre + "master" in part for part in parts):
      continue
    used = max((off + p.numel() for p, off in zip(b.params, b.offsets)), default=0)
    comm, sharded = trainer.dp_comms[s], trainer._sharded[s]
    lo, hi = b.shard_range(comm.rank if sharded else 0, comm.size if sharded else 1)
    sd = {"step": int(parts[0][pre + "step"])}
    ok = True
    for k in ("master", "m", "v"):
      if getattr(o, k, None) is None:
        sd[k] = None
        continue
      if not all(part.get(pre + k) is not None for part in parts):
        ok = False
        break
      full = torch.cat([part[pre + k].reshape(-1) for part in parts])
      if full.numel() < used:
        ok = False
        break
      piece = torch.zeros(hi - lo, dtype=full.dtype)
      n = max(min(hi, full.numel()) - lo, 0)
      if n:
        piece[:n] = full[lo:lo + n]
      sd[k] = piece
    if ok:
      o.load_state_dict(sd)
      restored.add((s, i))
      get_logger().info("optimizer state of bucket %d.%d re-sliced from %d checkpoint shard(s)", s, i, len(parts))


def load_checkpoint(trainer, directory: str) -> int:
  """Restore on every rank from the files written by ``save_checkpoint``; returns the restored global step."""
  trainer.build()
  mine = os.path.join(directory, "rank%d" % _rank())
  sub = mine
  if not os.path.exists(os.path.join(sub, "model.index.json")):
    # written by the first replica only: find the writer that holds the same stage
    sub = directory
    if trainer.plan.num_stages > 1:
      stage = trainer.plan.local_stages[0]
      sub = os.path.join(directory, "rank%d" % trainer.plan.stage_ranks[stage][0][0])
  builder = MemoryEfficientBuilder(sub)
  state, extra = builder.load("model")
  zero3 = getattr(trainer, "zero3", {})
  if not zero3:
    for s in trainer.plan.local_stages:
      prefix = "stage%d." % s
      sd = {k[len(prefix):]: v for k, v in state.items() if k.startswith(prefix)}
      own = trainer.stage_modules[s].state_dict()
      for k, v in sd.items():
        if k in own:
          own[k].copy_(v.to(own[k].dtype))
  # optimizer state: this rank's own shard if one was written, else the first replica's full state
  opt_dir = mine if os.path.exists(os.path.join(mine, "optim.index.json")) else sub
  restored = set()
  if os.path.exists(os.path.join(opt_dir, "optim.index.json")):
    opt, _ = MemoryEfficientBuilder(opt_dir).load("optim")
    for s in trainer.group_keys:
      for i, o in enumerate(trainer.optimizers[s]):
        pre = "g%d.b%d." % (s, i)
        if pre + "master" in opt and opt[pre + "master"].numel() == o.master.numel():
          sd = {"step": int(opt[pre + "step"]), "master": opt[pre + "master"], "m": opt.get(pre + "m"), "v": opt.get(pre + "v")}
          sd.update({k[len(pre):]: v for k, v in opt.items() if k.startswith(pre + "t")})      # wrapped torch optimizers: "t<i>.<name>"
          o.load_state_dict(sd)
          restored.add((s, i))
    for s, z in zero3.items():
      sds = []
      for i, u in enumerate(z.units):
        pre = "z%d.u%d." % (s, i)
        sds.append({k[len(pre):]: (int(v) if k.endswith(".step") else v) for k, v in opt.items() if k.startswith(pre)})
      if all("shard_param" in sd for sd in sds) and all(
          sd["shard_param"].numel() == u.shard_numel for sd, u in zip(sds, z.units)):
        z.load_state_dict(sds)
        restored.add(("zero3", s))
    for s, z in zero3.items():
      if ("zero3", s) in restored:
        continue
      # no shard state for this rank (checkpoint written without ZeRO-3, with another world size, or a rank directory is
      # missing): rebuild the shards from the gathered model weights instead of silently keeping the initial values
      prefix = "stage%d." % s
      sd = {k[len(prefix):]: v for k, v in state.items() if k.startswith(prefix)}
      names = {id(p): n for n, p in trainer.stage_modules[s].named_parameters()}
      missing = [names.get(id(p), "?") for u in z.units for p in u.params if names.get(id(p)) not in sd]
      if missing:
        raise RuntimeError("checkpoint %s holds neither this rank's ZeRO-3 shards nor the full weights of %s" % (directory, missing[:5]))
      for u in z.units:
        full = torch.zeros(u.numel, dtype=u.dtype, device=u.device)
        for p, o in zip(u.params, u.offsets):
          src = sd[names[id(p)]]
          full[o:o + src.numel()].copy_(src.reshape(-1).to(u.dtype))
        lo = u.comm.rank * u.shard_numel
        shard = full[lo:lo + u.shard_numel]
        (u.shard_host if u.offload else u.shard_param).copy_(shard)
        u.opt.master.copy_(shard.to(u.opt.master.dtype))           # moments restart at zero: the optimizer state was not in the checkpoint
      restored.add(("zero3", s))
  _reshard_optimizer_state(trainer, directory, restored)
  # optimizers without restored state restart from the restored weights (fp32 master = parameters)
  for s in trainer.group_keys:
    comm, flat = trainer.dp_comms[s], trainer.flats[s]
    for i, (b, o) in enumerate(zip(flat.buckets, trainer.optimizers[s])):
      if (s, i) not in restored:
        sharded = trainer._sharded[s]
        lo, hi = b.shard_range(comm.rank if sharded else 0, comm.size if sharded else 1)
        o.master.copy_(b.flat_param[lo:hi].to(o.master.dtype))
  trainer.global_step = int(extra.get("global_step", 0))
  if hasattr(trainer.scaler, "loss_scale") and "loss_scale" in extra:
    trainer.scaler.loss_scale = extra["loss_scale"]
  if hasattr(trainer.scaler, "good_steps") and "loss_scale_good_steps" in extra:
    trainer.scaler.good_steps = int(extra["loss_scale_good_steps"])
  return trainer.global_step
