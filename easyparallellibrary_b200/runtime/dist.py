"""Process-group bootstrap (control plane).

The reference bootstraps through a per-worker ``tf.train.Server`` and a TF
collective broadcast of the NCCL id (``epl/env.py:171-183``,
``communicators/base.py:44-73``).  Here the control plane is
``torch.distributed``'s TCPStore; the launcher (``torchrun`` or ``epl-launch``)
provides ``RANK/WORLD_SIZE/MASTER_ADDR/MASTER_PORT``; EPL-style ``TF_CONFIG``
is translated when those are absent.
"""
from __future__ import annotations

import datetime
import json
import os

import torch
import torch.distributed as dist


def _translate_tf_config() -> None:
  cfg = os.environ.get("TF_CONFIG")
  if not cfg or "RANK" in os.environ:
    return
  spec = json.loads(cfg)
  workers = spec.get("cluster", {}).get("chief", []) + spec.get("cluster", {}).get("worker", [])
  task = spec.get("task", {})
  idx = int(task.get("index", 0)) + (len(spec.get("cluster", {}).get("chief", [])) if task.get("type") == "worker" else 0)
  if len(workers) > 1:
    host, port = workers[0].rsplit(":", 1)
    os.environ.setdefault("MASTER_ADDR", host)
    os.environ.setdefault("MASTER_PORT", port)
    os.environ.setdefault("WORLD_SIZE", str(len(workers)))
    os.environ.setdefault("RANK", str(idx))


def local_device() -> torch.device:
  if torch.cuda.is_available():
    idx = int(os.environ.get("LOCAL_RANK", "0")) % max(torch.cuda.device_count(), 1)
    return torch.device("cuda", idx)
  return torch.device("cpu")


def ensure_process_group(timeout_s: int = 1800) -> None:
  _translate_tf_config()
  if dist.is_initialized():
    return
  world = int(os.environ.get("WORLD_SIZE", "1"))
  if world <= 1:
    return
  dev = local_device()
  if dev.type == "cuda":
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", timeout=datetime.timedelta(seconds=timeout_s), device_id=dev)
  else:
    dist.init_process_group("gloo", timeout=datetime.timedelta(seconds=timeout_s))


def world_size() -> int:
  return dist.get_world_size() if dist.is_initialized() else 1


def rank() -> int:
  return dist.get_rank() if dist.is_initialized() else 0


def shutdown() -> None:
  if dist.is_initialized():
    dist.destroy_process_group()
