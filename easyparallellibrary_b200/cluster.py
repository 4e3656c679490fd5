"""Cluster description and device layouts.

A *cluster* is ``worker_num`` workers (nodes / launch groups) each exposing
``gpu_num_per_worker`` B200s.  At run time every GPU is driven by its own
process (SPMD, ``torch.distributed``), so a :class:`Device` also carries the
global rank of the process that owns it.  A :class:`VirtualDevice` is the 2-D
``slice_devices[replica][device]`` table that one taskgraph runs on.

Parity with the reference (``epl/cluster.py``): layouts ``all`` (108-118),
``auto`` (146-159, replicas = total GPUs / sum of devices-per-replica, error if
not divisible), ``specific`` (162-166) and ``aware_row`` (169-241, 1 GPU per
worker, hosts regrouped by machine); row-major placement when
``cluster.device_place_prefer_intra_node`` else column-major (121-143);
``TF_CONFIG`` parsing with chief normalisation (301-351).  The torchrun
environment (``RANK/WORLD_SIZE/LOCAL_WORLD_SIZE``) is the native input here.
"""
from __future__ import annotations

import json
import os
from typing import Iterable, List, NamedTuple, Optional, Sequence, Union

from easyparallellibrary_b200.utils import constant


class Device(NamedTuple):
  worker: int
  index: int
  rank: int
  kind: str = "GPU"

  def __str__(self) -> str:  # familiar spelling for users coming from the reference
    return "/job:worker/replica:0/task:%d/device:%s:%d" % (self.worker, self.kind, self.index)


def parse_device(spec: Union[str, Device], gpus_per_worker: int) -> Device:
  if isinstance(spec, Device):
    return spec
  task, index = 0, 0
  for part in spec.strip("/").split("/"):
    if part.startswith("task:"):
      task = int(part[5:])
    elif part.startswith("device:"):
      index = int(part.rsplit(":", 1)[1])
  return Device(task, index, task * gpus_per_worker + index)


class VirtualDevice(object):
  """Devices of one taskgraph: ``slice_devices[replica] -> [Device, ...]``."""

  def __init__(self, slice_devices: Sequence[Sequence[Device]], local_worker: int, local_rank_set: Iterable[int]):
    self._slices = [list(s) for s in slice_devices]
    self._local_worker = local_worker
    self._local_ranks = set(local_rank_set)

  @property
  def slice_devices(self) -> List[List[Device]]:
    return self._slices

  @property
  def num_replicas(self) -> int:
    return len(self._slices)

  @property
  def num_devices_per_replica(self) -> int:
    return len(self._slices[0]) if self._slices else 0

  @property
  def all_devices(self) -> List[Device]:
    return [d for s in self._slices for d in s]

  @property
  def local_devices(self) -> List[Device]:
    """Devices hosted by this worker (reference semantics: task == worker_index)."""
    return [d for d in self.all_devices if d.worker == self._local_worker]

  @property
  def owned_devices(self) -> List[Device]:
    """Devices driven by *this process* (SPMD: normally exactly one)."""
    return [d for d in self.all_devices if d.rank in self._local_ranks]

  def get_device(self, replica_idx: int, device_idx: int) -> Device:
    return self._slices[replica_idx][device_idx]

  def replica_of_rank(self, rank: int) -> Optional[int]:
    for r, devs in enumerate(self._slices):
      if any(d.rank == rank for d in devs):
        return r
    return None

  def ranks(self) -> List[List[int]]:
    return [[d.rank for d in s] for s in self._slices]

  def __repr__(self) -> str:
    return "VirtualDevice(%s)" % [[str(d) for d in s] for s in self._slices]


def _device_stream(worker_num: int, gpus: int, prefer_row: bool):
  if prefer_row:
    for w in range(worker_num):
      for g in range(gpus):
        yield Device(w, g, w * gpus + g)
  else:
    for g in range(gpus):
      for w in range(worker_num):
        yield Device(w, g, w * gpus + g)


def _group_hosts_by_machine(hosts: List[str], worker_index: int):
  """aware_row: put workers of one machine next to each other, rank 0 first."""
  by_machine: "Dict[str, List[int]]" = {}
  for i, h in enumerate(hosts):
    by_machine.setdefault(h.split(":")[0], []).append(i)
  sizes = {len(v) for v in by_machine.values()}
  if len(sizes) > 1:
    raise RuntimeError("Number of workers must be the same for each machine.")
  order: List[int] = []
  for members in by_machine.values():
    if 0 in members:
      order = members + order
    else:
      order = order + members
  new_hosts = [hosts[i] for i in order]
  return new_hosts, new_hosts.index(hosts[worker_index]), sizes.pop()


class Cluster(object):
  """``Cluster(worker_hosts=..., worker_index=..., layout=...)``.

  ``layout`` is ``None`` (lazy: decided when the taskgraphs are known),
  ``"all"``, ``"auto"``, ``{"auto": [n0, n1, ...]}`` (devices per replica of
  each taskgraph), ``{"specific": [[[dev,...],...],...]}`` or
  ``{"aware_row": gpus_per_slice}``.
  """

  def __init__(self, worker_hosts: Optional[str] = None, ps_hosts: Optional[str] = None,
               job_name: str = "worker", worker_index: Optional[int] = None, layout=None,
               gpus_per_worker: Optional[int] = None, rank: Optional[int] = None,
               prefer_intra_node: Optional[bool] = None):
    self.ps_hosts = ps_hosts
    self.job_name = job_name
    self._virtual_devices: List[VirtualDevice] = []
    self._prefer_row = True if prefer_intra_node is None else prefer_intra_node
    self._from_env(worker_hosts, worker_index, gpus_per_worker, rank)
    self._layout = self._normalise_layout(layout)
    if self._layout and "aware_row" in self._layout:
      hosts, self.worker_index, self.worker_num_per_machine = _group_hosts_by_machine(
          self.hosts.split(","), self.worker_index)
      self.hosts = ",".join(hosts)
    if self._layout and ("auto" not in self._layout or isinstance(self._layout["auto"], (list, tuple))):
      self.generate_virtual_devices(self._layout)

  # ------------------------------------------------------------------ discovery
  def _from_env(self, worker_hosts, worker_index, gpus_per_worker, rank) -> None:
    env = os.environ
    tf_config = env.get(constant.ENV_TF_CONFIG)
    if worker_hosts is None and tf_config:
      cfg = json.loads(tf_config)
      spec = cfg.get("cluster", {})
      task = cfg.get("task", {})
      workers = list(spec.get("worker", []))
      chief = list(spec.get("chief", []))
      ttype, tindex = task.get("type", "worker"), int(task.get("index", 0))
      # a chief is worker 0; the remaining workers shift by one (reference cluster.py:301-351)
      if chief:
        workers = chief + workers
        if ttype == "worker":
          tindex += len(chief)
      if ttype == "ps":
        raise RuntimeError("parameter-server tasks are not supported; run workers only")
      worker_hosts = ",".join(workers) if workers else None
      if worker_index is None:
        worker_index = tindex
      self.ps_hosts = ",".join(spec.get("ps", [])) or self.ps_hosts
      # epl-launch exports both TF_CONFIG (workers) and the torch.distributed variables (one rank per GPU); several workers
      # may share a machine, so LOCAL_WORLD_SIZE counts the machine, not the worker: ranks per worker = WORLD_SIZE / workers
      if gpus_per_worker is None and workers and "WORLD_SIZE" in env and int(env["WORLD_SIZE"]) % len(workers) == 0:
        gpus_per_worker = int(env["WORLD_SIZE"]) // len(workers)
    if gpus_per_worker is None:
      gpus_per_worker = self.available_gpus()
    self.gpu_num_per_worker = max(int(gpus_per_worker), 1)
    if worker_hosts is None:
      world = int(env.get("WORLD_SIZE", "1"))
      # torchrun: every GPU is a rank; group ranks into nodes
      nodes = max(world // self.gpu_num_per_worker, 1) if world >= self.gpu_num_per_worker else 1
      if world < self.gpu_num_per_worker:
        self.gpu_num_per_worker = world
      worker_hosts = ",".join("127.0.0.1:%d" % (20000 + i) for i in range(nodes))
      if worker_index is None:
        worker_index = int(env.get("GROUP_RANK", int(env.get("RANK", "0")) // self.gpu_num_per_worker))
    self.hosts = worker_hosts
    self.worker_num = len(worker_hosts.split(","))
    self.worker_index = int(worker_index or 0)
    if rank is None:
      local = int(env.get("LOCAL_RANK", "0"))
      rank = int(env.get("RANK", self.worker_index * self.gpu_num_per_worker + local))
      if "RANK" not in env and "LOCAL_RANK" not in env and self.gpu_num_per_worker > 1:
        rank = None  # a single controller owning all local GPUs (tests / planning mode)
    self.rank = rank

  @staticmethod
  def available_gpus() -> int:
    """GPUs each worker exposes (overridable in tests, like the reference's mock point)."""
    env = os.environ
    if "LOCAL_WORLD_SIZE" in env:
      return int(env["LOCAL_WORLD_SIZE"])
    visible = env.get("EPL_CLUSTER_RUN_VISIBLE_DEVICES") or env.get("CUDA_VISIBLE_DEVICES")
    if visible:
      return len([v for v in visible.split(",") if v.strip()])
    try:
      import torch
      n = torch.cuda.device_count()
      return n if n > 0 else 1
    except Exception:  # pragma: no cover
      return 1

  # ------------------------------------------------------------------ properties
  @property
  def total_gpu_num(self) -> int:
    return self.worker_num * self.gpu_num_per_worker

  @property
  def virtual_devices(self) -> List[VirtualDevice]:
    return self._virtual_devices

  @property
  def available_devices(self) -> List[Device]:
    return list(_device_stream(self.worker_num, self.gpu_num_per_worker, True))

  @property
  def current_worker_cpu_device(self) -> Device:
    return Device(self.worker_index, 0, self.rank if self.rank is not None else 0, "CPU")

  def _local_rank_set(self) -> List[int]:
    if self.rank is not None:
      return [self.rank]
    g = self.gpu_num_per_worker
    return list(range(self.worker_index * g, (self.worker_index + 1) * g))

  # ------------------------------------------------------------------ layouts
  @staticmethod
  def _normalise_layout(layout) -> Optional[dict]:
    if layout is None:
      return None
    if isinstance(layout, str):
      return {layout.lower(): True}
    layout = {k.lower(): v for k, v in dict(layout).items()}
    if sum(k in layout for k in ("all", "auto", "specific")) > 1:
      raise ValueError("Can't set multiple layout to slice cluster. Layout: %s" % layout)
    return layout

  def _slices(self, layout: dict, device_counts: Optional[Sequence[int]]):
    g = self.gpu_num_per_worker
    if "all" in layout:
      return [[[d] for d in _device_stream(self.worker_num, g, True)]]
    if "specific" in layout:
      return [[[parse_device(d, g) for d in replica] for replica in tg] for tg in layout["specific"]]
    if "aware_row" in layout:
      per_slice = int(layout["aware_row"])
      if g != 1 or per_slice <= 0 or self.worker_num < per_slice or self.worker_num % per_slice:
        raise RuntimeError("aware_row needs 1 GPU per worker and worker count divisible by the slice size "
                           "(workers %d, GPUs per worker %d, slice %d)" % (self.worker_num, g, per_slice))
      n_slices = self.worker_num // per_slice
      return [[[Device(w, 0, w)] for w in range(s * per_slice, (s + 1) * per_slice)] for s in range(n_slices)]
    if "auto" in layout:
      counts = layout["auto"] if isinstance(layout["auto"], (list, tuple)) else device_counts
      if not counts:
        raise RuntimeError("auto layout needs the number of devices per replica of every taskgraph")
      per_replica = sum(counts)
      if per_replica <= 0 or self.total_gpu_num % per_replica:
        raise RuntimeError("Total devices {} is not divisible by num_device_per_replica {}".format(
            self.total_gpu_num, per_replica))
      replicas = self.total_gpu_num // per_replica
      stream = _device_stream(self.worker_num, g, self._prefer_row)
      slices = [[[] for _ in range(replicas)] for _ in counts]
      for r in range(replicas):
        for t, n in enumerate(counts):
          for _ in range(n):
            slices[t][r].append(next(stream))
      return slices
    raise RuntimeError("Layout is not supported. Layout: %s ." % layout)

  def generate_virtual_devices(self, layout="auto", device_counts: Optional[Sequence[int]] = None):
    """(Re)build the virtual devices; returns them."""
    layout = self._normalise_layout(layout)
    self._layout = layout
    local = self._local_rank_set()
    self._virtual_devices = [VirtualDevice(tg, self.worker_index, local)
                             for tg in self._slices(layout, device_counts)]
    return self._virtual_devices

  def set_prefer_intra_node(self, flag: bool) -> None:
    self._prefer_row = bool(flag)

  # context-manager form kept for API familiarity (reference cluster.py:478-484)
  def __enter__(self):
    from easyparallellibrary_b200.env import Env
    self._prev = Env.get().cluster
    Env.get().cluster = self
    return self

  def __exit__(self, *exc):
    from easyparallellibrary_b200.env import Env
    Env.get().cluster = self._prev
    return False

  def __repr__(self) -> str:
    return "Cluster(workers=%d, gpus_per_worker=%d, worker_index=%d, rank=%s, layout=%s)" % (
        self.worker_num, self.gpu_num_per_worker, self.worker_index, self.rank, self._layout)
