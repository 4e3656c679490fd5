"""From annotations to a parallel plan.

The reference transforms the graph at ``finalize`` (``parallel.py:211-231``);
the eager equivalent is a one-shot *plan*: lay the taskgraphs out on the
cluster (``Cluster`` / ``VirtualDevice``), then answer, for *this rank*: which
pipeline stage(s) do I run, who are my data-parallel peers for each stage, who
are my pipeline neighbours, who is in my tensor-parallel group.

Layout rules follow the reference: replicas = total GPUs / sum(devices per
replica over taskgraphs) (``cluster.py:146-159``); with
``cluster.colocate_split_and_replicate`` every GPU is a 1-device replica and a
``split`` scope spans all of them (``cluster.py:108-118``,
``strategy_context.py:76-79``); one data-parallel communicator per stage
(``parallel/ops.py:461-465``).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

from easyparallellibrary_b200.ir.graph import Graph
from easyparallellibrary_b200.ir.taskgraph import Taskgraph


@dataclass
class StagePlacement:
  taskgraph: int
  replica: int                 # which model replica this rank belongs to (for this stage)
  device_slot: int             # index inside the replica's device list
  dp_ranks: List[int]          # same stage+slot across replicas (gradient reduction group)
  tp_ranks: List[int]          # devices of this replica for this taskgraph


@dataclass
class ParallelPlan:
  rank: int
  world: int
  num_stages: int
  num_replicas: int
  num_micro_batch: int
  pipeline: bool
  colocated: bool                                   # every taskgraph lives on every rank
  stage_of_rank: Dict[int, List[int]] = field(default_factory=dict)   # rank -> stage indices it runs
  local_stages: List[int] = field(default_factory=list)
  placements: Dict[int, StagePlacement] = field(default_factory=dict) # taskgraph index -> placement on this rank
  stage_ranks: List[List[List[int]]] = field(default_factory=list)    # [stage][replica] -> ranks
  split_taskgraphs: List[int] = field(default_factory=list)

  def prev_rank(self, stage: int) -> Optional[int]:
    if stage <= 0:
      return None
    pl = self.placements[self.stage_taskgraphs[stage]]
    return self.stage_ranks[stage - 1][pl.replica][0]

  def next_rank(self, stage: int) -> Optional[int]:
    if stage >= self.num_stages - 1:
      return None
    pl = self.placements[self.stage_taskgraphs[stage]]
    return self.stage_ranks[stage + 1][pl.replica][0]

  stage_taskgraphs: List[int] = field(default_factory=list)

  def describe(self) -> str:
    lines = ["ParallelPlan(rank=%d/%d stages=%d replicas=%d micro_batches=%d pipeline=%s colocated=%s)" % (
        self.rank, self.world, self.num_stages, self.num_replicas, self.num_micro_batch, self.pipeline, self.colocated)]
    for s, reps in enumerate(self.stage_ranks):
      lines.append("  stage %d: %s" % (s, reps))
    return "\n".join(lines)


def build_plan(graph: Optional[Graph] = None, cluster=None, config=None, rank: Optional[int] = None) -> ParallelPlan:
  from easyparallellibrary_b200.env import Env
  env = Env.get()
  graph = graph or Graph.get()
  cluster = cluster or env.cluster
  config = config or env.config
  if cluster is None:
    from easyparallellibrary_b200.cluster import Cluster
    cluster = env.cluster = Cluster()
  if not graph.taskgraphs:
    graph.current_taskgraph(create=True)
  tgs: List[Taskgraph] = graph.taskgraphs
  stage_tgs = [t for t in tgs if t.is_replicate]
  split_tgs = [t for t in tgs if t.is_split]
  world = cluster.total_gpu_num
  me = cluster.rank if rank is None else rank
  if me is None:
    me = 0
  M = config.pipeline.num_micro_batch
  colocate = config.cluster.colocate_split_and_replicate or world == 1

  plan = ParallelPlan(rank=me, world=world, num_stages=max(len(stage_tgs), 1), num_replicas=1, num_micro_batch=M,
                      pipeline=len(stage_tgs) > 1 and M > 1, colocated=False)
  plan.stage_taskgraphs = [t.index for t in stage_tgs] or [tgs[0].index]
  plan.split_taskgraphs = [t.index for t in split_tgs]

  if colocate:
    # every GPU is a 1-device replica of every replicate taskgraph; split scopes span all GPUs
    vds = cluster.generate_virtual_devices("all")
    all_ranks = [d.rank for d in vds[0].all_devices]
    for t in tgs:
      t.virtual_device = vds[0]
    plan.colocated = True
    plan.num_replicas = world
    plan.stage_ranks = [[[r] for r in all_ranks] for _ in plan.stage_taskgraphs]
    for t in tgs:
      if t.is_split:
        n = t.strategy.device_count or world
        if world % n:
          raise RuntimeError("split(device_count=%d) does not divide the %d available GPUs" % (n, world))
        base = (me // n) * n
        tp = list(range(base, base + n))
        plan.placements[t.index] = StagePlacement(t.index, me // n, me - base, [r for r in all_ranks if r % n == me % n], tp)
      else:
        plan.placements[t.index] = StagePlacement(t.index, me, 0, all_ranks, [me])
    plan.local_stages = list(range(len(plan.stage_taskgraphs)))
    plan.stage_of_rank = {r: list(plan.local_stages) for r in all_ranks}
    plan.pipeline = False       # all stages on every rank: micro-batches become gradient accumulation
    if len(stage_tgs) > 1 and world > 1:
      from easyparallellibrary_b200.utils.logging import get_logger
      get_logger().warning(
          "%d replicate taskgraphs are colocated on every GPU (cluster.colocate_split_and_replicate): they run back to back, "
          "not as a pipeline; %s", len(stage_tgs),
          "the %d micro-batches are gradient accumulation" % M if M > 1 else "set the flag to False for inter-layer placement")
    return plan

  counts = [t.num_device_per_replica for t in tgs]
  vds = cluster.generate_virtual_devices("auto", counts)
  for t, vd in zip(tgs, vds):
    t.virtual_device = vd
  plan.num_replicas = vds[0].num_replicas
  plan.stage_ranks = [tgs_vd.ranks() for tgs_vd in (tgs[i].virtual_device for i in plan.stage_taskgraphs)]
  for t in tgs:
    vd = t.virtual_device
    rep = vd.replica_of_rank(me)
    if rep is None:
      continue
    devs = [d.rank for d in vd.slice_devices[rep]]
    slot = devs.index(me)
    dp = [vd.slice_devices[r][slot].rank for r in range(vd.num_replicas)]
    plan.placements[t.index] = StagePlacement(t.index, rep, slot, dp, devs)
  plan.local_stages = [s for s, ti in enumerate(plan.stage_taskgraphs) if ti in plan.placements]
  for s, reps in enumerate(plan.stage_ranks):
    for devs in reps:
      for r in devs:
        plan.stage_of_rank.setdefault(r, []).append(s)
  return plan
