"""The training engine: turns an annotated model into a running hybrid-parallel job.

This is the eager counterpart of the reference's ``Parallel.do_parallelism``
pipeline (``epl/parallel/parallel.py:211-231``: offload -> micro-batch clone ->
replica clone -> gradient aggregation -> schedule -> IO slicing -> output
merging).  Nothing is cloned: replicas are other ranks, micro-batches are loop
iterations, schedules are instruction lists, gradient aggregation is a set of
flat buckets in NVLink symmetric memory, each reduced + applied + re-gathered by one kernel.

Semantics kept from the reference:

* the batch handed to :meth:`Trainer.step` is split into
  ``pipeline.num_micro_batch`` micro-batches; with one stage that is gradient
  accumulation (``runtime/gradient_accumulation.py:40-50``), with several it is
  the pipeline;
* gradients: sum over micro-batches (/M if ``mean``) -> reduce over replicas
  (/N if ``mean``) -> clip -> apply (``graph_editor.py:610-725``);
  ``communication.clip_after_allreduce`` picks clip-then-reduce vs
  reduce-then-clip (``hooks.py:173-179``);
* weights are broadcast from the first replica once (``hooks.py:330-357``);
* ``GraphKeys`` collections are merged over micro-batches and replicas and
  returned to the caller (``parallel.py:233-353``);
* ZeRO ``v0`` shards optimizer state, ``v1`` additionally shards gradients
  (``runtime/zero.py:88-175``) — here as equal flat shards per bucket instead of
  variable-granular ownership, which is what makes reduce-scatter (and the fused
  NVLink kernel) applicable; grouped apply (``optimizer.num_apply_group``)
  bounds optimizer temporaries (``optimizer_helper.py:75-128``).
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import torch
from torch import nn

from easyparallellibrary_b200.communicators.collective_communicator import CollectiveCommunicator
from easyparallellibrary_b200.env import Env
from easyparallellibrary_b200.ir.graph import Graph, GraphKeys
from easyparallellibrary_b200.ir.phase import ModelPhase, phase_scope
from easyparallellibrary_b200.parallel.flat import Bucket, FlatParameters
from easyparallellibrary_b200.parallel.plan import ParallelPlan, build_plan
from easyparallellibrary_b200.runtime import amp as amp_lib
from easyparallellibrary_b200.runtime.optimizer import FlatOptimizer, make_hyper
from easyparallellibrary_b200.utils import constant
from easyparallellibrary_b200.utils.logging import get_logger


@dataclass
class StepOutput:
  loss: Optional[torch.Tensor] = None
  collections: "OrderedDict[str, List[Any]]" = field(default_factory=OrderedDict)
  skipped: bool = False
  grad_norm: Optional[torch.Tensor] = None
  loss_scale: float = 1.0

  def item(self) -> float:
    return float(self.loss) if self.loss is not None else float("nan")


def default_no_decay(p: nn.Parameter) -> bool:
  return p.dim() <= 1 or getattr(p, "epl_no_decay", False)


def _split_batch(batch: Sequence[Any], m: int) -> List[Tuple[Any, ...]]:
  if m == 1:
    return [tuple(batch)]
  cols = []
  for x in batch:
    if isinstance(x, torch.Tensor) and x.dim() > 0:
      if x.shape[0] % m:
        raise ValueError("batch dimension %d is not divisible by pipeline.num_micro_batch=%d" % (x.shape[0], m))
      cols.append(x.chunk(m, 0))
    else:
      cols.append([x] * m)
  return [tuple(c[i] for c in cols) for i in range(m)]


def sequential_layers(model: nn.Module) -> Optional[List[nn.Module]]:
  fn = getattr(model, "epl_sequential", None)
  if callable(fn):
    return list(fn())
  if isinstance(model, nn.Sequential):
    return list(model)
  return None


class Trainer(object):
  """``Trainer(model, optimizer="adamw", lr=..., loss_fn=...)`` then ``trainer.step(*batch)``.

  ``loss_fn(output, *batch[1:])`` is applied to the model output; without a
  ``loss_fn`` the model itself must return the loss (``model(*batch)``).
  """

  def __init__(self, model: nn.Module, optimizer: str = "adamw", loss_fn: Optional[Callable] = None,
               device: Optional[torch.device] = None, max_grad_norm: Optional[float] = None,
               no_decay: Callable[[nn.Parameter], bool] = default_no_decay, example_inputs: Optional[Sequence[Any]] = None,
               baseline: bool = False, cuda_graph: Optional[bool] = None, **opt_kwargs):
    env = Env.get()
    if not env.is_initialized:
      env.init(None)
    self.env = env
    self.config = env.config
    self.model = model
    self.loss_fn = loss_fn
    if isinstance(optimizer, torch.optim.Optimizer):
      # an optimizer INSTANCE built over model.parameters() the usual PyTorch way: its class and defaults are taken over (the
      # instance itself is not used: the engine owns flat fp32 master shards).  Per-group settings other than weight decay are
      # not carried; parameters of groups with weight_decay == 0 join the no-decay set.
      inst = optimizer
      defaults = {k: v for k, v in inst.defaults.items() if k not in ("params",)}
      if len({g.get("lr", defaults.get("lr")) for g in inst.param_groups}) > 1:
        get_logger().warning("Trainer: the optimizer's param groups use different learning rates; only the default (%s) is used",
                             defaults.get("lr"))
      wds = [g.get("weight_decay", defaults.get("weight_decay", 0.0)) for g in inst.param_groups]
      zero_wd = {id(p) for g, w in zip(inst.param_groups, wds) if not w for p in g["params"]}
      if any(wds):
        defaults["weight_decay"] = max(wds)
      if no_decay is default_no_decay:             # follow the instance exactly: torch decays biases too unless a group says otherwise
        no_decay = lambda p, _z=zero_wd: id(p) in _z                           # noqa: E731
      else:
        user_nd = no_decay
        no_decay = lambda p, _u=user_nd, _z=zero_wd: id(p) in _z or _u(p)      # noqa: E731
      defaults.update(opt_kwargs)
      opt_kwargs = defaults
      optimizer = type(inst)
    if isinstance(optimizer, type) and issubclass(optimizer, torch.optim.Optimizer):
      # any torch optimizer class (Adagrad, RMSprop, ...): runs on the fp32 master shards through the library path, so ZeRO
      # sharding, gradient accumulation, clipping, loss scaling and checkpoints keep working; the fused kernels and the
      # CUDA-graph step need one of the built-in optimizers
      opt_kwargs = dict(opt_kwargs, factory=optimizer)
      optimizer = "torch"
    if not isinstance(optimizer, str):          # an optimizer description object (e.g. ops.AdamWeightDecayOptimizer)
      extra = optimizer.trainer_kwargs(model)
      no_decay = extra.pop("no_decay", no_decay)
      extra.update(opt_kwargs)
      opt_kwargs = extra
      optimizer = optimizer.kind
    self.opt_kind = optimizer.lower()
    # lr may be a schedule: a callable global_step -> learning rate (runtime/lr_schedule.py), evaluated before every step.  Every
    # optimizer shard shares this one hyper-parameter object and reads lr per step (the CUDA-graph / fused paths through their
    # device-side step values), so a schedule costs nothing and survives checkpoints (it is a function of global_step)
    self.lr_schedule = opt_kwargs["lr"] if callable(opt_kwargs.get("lr")) else None
    if self.lr_schedule is not None:
      opt_kwargs = dict(opt_kwargs, lr=float(self.lr_schedule(0)))
    self.hyper = make_hyper(self.opt_kind, **opt_kwargs)
    self.max_grad_norm = max_grad_norm
    self.no_decay = no_decay
    self.example_inputs = example_inputs
    self.baseline = baseline            # reference-equivalent library path (all-reduce + unfused optimizer)
    self.cuda_graph = cuda_graph        # True: capture the whole step in a CUDA graph when the configuration allows it
    self._device = device
    self._built = False
    self.global_step = 0
    self.scaler = amp_lib.make_scaler(self.config.amp.level, self.config.amp.loss_scale)
    self.hooks: List[Any] = []          # profiler hooks: before_step(trainer) / after_step(trainer, out)
    self.launch_count = 0

  # ================================================================== build
  def build(self) -> "Trainer":
    if self._built:
      return self
    from easyparallellibrary_b200.runtime.dist import local_device
    cfg, env = self.config, self.env
    graph = Graph.get()
    self.device = self._device or local_device()
    if self.opt_kind == "torch" and cfg.offload.level:
      raise ValueError("offload.level=%s streams the optimizer state through a device window for the built-in optimizers "
                       "(adam | adamw | sgd); a torch.optim class keeps its own state tensors" % cfg.offload.level)
    if cfg.auto.auto_parallel and cfg.pipeline.num_stages > 1:
      self._auto_stages(graph)
    self.plan: ParallelPlan = build_plan(graph, env.cluster, cfg)
    env.parallel_information[constant.STAGE_POLICY_HEURISTIC] = self.plan.describe()
    get_logger().info(self.plan.describe())

    # ---- which modules run here ------------------------------------------------------
    self.stage_modules: Dict[int, nn.Module] = {}
    layers = sequential_layers(self.model)
    if self.plan.num_stages > 1:
      if layers is None:
        raise RuntimeError("several replicate taskgraphs need a sequential model: an nn.Sequential or a model "
                           "with epl_sequential() returning its layers in call order")
      assign = self._assign_layers(layers, graph)
      for s in range(self.plan.num_stages):
        mods = [l for l, a in zip(layers, assign) if a == s]
        self.stage_modules[s] = nn.Sequential(*mods)
    else:
      self.stage_modules[0] = self.model
    local = [self.stage_modules[s] for s in self.plan.local_stages]

    # ---- precision, placement, recompute ---------------------------------------------
    self.compute_dtype = amp_lib.compute_dtype(cfg.amp.level)
    self.o1 = (cfg.amp.level or "").lower() == "o1"       # fp32 parameters, per-op fp16/fp32 policy (runtime/amp.py)
    from easyparallellibrary_b200.ops import fp8 as fp8_lib
    fp8_lib.ENABLED = (cfg.amp.level or "").lower() == "fp8" and self.device.type == "cuda"
    env.parallel_plan = self.plan
    for m in local:
      if self.compute_dtype is not None:
        amp_lib.cast_module(m, self.compute_dtype, cfg.amp.debug_log)
      _materialize(m, self.device)
    if cfg.gradient_checkpoint.type:
      from easyparallellibrary_b200.runtime.gradient_checkpoint import apply_gradient_checkpoint
      nodes = None
      if cfg.gradient_checkpoint.type == constant.GC_AUTO and self.example_inputs is not None and self.plan.num_stages == 1:
        from easyparallellibrary_b200.ir.capture import trace_module_costs
        try:
          ex = [_to_device(x, self.device) for x in self.example_inputs]
          nodes = trace_module_costs(self.model, ex)
        except Exception as e:  # pragma: no cover - tracing is best effort
          get_logger().warning("auto gradient checkpoint: tracing failed (%s); using module structure", e)
      for m in local:
        wrapped = apply_gradient_checkpoint(m, cfg.gradient_checkpoint.type, nodes,
                                            graph.get_collection(GraphKeys.GC_CHECKPOINTS),
                                            cfg.gradient_checkpoint.end_taskgraph)
        get_logger().info("gradient checkpoint: %d segment(s)", len(wrapped))

    # ---- data-parallel groups, flat storage, optimizer ---------------------------------
    self.dp_comms: Dict[int, CollectiveCommunicator] = {}
    self.flats: Dict[int, FlatParameters] = {}
    self.optimizers: Dict[int, List[FlatOptimizer]] = {}
    zero = cfg.zero.level
    # process groups are created collectively and in the same order on every rank
    from easyparallellibrary_b200.communicators.backend import register_groups
    dp_groups = []
    for reps in self.plan.stage_ranks:
      for slot in range(len(reps[0])):
        ranks = [devs[slot] for devs in reps]
        if ranks not in dp_groups:
          dp_groups.append(ranks)
    register_groups(dp_groups, copies=cfg.communication.num_communicators)
    # parameter groups: one per local pipeline stage (data-parallel over that stage's replicas) plus one per
    # split taskgraph (tensor-parallel shards: reduced only over the replicas that hold the *same* shard)
    split_owner = {}
    for ti in self.plan.split_taskgraphs:
      for p in graph.taskgraphs[ti].parameters:
        split_owner[id(p)] = ti
    self.group_keys: List[int] = []
    group_params: Dict[int, List[nn.Parameter]] = {}
    group_ranks: Dict[int, List[int]] = {}
    seen = set()
    # parameters whose gradient is sparse (nn.Embedding(sparse=True)): kept out of the dense flat buckets; their gradients
    # travel as (indices, values) all-gathers (reference rewriters/sparse_allreduce.py:127-160) unless
    # communication.sparse_as_dense densifies them first (reference parallel/hooks.py:162-167)
    sparse_ids = set()
    for mod in self.stage_modules.values():
      for m in mod.modules():
        if isinstance(m, (nn.Embedding, nn.EmbeddingBag)) and getattr(m, "sparse", False):
          sparse_ids.add(id(m.weight))
    self._sparse: Dict[int, List[Tuple[nn.Parameter, FlatOptimizer]]] = {}
    for s in self.plan.local_stages:
      pl = self.plan.placements[self.plan.stage_taskgraphs[s]]
      mine = []
      for p in self.stage_modules[s].parameters():
        if id(p) in seen or not p.requires_grad:
          continue
        seen.add(id(p))
        if id(p) in sparse_ids:
          self._sparse.setdefault(s, []).append(p)
          continue
        ti = split_owner.get(id(p))
        if ti is not None and ti in self.plan.placements:
          key = 1000 + ti
          if key not in group_params:
            group_params[key], group_ranks[key] = [], self.plan.placements[ti].dp_ranks
          group_params[key].append(p)
        else:
          mine.append(p)
      group_params[s], group_ranks[s] = mine, pl.dp_ranks
    self.group_keys = [k for k in list(self.plan.local_stages) + sorted(k for k in group_params if k >= 1000)]
    self.has_split = any(k >= 1000 for k in self.group_keys)
    from easyparallellibrary_b200.ops.tensor_parallel import Replica2Split
    self._split_gathers_batch = any(isinstance(m, Replica2Split) and (m.group is None or m.group.size > 1)
                                    for mod in self.stage_modules.values() for m in mod.modules())
    self._split_size = 1
    for ti in self.plan.split_taskgraphs:
      self._split_size = max(self._split_size, graph.taskgraphs[ti].strategy.device_count or self.plan.world)
    extra = []
    for ti in self.plan.split_taskgraphs:       # every rank registers every shard-replica group (collective creation)
      n = graph.taskgraphs[ti].strategy.device_count or self.plan.world
      for k in range(n):
        ranks = [r for r in range(self.plan.world) if r % n == k]
        if ranks not in dp_groups and ranks not in extra:
          extra.append(ranks)
    register_groups(extra, copies=cfg.communication.num_communicators)
    self._sharded: Dict[int, bool] = {}
    self.zero3: Dict[int, Any] = {}
    for s in self.group_keys:
      comm = CollectiveCommunicator("DATA_PARALLEL_GRADS_REDUCE_%d" % s, group_ranks[s], device=self.device)
      self.dp_comms[s] = comm
      self._cur_stage = s
      params = group_params[s]
      self.sharded = (zero in ("v0", "v1", "v2", "v3") or (cfg.communication.fused_kernels and self.device.type == "cuda"
                                                           and not self.baseline)) and comm.size > 1
      self._sharded[s] = self.sharded
      shard_world = comm.size if self.sharded else 1
      # offload.level=v0 keeps the WEIGHTS on the host too (reference graph_editor.py:727-751): that is the per-layer
      # partition / gather-just-in-time engine of ZeRO-3 with its shards in pinned host memory, at any world size (1 included)
      if (zero == "v3" or (cfg.offload.level == "v0" and cfg.offload.weights)) and s < 1000:
        from easyparallellibrary_b200.parallel.zero3 import Zero3Engine
        root = self.stage_modules[s]
        units = sequential_layers(root) or [c for c in root.children()]
        covered = {id(p) for u in units for p in u.parameters()}
        if any(id(p) not in covered for p in root.parameters()):
          units = [root]                         # parameters live directly on the root: treat it as one unit
        self.zero3[s] = Zero3Engine(self, s, units, comm)
        params = []
      allocator = self._bucket_allocator(comm)
      splits = max(cfg.communication.max_splits, cfg.communication.fused_splits) if allocator is not None else cfg.communication.max_splits
      flat = FlatParameters(params, splits, shard_world, allocator=allocator)
      self.flats[s] = flat
      if comm.size > 1:
        for dt, buf in flat.flat_params.items():
          comm.broadcast(buf, root=0)
        if s < 1000:
          for b in self.stage_modules[s].buffers():
            comm.broadcast(b, root=0)
      opts = []
      offload = cfg.offload.level == "v0"
      for b in flat.buckets:
        lo, hi = b.shard_range(comm.rank if self.sharded else 0, shard_world)
        if b.dtype == torch.float32 and not offload:
          master = b.flat_param[lo:hi]                     # fp32 weights are their own master copy
        else:
          master = b.flat_param[lo:hi].to(torch.float32)
        mask = flat.decay_mask(b, self.no_decay)
        if offload:
          from easyparallellibrary_b200.runtime.offload import OffloadedOptimizer
          opts.append(OffloadedOptimizer(self.opt_kind, self.hyper, master, None if mask is None else mask[lo:hi], self.device))
        else:
          opts.append(FlatOptimizer(self.opt_kind, self.hyper, master, None if mask is None else mask[lo:hi]))
      self.optimizers[s] = opts
    for s_, plist in list(self._sparse.items()):
      items = []
      for p in plist:
        if self.dp_comms[s_].size > 1:
          self.dp_comms[s_].broadcast(p.data, root=0)
        master = p.data.view(-1) if p.dtype == torch.float32 else p.data.view(-1).to(torch.float32)
        mask = None if not self.no_decay(p) else torch.zeros(p.numel(), dtype=torch.float32, device=p.device)
        items.append((p, FlatOptimizer(self.opt_kind, self.hyper, master, mask)))
      self._sparse[s_] = items
    self._setup_fused(cfg)
    self._install_grad_hooks()
    if self.plan.pipeline:
      from easyparallellibrary_b200.parallel.pipeline import PipelineExecutor
      self.pipe = PipelineExecutor(self)
    self._graphed = None
    self._capturing = False
    import os as _os
    want = self.cuda_graph if self.cuda_graph is not None else (_os.environ.get("EPL_CUDA_GRAPH", "0") == "1")
    if want and self.device.type == "cuda":
      from easyparallellibrary_b200.parallel.graph_step import GraphedStep
      if GraphedStep.eligible(self):
        self._graphed = GraphedStep(self)
      else:
        get_logger().info("cuda_graph requested but this configuration is not capturable; running eagerly")
    self._built = True
    return self

  def _bucket_allocator(self, comm):
    """Flat weight / gradient buffers live in NVLink symmetric memory when the fused data-parallel kernel
    will run (same allocation order on every rank, so the buffers pair up)."""
    from easyparallellibrary_b200.parallel.fused_dp import FusedDataParallel
    if not FusedDataParallel.eligible(self, comm):
      return None
    from easyparallellibrary_b200.runtime.symmetric import SymmetricBuffer
    if not hasattr(self, "_symm_buffers"):
      self._symm_buffers = {}
    stage = len(self.dp_comms) - 1
    state = {"n": 0}

    def alloc(numel, dtype, device):
      kind = "param" if state["n"] % 2 == 0 else "grad"
      state["n"] += 1
      if dtype not in (torch.bfloat16, torch.float16):
        return torch.zeros(numel, dtype=dtype, device=device)
      pg = getattr(comm.primary, "group", None)
      buf = None
      import os as _os
      if _os.environ.get("EPL_K1", "v2") == "nvls" and pg is not None:
        # NVLS: allocate the bucket through the VMM + multicast path (runtime/nvls.py); the peer pointers of the same allocation
        # keep the v2 kernel usable when the platform has no multicast
        try:
          from easyparallellibrary_b200.runtime.nvls import MulticastBuffer
          buf = MulticastBuffer(numel * 2, device, pg)
        except Exception as e:      # pragma: no cover - platform dependent
          get_logger().warning("NVLS bucket allocation failed (%s); using cudaIpc symmetric memory", e)
          buf = None
      if buf is None:
        buf = SymmetricBuffer(numel * 2, comm.ranks, device, group=pg)
      self._symm_buffers[(self._cur_stage, kind, dtype)] = buf
      return buf.tensor(dtype, numel)
    return alloc

  def _setup_fused(self, cfg) -> None:
    """Fused reduce-scatter + Adam + all-gather over NVLink peer memory (K1)."""
    self.fused = None
    if self.device.type != "cuda" or self.baseline or not cfg.communication.fused_kernels:
      return
    try:
      from easyparallellibrary_b200.parallel.fused_dp import FusedDataParallel
    except ImportError:
      return
    self.fused = FusedDataParallel.maybe_create(self)

  def _auto_stages(self, graph: Graph) -> None:
    """auto.auto_parallel: cut the sequential layers into pipeline.num_stages taskgraphs."""
    from easyparallellibrary_b200.ir.capture import trace_module_costs
    from easyparallellibrary_b200.ir.node import Node
    from easyparallellibrary_b200.parallel.planner import AutoStageGenerator
    from easyparallellibrary_b200.strategies.base import Replicate
    layers = sequential_layers(self.model)
    if layers is None:
      raise RuntimeError("auto.auto_parallel needs a sequential model")
    n_stages = self.config.pipeline.num_stages
    nodes = None
    if self.example_inputs is not None:
      try:
        traced = trace_module_costs(self.model, self.example_inputs)
        owner = {}
        for li, l in enumerate(layers):
          for m in l.modules():
            owner[id(m)] = li
        per_layer: Dict[int, Node] = {}
        for n in traced:
          li = owner.get(id(n.module))
          if li is None:
            continue
          agg = per_layer.setdefault(li, Node(name="layer.%d" % li, type=type(layers[li]).__name__))
          agg.flops += n.flops
          agg.param_count += n.param_count
          agg.act_bytes += n.act_bytes
        nodes = [per_layer.get(i, Node(name="layer.%d" % i, type=type(layers[i]).__name__)) for i in range(len(layers))]
      except Exception as e:  # pragma: no cover - tracing is best effort
        get_logger().warning("auto stage tracing failed (%s); falling back to parameter counts", e)
    if nodes is None:
      nodes = [Node(name="layer.%d" % i, type=type(l).__name__, param_count=sum(p.numel() for p in l.parameters()))
               for i, l in enumerate(layers)]
    stages = AutoStageGenerator(num_stages=n_stages).search(nodes)
    graph._taskgraphs = []
    graph._by_strategy = {}
    graph._param_tg.clear()
    graph._module_tg.clear()
    for s, st in enumerate(stages):
      tg = graph.new_taskgraph(Replicate(1, name="auto_stage_%d" % s))
      for n in st:
        li = int(n.name.split(".")[1])
        graph._module_tg[layers[li]] = tg.index
        tg.add_module(layers[li])
        for p in layers[li].parameters():
          if p not in graph._param_tg:
            graph._param_tg[p] = tg.index
            tg.add_parameter(p)

  def _assign_layers(self, layers: List[nn.Module], graph: Graph) -> List[int]:
    stage_of_tg = {ti: s for s, ti in enumerate(self.plan.stage_taskgraphs)}
    assign, cur = [], 0
    for l in layers:
      tg = graph.taskgraph_of(l)
      if tg is not None and tg.index in stage_of_tg:
        cur = max(cur, stage_of_tg[tg.index])     # stages are monotone in call order
      assign.append(cur)
    return assign

  # ================================================================== gradient hooks
  def _install_grad_hooks(self) -> None:
    self._bucket_of: Dict[int, Tuple[int, Bucket, int]] = {}
    self._grad_view: Dict[int, Tuple[torch.Tensor, int]] = {}      # parameter -> (its view of the flat gradient, data_ptr)
    self._overlap = False
    self._first_micro_batch = True
    self._last_micro_batch = True
    self._mean = True
    # a parameter referenced from several places (tied embeddings: the LM head multiplies by wte.weight) receives several
    # gradient contributions per backward; such a parameter is "ready" only when autograd has accumulated the last one
    uses: Dict[int, int] = {}
    for m in self.stage_modules.values():
      for _, p in m.named_parameters(remove_duplicate=False):
        uses[id(p)] = uses.get(id(p), 0) + 1
    self._sink_params: List[nn.Parameter] = []
    for s, flat in self.flats.items():
      for b in flat.buckets:
        b.ready_ids = set()
        for p, o in zip(b.params, b.offsets):
          self._bucket_of[id(p)] = (s, b, o)
          view = b.flat_grad[o:o + p.numel()].view(p.shape)      # the buckets are persistent: build the view once, not per hook call
          self._grad_view[id(p)] = (view, view.data_ptr())
          p.register_post_accumulate_grad_hook(self._on_grad_ready)
          if p.is_cuda and b.flat_grad.dtype == p.dtype:
            # weight-gradient GEMMs accumulate straight into the flat bucket (ops/linear.py:_sink_weight_grad)
            p.epl_main_grad = b.flat_grad[o:o + p.numel()]
            if uses.get(id(p), 1) <= 1 and not getattr(p, "epl_shared", False):
              # exactly one contribution: the GEMM itself reports readiness (autograd sees no gradient for this weight) and
              # the first micro-batch's GEMM may store instead of accumulate (epl_sink_fresh, re-armed every step)
              p.epl_grad_ready = self._on_grad_ready
              self._sink_params.append(p)
    self._pending: List[Tuple[int, Bucket, Any]] = []
    self._launched_buckets = set()

  def _on_grad_ready(self, p: nn.Parameter) -> None:
    s, b, o = self._bucket_of[id(p)]
    g = p.grad
    if g is not None:
      view, ptr = self._grad_view[id(p)]
      if g.data_ptr() != ptr:                        # autograd replaced the view: fold it back
        view.add_(g)
        p.grad = view if view.dtype == p.dtype else None
    b.ready += 1
    b.ready_ids.add(id(p))
    if len(b.ready_ids) == len(b.params) and not self._last_micro_batch:
      b.ready_ids = set()                          # complete for this micro-batch: count again for the next one
    elif len(b.ready_ids) == len(b.params) and (s, b.index) not in self._launched_buckets:
      if self.fused is not None and self.fused.overlap and not self.plan.pipeline and self.max_grad_norm is None:
        self._launched_buckets.add((s, b.index))
        self.fused.launch_bucket_async(s, b.index)
      elif self._overlap:
        self._launched_buckets.add((s, b.index))
        self._launch_bucket_reduce(s, b)

  @property
  def lr(self) -> float:
    """Learning rate of the next step (assign to change it; with a schedule the schedule wins)."""
    return self.hyper.lr

  @lr.setter
  def lr(self, value: float) -> None:
    self.hyper.lr = float(value)

  # ================================================================== one step
  def step(self, *batch, **kwargs) -> StepOutput:
    if not self._built:
      self.build()
    for h in self.hooks:
      h.before_step(self)
    if self.lr_schedule is not None:
      self.hyper.lr = float(self.lr_schedule(self.global_step))
    cfg = self.config
    mean = cfg.communication.gradients_reduce_method == constant.REDUCE_MEAN
    self._mean = mean
    graph = Graph.get()
    batch = tuple(_to_device(x, self.device) for x in batch)
    if self.compute_dtype is not None and batch and isinstance(batch[0], torch.Tensor) and batch[0].is_floating_point():
      batch = (batch[0].to(self.compute_dtype),) + batch[1:]          # the model input follows the compute dtype (AMP)
    graphed_loss = self._graphed.step(batch, kwargs) if self._graphed is not None else None
    if graphed_loss is not None:                   # the whole step was one CUDA-graph replay (parallel/graph_step.py)
      self.global_step += 1
      out = StepOutput(skipped=False, grad_norm=None, loss_scale=self.scaler.loss_scale, loss=graphed_loss.float().clone())
    else:
      if self._graphed is not None:
        self._graphed._refresh(mean, count=False)
      losses, collected, skipped, gnorm = self._run_step(batch, kwargs, in_graph=False)
      self.global_step += 0 if skipped else 1
      out = StepOutput(skipped=skipped, grad_norm=gnorm, loss_scale=self.scaler.loss_scale)
      if losses:
        out.loss = torch.stack([l.float() for l in losses]).mean() if mean else torch.stack([l.float() for l in losses]).sum()
      out.collections = self._merge_collections(collected)
    graph.current_micro_batch = None
    for h in self.hooks:
      h.after_step(self, out)
    return out

  def _eager_body(self, batch, kwargs, in_graph: bool = True) -> torch.Tensor:
    """One full step on ``batch`` (already on the device); used by the CUDA-graph capture.  Returns the loss tensor."""
    losses, collected, skipped, _ = self._run_step(batch, kwargs, in_graph=in_graph)
    if any(collected):
      raise RuntimeError("collections are filled on the host and cannot be replayed from a CUDA graph")
    return losses[0] if len(losses) == 1 else torch.stack([l.float() for l in losses]).mean()

  def _run_step(self, batch, kwargs, in_graph: bool):
    """Zero the gradient buckets, forward + backward over the micro-batches (or the pipeline program), reduce + apply."""
    cfg = self.config
    M = cfg.pipeline.num_micro_batch
    mean = self._mean
    graph = Graph.get()
    graph.pop_collections()
    for flat in self.flats.values():
      flat.zero_grad()
    for z in self.zero3.values():
      z.zero_grad()
    self._pending = []
    self._launched_buckets = set()
    for flat in self.flats.values():
      for b in flat.buckets:
        b.ready_ids = set()
    for p in self._sink_params:
      p.epl_sink_fresh = True
    if self.fused is not None:
      self.fused.reset_step()
      if in_graph:
        self.fused._prepared = True                # the host refreshed the device-side step values before the capture / replay
      else:
        self.fused.begin_step(mean)
    micro = _split_batch(batch, M)
    losses: List[torch.Tensor] = []
    collected: List["OrderedDict[str, List[Any]]"] = []
    dp_size = max(c.size for c in self.dp_comms.values()) if self.dp_comms else 1
    self._overlap = (dp_size > 1 and not cfg.communication.clip_after_allreduce and self.max_grad_norm is None
                     and self.fused is None)
    if cfg.gradient_checkpoint.check_gradients and cfg.gradient_checkpoint.type and not getattr(self, "_gc_checked", False):
      self._gc_checked = True
      self._check_recompute_gradients(micro[0], kwargs)
    with amp_lib.o1_autocast(self.device.type, enabled=self.o1, debug_log=cfg.amp.debug_log):
      if self.plan.pipeline:
        losses, collected = self.pipe.run(micro, mean)
      else:
        for i, mb in enumerate(micro):
          self._first_micro_batch = i == 0
          self._last_micro_batch = i == M - 1
          graph.current_micro_batch = mb
          with phase_scope(ModelPhase.FORWARD):
            loss = self._forward_loss(mb, kwargs)
          collected.append(graph.pop_collections())
          losses.append(loss.detach())
          scaled = self.scaler.scale(loss)
          if mean and M > 1:
            scaled = scaled / M
          with phase_scope(ModelPhase.BACKWARD):
            scaled.backward()
    for z in self.zero3.values():
      z.finish_backward()
    with phase_scope(ModelPhase.APPLY):
      skipped, gnorm = self._reduce_and_apply(mean)
    return losses, collected, skipped, gnorm

  def _check_recompute_gradients(self, mb: Tuple[Any, ...], kwargs) -> None:
    """``gradient_checkpoint.check_gradients`` (reference gc/gradient_checkpoint.py:310-325): before the first step, the
    gradients of one micro-batch are computed with and without recomputation and compared; a mismatch (a segment that is not
    a pure function of its inputs: unseeded randomness, in-place state) raises instead of training on wrong gradients."""
    from easyparallellibrary_b200.runtime import gradient_checkpoint as gc_lib
    if self.plan.pipeline or self.zero3:
      get_logger().warning("gradient_checkpoint.check_gradients: skipped (pipeline / ZeRO-3 stages own their backward)")
      return
    saved_last, self._last_micro_batch = self._last_micro_batch, False      # no bucket may be launched from the hooks
    snaps = []
    try:
      for enabled in (False, True):
        gc_lib.RECOMPUTE_ENABLED = enabled
        for flat in self.flats.values():
          flat.zero_grad()
        for p in self._sink_params:
          p.epl_sink_fresh = True
        with amp_lib.o1_autocast(self.device.type, enabled=self.o1):
          self._forward_loss(mb, kwargs).backward()
        snaps.append([g.detach().float().clone() for flat in self.flats.values() for g in flat.flat_grads.values()])
    finally:
      gc_lib.RECOMPUTE_ENABLED = True
      self._last_micro_batch = saved_last
    worst = 0.0
    for a, b in zip(*snaps):
      worst = max(worst, float((a - b).abs().max() / (a.abs().max() + 1e-12)))
    tol = 1e-5 if self.compute_dtype is None and not self.o1 else 2e-2
    get_logger().info("gradient checkpoint check: max relative gradient difference %.3e (tolerance %.1e)", worst, tol)
    if not worst <= tol:
      raise RuntimeError("gradient_checkpoint.check_gradients: recomputed gradients differ from plain ones (max relative "
                         "difference %.3e > %.1e)" % (worst, tol))
    for flat in self.flats.values():
      flat.zero_grad()
      for b in flat.buckets:
        b.ready_ids = set()
    for p in self._sink_params:
      p.epl_sink_fresh = True
    self.gc_check_result = worst

  def _forward_loss(self, mb: Tuple[Any, ...], kwargs) -> torch.Tensor:
    if self.plan.num_stages > 1:          # colocated stages: run them back to back
      x = mb[0]
      for s in range(self.plan.num_stages):
        x = self.stage_modules[s](x)
      out = x
      if self.loss_fn is not None and (len(mb) > 1 or torch.is_grad_enabled()):
        return self.loss_fn(out, *mb[1:])
      return out if not torch.is_grad_enabled() else _as_loss(out)       # eval_step(inputs) without labels: the model's output
    if self.loss_fn is not None:
      out = self.model(mb[0], **kwargs)
      if len(mb) == 1 and not torch.is_grad_enabled():
        return out                                                       # eval_step(inputs) without labels: the model's output
      return self.loss_fn(out, *mb[1:])
    return _as_loss(self.model(*mb, **kwargs))

  # ------------------------------------------------------------------ reduce + apply
  def _launch_bucket_reduce(self, s: int, b: Bucket) -> None:
    comm = self.dp_comms[s]
    if comm.size <= 1:
      return
    slot = len(self._pending) % comm.pool.size
    be = comm.pool.backends[slot]
    zero = self.config.zero.level
    ccfg = self.config.communication
    if ccfg.fp16 and b.flat_grad.dtype == torch.float32:
      # 16-bit wire format for fp32 gradients (reference coalescing.py:341-342,374-378: cast(grad * fp16_scale) before the
      # all-reduce, cast back and divide after): halves the bytes on the link; the decompression runs when the reduce is waited on
      wire = (b.flat_grad * float(ccfg.fp16_scale)).to(torch.float16)
      w = be.all_reduce_async(wire, "sum")
      self._pending.append((s, b, _Decompress(w, wire, b.flat_grad, float(ccfg.fp16_scale))))
      return
    if self._sharded[s] and zero != "v0":
      lo, hi = b.shard_range(comm.rank, comm.size)
      w = be.reduce_scatter_into(b.flat_grad[lo:hi], b.flat_grad, "sum", async_op=True)
    else:
      w = be.all_reduce_async(b.flat_grad, "sum")
    self._pending.append((s, b, w))

  def _reduce_and_apply(self, mean: bool) -> Tuple[bool, Optional[torch.Tensor]]:
    cfg = self.config
    clip_after = cfg.communication.clip_after_allreduce
    inv = self.scaler.inv_scale
    gnorm = None
    # (1) clip-then-reduce: local norm, local clip coefficient folded into the grad scale
    local_coef = 1.0
    local_coef_applied = 1.0            # sparse gradients are not stored in the buckets: their clip coefficient rides on the scale
    if self.max_grad_norm is not None and not clip_after:
      gnorm = self._grad_norm(reduced=False) * inv
      c = float(min(1.0, self.max_grad_norm / (float(gnorm) + 1e-6)))
      local_coef_applied = c
      if c < 1.0:                       # each replica clips its own gradient, then the clipped ones are reduced
        for s_ in self.group_keys:
          for g_ in self.flats[s_].flat_grads.values():
            g_.mul_(c)
          if s_ in self.zero3:           # ZeRO-3 gradients live in the units' shards (already reduce-scattered)
            self.zero3[s_].scale_grads(c)
    if self.fused is not None:          # the gradient buckets live in symmetric memory: the fused kernel reads the clipped values
      skipped, _ = self.fused.reduce_and_apply(mean)
      self._reduce_apply_sparse(mean, inv * local_coef_applied)
      return skipped, gnorm
    # (2) reduce, last bucket first (its gradients were produced first)
    launched = {(s, b.index) for s, b, _ in self._pending}
    for s in self.group_keys:
      for b in reversed(self.flats[s].buckets):
        if (s, b.index) not in launched:
          self._launch_bucket_reduce(s, b)
    for _, _, w in self._pending:
      if w is not None:
        w.wait()
    # (3) finite check (fp16 only), agreed on by every replica
    found_inf = False
    if isinstance(self.scaler, (amp_lib.DynamicLossScale, amp_lib.FixedLossScale)):
      bad = torch.zeros(1, device=self.device)
      for s in self.group_keys:
        for b in self.flats[s].buckets:
          bad += (~torch.isfinite(b.flat_grad)).any().float()
        if s in self.zero3:
          bad += self.zero3[s].has_non_finite()
      for comm in self.dp_comms.values():
        if comm.size > 1:
          comm.primary.all_reduce(bad, "max")
      if self.plan.pipeline:                   # ... and by every stage: a step is skipped everywhere or nowhere
        self.pipe.all_reduce_over_stages(bad)
      found_inf = bool(bad.item() > 0)
    if self.scaler.update(found_inf):
      return True, gnorm
    # (4) reduce-then-clip
    coef = local_coef
    if self.max_grad_norm is not None and clip_after:
      n = max(self.mean_divisor(s_) for s_ in self.group_keys)
      gnorm = self._grad_norm(reduced=True) * inv / (n if mean else 1)
      coef = float(min(1.0, self.max_grad_norm / (float(gnorm) + 1e-6)))
    # (5) apply
    for s in self.group_keys:
      self._apply_group(s, mean, inv * coef)
    self._reduce_apply_sparse(mean, inv * coef * local_coef_applied)
    return False, gnorm

  def _reduce_apply_sparse(self, mean: bool, scale0: float) -> None:
    if not self._sparse:
      return
    from easyparallellibrary_b200.communicators.sparse import sparse_all_reduce
    densify = self.config.communication.sparse_as_dense
    for s, items in self._sparse.items():
      comm = self.dp_comms[s]
      scale = scale0 / (self.mean_divisor(s) if mean else 1)
      for p, opt in items:
        g = p.grad
        if g is None:
          continue
        if g.is_sparse:
          if densify:
            g = g.to_dense()
            if comm.size > 1:
              comm.primary.all_reduce(g, "sum")
          else:
            if comm.size > 1:
              g = sparse_all_reduce(comm, g)            # all-gather of indices and values: bytes ~ touched rows, not the table
            self.sparse_wire_elems = getattr(self, "sparse_wire_elems", 0) + int(g._nnz()) * int(g.shape[1] if g.dim() > 1 else 1)
            g = g.to_dense()
        elif comm.size > 1:
          comm.primary.all_reduce(g, "sum")
        opt.step(g.reshape(-1), p.data.view(-1), scale)
        p.grad = None

  def mean_divisor(self, s: int) -> int:
    """Number of independent data replicas whose (already batch-averaged) losses this group's reduction sums.

    Plain data parallelism: the group size.  Megatron-style tensor parallelism (``split(n)`` on k*n GPUs, the n ranks of a
    group see the same batch): still the group size — shard groups have k members, one per data replica; replicated
    parameters are reduced over all k*n ranks and every replica's gradient appears n times in that sum.  ``Replica2Split``
    bridges (the reference's colocated split head, ``bridging_layer.py:46-58``): every rank feeds its own batch, the n
    batches are gathered and the loss is the mean over the gathered batch, so the sum over a group already *is* the
    gradient of that mean and only the k groups are averaged."""
    comm = self.dp_comms[s]
    if not self.has_split:
      return comm.size
    if self._split_gathers_batch:
      return max(self.plan.world // max(self._split_size, 1), 1)
    return comm.size

  def _apply_group(self, s: int, mean: bool, scale0: float, only: Optional[Sequence[int]] = None) -> None:
    cfg = self.config
    comm, flat, opts = self.dp_comms[s], self.flats[s], self.optimizers[s]
    sharded = self._sharded[s]
    scale = scale0 / (self.mean_divisor(s) if mean else 1)
    if s in self.zero3:
      self.zero3[s].apply(scale)
    groups = max(1, cfg.optimizer.num_apply_group)
    gathers = []
    for b, opt in zip(flat.buckets, opts):
      if only is not None and b.index not in only:
        continue
      lo, hi = b.shard_range(comm.rank if sharded else 0, comm.size if sharded else 1)
      gshard, pshard = b.flat_grad[lo:hi], b.flat_param[lo:hi]
      if self.baseline:
        self._baseline_apply(opt, gshard, pshard, scale)
      else:
        n = hi - lo
        per = ((n + groups - 1) // groups + 7) // 8 * 8
        for g in range(groups - 1, -1, -1):           # groups applied last to first (optimizer_helper.py:95-120)
          opt.step(gshard, pshard, scale, min(g * per, n), min((g + 1) * per, n), count_step=(g == groups - 1))
      if sharded and comm.size > 1:
        gathers.append(comm.pool.backends[len(gathers) % comm.pool.size].all_gather_into(b.flat_param, pshard, async_op=True))
    for w in gathers:
      if w is not None:
        w.wait()

  def _apply_group_library(self, s: int, mean: bool, only: Optional[Sequence[int]] = None) -> None:
    """Reduce + apply one parameter group (or the buckets ``only`` of it) through the library path: used by the fused engine
    for what does not live in symmetric memory — fp32 buckets, single-replica groups."""
    for b in reversed(self.flats[s].buckets):
      if only is None or b.index in only:
        self._launch_bucket_reduce(s, b)
    for _, _, w in self._pending:
      if w is not None:
        w.wait()
    self._pending = []
    self._apply_group(s, mean, self.scaler.inv_scale, only=only)

  def _baseline_apply(self, opt: FlatOptimizer, g, p, scale) -> None:
    from easyparallellibrary_b200.runtime.optimizer import adamw_reference, sgd_reference
    if opt.kind == "torch":
      opt.step(g, p, scale)
      return
    opt.step_count += 1
    out = p if p.data_ptr() != opt.master.data_ptr() else None
    if opt.kind == "sgd":
      sgd_reference(opt.master, g, opt.m, opt.hyper, scale, out)
    else:
      adamw_reference(opt.master, g, opt.m, opt.v, opt.step_count, opt.hyper, scale, opt.decay_mask, out)

  def _grad_norm(self, reduced: bool) -> torch.Tensor:
    sq = torch.zeros(1, device=self.device, dtype=torch.float32)
    for s in self.group_keys:
      comm = self.dp_comms[s]
      part = torch.zeros(1, device=self.device, dtype=torch.float32)
      for b in self.flats[s].buckets:
        if reduced and self._sharded[s] and self.config.zero.level != "v0" and comm.size > 1:
          lo, hi = b.shard_range(comm.rank, comm.size)
          part += b.flat_grad[lo:hi].float().pow(2).sum()
        else:
          part += b.flat_grad.float().pow(2).sum()
      if reduced and self._sharded[s] and self.config.zero.level != "v0" and comm.size > 1:
        comm.primary.all_reduce(part, "sum")
      if s >= 1000:                     # tensor-parallel shards: every rank of the TP group holds a different slice
        tp = self._tp_comm(s - 1000)
        if tp is not None and tp.size > 1:
          tp.primary.all_reduce(part, "sum")
      sq += part
      if s in self.zero3:
        sq += self.zero3[s].grad_sq_norm(local=not reduced, divisor=comm.size if (self._mean and not reduced) else 1)
    for s, items in self._sparse.items():
      for p, _ in items:
        if p.grad is not None:
          g = p.grad.coalesce().values() if p.grad.is_sparse else p.grad
          part = g.float().pow(2).sum().reshape(1)
          if reduced and self.dp_comms[s].size > 1:          # not reduced yet at this point: sum of the replicas' squared norms
            self.dp_comms[s].primary.all_reduce(part, "sum")
          sq += part
    if self.plan.pipeline and self.plan.num_stages > 1:
      sq = self.pipe.all_reduce_over_stages(sq)
    return sq.sqrt()

  def _tp_comm(self, ti: int):
    """Communicator over the tensor-parallel group of split taskgraph ``ti`` (the one the split ops themselves use)."""
    try:
      from easyparallellibrary_b200.ops.tensor_parallel import current_tp_group
      return current_tp_group(Graph.get().taskgraphs[ti].strategy).comm
    except Exception:       # pragma: no cover - no TP group (single device)
      return None

  # ------------------------------------------------------------------ collections
  def _merge_collections(self, collected: List["OrderedDict[str, List[Any]]"]) -> "OrderedDict[str, List[Any]]":
    """LOCAL_*: merge over micro-batches; GLOBAL_*: additionally over replicas."""
    merged: "OrderedDict[str, List[Any]]" = OrderedDict()
    if not collected or not any(collected):
      return merged
    keys = [k for k in GraphKeys.ALL_COLLECTION_KEYS if any(k in c for c in collected)]
    comm = next(iter(self.dp_comms.values())) if self.dp_comms else None
    for k in keys:
      per_mb = [c[k] for c in collected if k in c]
      n_obj = min(len(x) for x in per_mb)
      outs = []
      for j in range(n_obj):
        vals = [torch.as_tensor(x[j]).detach() for x in per_mb]
        vals = [v.to(self.device) for v in vals]
        if "concat" in k:
          v = torch.cat([t.reshape(1) if t.dim() == 0 else t for t in vals], 0)
        elif "mean" in k:
          v = torch.stack([t.float() for t in vals]).mean(0)
        else:
          v = torch.stack(vals).sum(0)
        if k.startswith("global") and comm is not None and comm.size > 1:
          if "concat" in k:
            v = comm.allgather(v.contiguous())
          else:
            v = comm.allreduce(v.contiguous(), mean="mean" in k)
        outs.append(v)
      merged[k] = outs
    return merged

  # ------------------------------------------------------------------ evaluation & state
  @torch.no_grad()
  def eval_step(self, *batch, **kwargs):
    """Evaluation is not parallelised (reference ``ir/graph.py:926-933``): plain forward."""
    if not self._built:
      self.build()
    batch = tuple(_to_device(x, self.device) for x in batch)
    if self.compute_dtype is not None and batch and isinstance(batch[0], torch.Tensor) and batch[0].is_floating_point():
      batch = (batch[0].to(self.compute_dtype),) + batch[1:]          # same input cast as step() (AMP)
    was = self.model.training
    self.model.eval()
    Graph.get().current_micro_batch = batch
    try:
      with amp_lib.o1_autocast(self.device.type, enabled=self.o1):       # same per-op precision (and stage-boundary dtypes) as training
        if self.plan.pipeline:
          return self.pipe.forward_only(batch)
        return self._forward_loss(batch, kwargs)
    finally:
      self.model.train(was)
      for z in self.zero3.values():
        z.release_all()

  def state_dict(self) -> Dict[str, Any]:
    sd = {"global_step": self.global_step, "model": {}, "optim": {}, "loss_scale": self.scaler.loss_scale}
    for s in self.plan.local_stages:
      sd["model"][s] = self.stage_modules[s].state_dict()
    for s in self.group_keys:
      sd["optim"][s] = [o.state_dict() for o in self.optimizers[s]]
    return sd

  def load_state_dict(self, sd: Dict[str, Any]) -> None:
    self.global_step = int(sd["global_step"])
    if hasattr(self.scaler, "loss_scale"):
      self.scaler.loss_scale = sd.get("loss_scale", self.scaler.loss_scale)
    for s in self.plan.local_stages:
      self.stage_modules[s].load_state_dict(sd["model"][s])
    for s in self.group_keys:
      for o, osd in zip(self.optimizers[s], sd["optim"][s]):
        o.load_state_dict(osd)

  @property
  def is_first_replica(self) -> bool:
    return all(c.rank == 0 for c in self.dp_comms.values())


class _Decompress(object):
  """Work handle of a 16-bit-compressed bucket all-reduce: ``wait()`` also restores the fp32 gradient."""

  def __init__(self, work, wire: torch.Tensor, dst: torch.Tensor, scale: float):
    self.work, self.wire, self.dst, self.scale = work, wire, dst, scale

  def wait(self) -> None:
    if self.work is not None:
      self.work.wait()
    self.dst.copy_(self.wire.to(torch.float32).div_(self.scale))


def _as_loss(out) -> torch.Tensor:
  if isinstance(out, torch.Tensor):
    return out
  if isinstance(out, dict):
    return out["loss"]
  if isinstance(out, (tuple, list)):
    return out[0]
  raise TypeError("model output %r cannot be interpreted as a loss" % type(out))


def _to_device(x, device):
  if isinstance(x, torch.Tensor) and x.device != device:
    return x.to(device, non_blocking=True)
  return x


def _materialize(module: nn.Module, device: torch.device) -> None:
  """Move to ``device``; parameters created on the meta device are allocated and re-initialised."""
  has_meta = any(p.is_meta for p in module.parameters())
  if has_meta:
    module.to_empty(device=device)
    for m in module.modules():
      reset = getattr(m, "reset_parameters", None)
      if callable(reset):
        reset()
  else:
    module.to(device)


def prepare(model: nn.Module, optimizer: str = "adamw", **kw) -> Trainer:
  """Functional spelling of ``Trainer(...).build()``."""
  return Trainer(model, optimizer, **kw).build()
