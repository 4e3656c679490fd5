"""EPL-B200: a Blackwell-native hybrid-parallel training framework with the
capabilities of alibaba/EasyParallelLibrary.

Public surface (reference ``epl/__init__.py:23-55``)::

    import easyparallellibrary_b200 as epl
    epl.init(epl.Config({"pipeline.num_micro_batch": 4}))
    with epl.replicate(device_count=1):
        model = Net()
    trainer = epl.Trainer(model, optimizer="adamw", lr=1e-4, loss_fn=loss_fn)
    out = trainer.step(inputs, labels)
"""
from __future__ import annotations

from easyparallellibrary_b200.utils.version import VERSION
from easyparallellibrary_b200.config import Config
from easyparallellibrary_b200.env import Env
from easyparallellibrary_b200.cluster import Cluster, VirtualDevice, Device
from easyparallellibrary_b200.ir.graph import (Graph, GraphKeys, add_to_collection, get_collection,
                                               get_all_collections, current_micro_batch)
from easyparallellibrary_b200.ir.phase import ModelPhase
from easyparallellibrary_b200.strategies import replicate, split, Replicate, Split

__version__ = VERSION


def init(config=None, init_process_group: bool = True):
  """Reset all state, install the capture hooks, discover the cluster.

  Parity: ``epl.init`` (reference ``epl/__init__.py:38-50``): ``Env.reset`` +
  ``Env.init(config)`` and a ``Cluster`` — with layout ``"all"`` when
  ``cluster.colocate_split_and_replicate`` is set, otherwise lazily laid out
  once the taskgraphs are known.  In addition, under a launcher
  (``RANK``/``WORLD_SIZE`` present) the ``torch.distributed`` process group is
  created here: NCCL when this rank has a GPU, gloo otherwise.
  """
  import os
  env = Env.get()
  env.reset()
  env.init(config)
  if env.config.pipeline.num_micro_batch > 1 or env.config.pipeline.num_stages > 1:
    # Pipeline schedules keep NCCL receives posted on the device; CUDA's lazy module loading would make the first launch of
    # any kernel wait for them (deadlock under 1F1B, see parallel/pipeline.py).  Takes effect if the CUDA context does not
    # exist yet; the executor additionally runs a communication-free warm-up pass of every stage.
    os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
  if env.config.cluster.run_visible_devices:
    # reference cluster.run_visible_devices: "visible devices for the session" (epl/config.py:160-164).  One process per GPU here:
    # the list restricts which physical GPUs this job's processes may use; it must be exported before the CUDA context exists
    import torch
    if torch.cuda.is_available() and torch.cuda.is_initialized():
      from easyparallellibrary_b200.utils.logging import get_logger
      get_logger().warning("cluster.run_visible_devices=%s ignored: the CUDA context already exists", env.config.cluster.run_visible_devices)
    else:
      os.environ["CUDA_VISIBLE_DEVICES"] = str(env.config.cluster.run_visible_devices)
  if init_process_group and int(os.environ.get("WORLD_SIZE", "1")) > 1:
    from easyparallellibrary_b200.runtime.dist import ensure_process_group
    ensure_process_group()
  cfg = env.config
  layout = "all" if cfg.cluster.colocate_split_and_replicate else None
  env.cluster = Cluster(layout=layout, prefer_intra_node=cfg.cluster.device_place_prefer_intra_node)
  return env


def shutdown() -> None:
  """Leave the job cleanly: wait for every rank, then destroy the process group.  Without it the rank that hosts the rendezvous
  store can exit while its peers are still tearing down and they die with a non-zero exit code (seen as a flaky launcher
  test); call it at the end of a training script."""
  try:
    import gc
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
      dist.barrier()
      # drop every cached communicator / process-group handle first: a ProcessGroup object that is still referenced from a
      # module-level cache is destroyed during interpreter finalisation, after the store it depends on, and gloo then calls
      # std::terminate ("terminate called without an active exception")
      from easyparallellibrary_b200.communicators import backend, collective_communicator
      try:                                   # in-tree NCCL communicators (GPU jobs): synchronise and destroy them explicitly
        from easyparallellibrary_b200.communicators import native as _native
        for be in list(getattr(_native, "_LIVE", [])):
          be.close()
      except Exception:
        pass
      collective_communicator._REGISTRY.clear()
      backend.reset_groups()
      Env.get().reset()
      gc.collect()
      dist.destroy_process_group()
  except Exception:
    pass


def set_default_strategy(strategy):
  """Everything created outside an explicit scope belongs to ``strategy``
  (replicate only).  Calling it again opens the next taskgraph — the idiom the
  reference's BERT example uses to cut pipeline stages
  (``examples/bert/modeling.py:829-834``)."""
  Graph.get().set_default_strategy(strategy)


def __getattr__(name):
  # heavy sub-systems are imported lazily so `import easyparallellibrary_b200` stays cheap
  if name in ("Trainer", "Engine"):
    from easyparallellibrary_b200.parallel.engine import Trainer
    return Trainer
  if name == "prepare":
    from easyparallellibrary_b200.parallel.engine import prepare
    return prepare
  if name == "summary":
    from easyparallellibrary_b200.utils import summary
    return summary
  if name in ("train", "evaluate", "train_and_evaluate"):
    from easyparallellibrary_b200.runtime import loop
    return getattr(loop, name)
  if name in ("ops", "models", "runtime", "profiler", "communicators", "parallel", "utils"):
    import importlib
    return importlib.import_module("easyparallellibrary_b200." + name)
  raise AttributeError("module 'easyparallellibrary_b200' has no attribute %r" % name)
