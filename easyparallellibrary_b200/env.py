"""Process-wide state (reference ``epl/env.py:38-183``): the config, the
cluster, the strategy stack, the IR graph, and — new here — the process
groups / native communicators created when a plan is built."""
from __future__ import annotations

from typing import Any, Dict, Optional

from easyparallellibrary_b200.config import Config


class Env(object):
  _instance: Optional["Env"] = None

  def __init__(self):
    self.config: Config = Config()
    self.cluster = None
    self.strategy_context = None
    self.graph = None
    self.parallel_information: Dict[str, Any] = {}
    self.comm_resources: Dict[str, Any] = {}
    self.is_initialized = False

  @classmethod
  def get(cls) -> "Env":
    if cls._instance is None:
      cls._instance = Env()
      from easyparallellibrary_b200.strategies.context import StrategyContext
      cls._instance.strategy_context = StrategyContext()
    return cls._instance

  def reset(self) -> None:
    from easyparallellibrary_b200.strategies.context import StrategyContext
    for res in list(self.comm_resources.values()):
      close = getattr(res, "close", None)
      if callable(close):
        try:
          close()
        except Exception:  # pragma: no cover - best effort during teardown
          pass
    self.__init__()
    self.strategy_context = StrategyContext()
    from easyparallellibrary_b200.communicators import backend, collective_communicator
    backend.reset_groups()
    collective_communicator._REGISTRY.clear()

  def init(self, config=None) -> None:
    from easyparallellibrary_b200.ir import capture
    from easyparallellibrary_b200.ir.graph import Graph
    from easyparallellibrary_b200.utils.logging import get_logger
    if config is None:
      config = Config()
    elif isinstance(config, dict):
      config = Config(config)
    self.config = config
    self.graph = Graph()
    capture.install_hooks()
    self.is_initialized = True
    from easyparallellibrary_b200.utils.version import VERSION
    get_logger().debug("EPL-B200 %s initialised with %r", VERSION, config)

  @property
  def default_graph(self):
    from easyparallellibrary_b200.ir.graph import Graph
    return Graph.get()
