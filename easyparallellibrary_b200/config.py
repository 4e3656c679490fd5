"""Configuration system.

Two-level ``group.attr`` keys; precedence is *code dict > ``EPL_<GROUP>_<ATTR>``
environment variable > default*; values are type-checked against the default;
unknown keys are rejected.  Key names and defaults follow the reference
(``epl/config.py:55-178``, doc table ``docs/en/api/config.md``) so a user's
``epl.Config({...})`` dictionary carries over unchanged.

Unlike the reference (one Python class per group with class attributes parsed
by ``inspect``) the schema here is a single declarative table, which is also
what the doc-sync test and ``Config.describe()`` iterate over.
"""
from __future__ import annotations

import os
from collections import OrderedDict
from typing import Any, Dict, Iterator, Tuple

from easyparallellibrary_b200.utils import constant

# (group, attr) -> (default, help)
_SCHEMA: "OrderedDict[str, OrderedDict[str, Tuple[Any, str]]]" = OrderedDict()


def _opt(group: str, attr: str, default: Any, doc: str) -> None:
  _SCHEMA.setdefault(group, OrderedDict())[attr] = (default, doc)


_opt("auto", "auto_parallel", False, "Search pipeline stages automatically.")
_opt("io", "drop_last_files", False, "Drop trailing files so every worker gets the same number.")
_opt("io", "unbalanced_io_slicing", False, "Allow workers to receive different numbers of files.")
_opt("io", "slicing", False, "Shard the input file list across workers automatically.")
_opt("communication", "sparse_as_dense", False, "Densify sparse gradients before reduction.")
_opt("communication", "max_splits", 5, "Maximum number of fused gradient buckets.")
_opt("communication", "num_communicators", 2, "Communicators (and side streams) per pool.")
_opt("communication", "fp16", False, "Compress gradients to 16 bit on the wire.")
_opt("communication", "fp16_scale", 128, "Scale applied before 16-bit wire compression.")
_opt("communication", "clip_after_allreduce", False, "Clip gradients after the reduction instead of before.")
_opt("communication", "gradients_reduce_method", constant.REDUCE_MEAN, "mean or sum over replicas and micro-batches.")
_opt("communication", "fused_kernels", True,
     "B200 extension: use in-kernel NVLink (P2P / multimem) fused collectives where available.")
_opt("communication", "fused_splits", 16,
     "B200 extension: number of gradient buckets when the fused reduce-scatter+AdamW+all-gather kernel runs (more, smaller "
     "buckets overlap better with backward; only the last one is exposed).")
_opt("pipeline", "num_stages", -1, "Number of stages for automatic partitioning.")
_opt("pipeline", "num_micro_batch", 1, "Micro-batches per step (pipeline depth or accumulation count).")
_opt("pipeline", "strategy", constant.DEFAULT_PIPELINE_STRATEGY,
     "PreferForward | PreferBackward | PreferBackwardOptimizer.")
_opt("gradient_checkpoint", "type", "", "'' | collection | auto.")
_opt("gradient_checkpoint", "end_taskgraph", -1, "Last taskgraph that auto checkpointing may touch.")
_opt("gradient_checkpoint", "check_gradients", False, "Validate recompute gradients against plain ones.")
_opt("zero", "level", "", "'' | v0 | v1 | v2 | v3.")
_opt("zero", "fused_gather", False,
     "B200 extension (v3): gather the weight of a layer's first GEMM inside that GEMM instead of before it.")
_opt("offload", "level", "", "'' | v0 (weights and optimizer state live on the host).")
_opt("offload", "weights", True,
     "B200 extension (v0): False keeps the weights on the device and offloads the optimizer state only.")
_opt("amp", "level", "", "'' | O1 (fp16 + loss scale) | bf16 | fp8 (B200 extension: bf16 weights, e4m3 forward GEMMs).")
_opt("amp", "debug_log", False, "Log the precision decision for every module.")
_opt("amp", "loss_scale", "dynamic", "'dynamic' or a fixed number.")
_opt("cluster", "device_place_prefer_intra_node", True, "Keep one model replica inside a node when possible.")
_opt("cluster", "run_visible_devices", "", "Comma separated device ordinals visible to this process.")
_opt("cluster", "colocate_split_and_replicate", False, "Place split and replicate taskgraphs on the same devices.")
_opt("optimizer", "num_apply_group", 1, "Apply the optimizer in this many sequential groups.")


class ConfigGroup(object):
  """One frozen attribute namespace (e.g. ``config.pipeline``)."""

  __slots__ = ("_name", "_values", "_locked")

  def __init__(self, name: str):
    object.__setattr__(self, "_name", name)
    object.__setattr__(self, "_values", OrderedDict((k, v[0]) for k, v in _SCHEMA[name].items()))
    object.__setattr__(self, "_locked", False)

  def __getattr__(self, key: str) -> Any:
    values = object.__getattribute__(self, "_values")
    if key in values:
      return values[key]
    raise AttributeError("config group %r has no attribute %r" % (self._name, key))

  def __setattr__(self, key: str, value: Any) -> None:
    if key not in self._values:
      raise AttributeError("config group %r has no attribute %r" % (self._name, key))
    self._values[key] = value

  def items(self) -> Iterator[Tuple[str, Any]]:
    return iter(self._values.items())

  def __repr__(self) -> str:
    body = "".join("    %s = %r,\n" % kv for kv in self._values.items())
    return "%s {\n%s}" % (self._name, body)


def _coerce_env(raw: str, default: Any, key: str) -> Any:
  if isinstance(default, bool):
    low = raw.lower()
    if low not in ("true", "false"):
      raise ValueError("Unknown bool parameter, key: %s, value: %s" % (key, raw))
    return low == "true"
  if isinstance(default, int):
    return int(raw)
  if isinstance(default, float):
    return float(raw)
  return raw


def _check_type(value: Any, default: Any, key: str) -> Any:
  if isinstance(value, str):
    value = value.lower()
  if default is None:
    return value
  # bool is an int subclass: refuse silently turning True into 1 and vice versa.
  if isinstance(default, bool) != isinstance(value, bool) or not isinstance(value, type(default)):
    raise ValueError("%s type error, expected: %s." % (key, type(default)))
  return value


class Config(object):
  """``Config({"pipeline.num_micro_batch": 4, ...})``."""

  def __init__(self, param_dict: Dict[str, Any] | None = None):
    params = dict(param_dict or {})
    groups = OrderedDict((g, ConfigGroup(g)) for g in _SCHEMA)
    object.__setattr__(self, "_groups", groups)
    known = {"%s.%s" % (g, a) for g in _SCHEMA for a in _SCHEMA[g]}
    unknown = sorted(set(params) - known)
    if unknown:
      raise AttributeError("Unknown config key(s): %s" % ", ".join(unknown))
    for gname, attrs in _SCHEMA.items():
      for attr, (default, _) in attrs.items():
        key = "%s.%s" % (gname, attr)
        value = default
        env_key = (constant.ENV_PREFIX + gname + "_" + attr).upper()
        if env_key in os.environ:
          value = _coerce_env(os.environ[env_key], default, env_key)
        if key in params:
          value = params[key]
        if key == "amp.loss_scale":
          if not (isinstance(value, str) and value.lower() == "dynamic"):
            value = float(value)
          else:
            value = "dynamic"
        else:
          value = _check_type(value, default, key)
        setattr(groups[gname], attr, value)
    self._validate()

  # attribute access: config.pipeline.num_micro_batch
  def __getattr__(self, key: str) -> ConfigGroup:
    groups = object.__getattribute__(self, "_groups")
    if key in groups:
      return groups[key]
    raise AttributeError("Config has no group %r" % key)

  def __setattr__(self, key: str, value: Any) -> None:
    raise AttributeError("Config groups are fixed; set config.<group>.<attr> instead")

  def _validate(self) -> None:
    if self.communication.gradients_reduce_method not in constant.REDUCE_METHODS:
      raise ValueError("Gradients reduce method error: %s, which should be one of %s."
                       % (self.communication.gradients_reduce_method, list(constant.REDUCE_METHODS)))
    if self.zero.level not in constant.ZERO_LEVELS:
      raise ValueError("zero.level must be one of %s" % (constant.ZERO_LEVELS,))
    if self.offload.level not in constant.OFFLOAD_LEVELS:
      raise ValueError("offload.level must be one of %s" % (constant.OFFLOAD_LEVELS,))
    if self.amp.level not in constant.AMP_LEVELS:
      raise ValueError("amp.level must be one of %s" % (constant.AMP_LEVELS,))
    if self.gradient_checkpoint.type not in ("", constant.GC_COLLECTION, constant.GC_AUTO):
      raise ValueError("gradient_checkpoint.type must be '', 'collection' or 'auto'")
    if self.pipeline.num_micro_batch < 1:
      raise ValueError("pipeline.num_micro_batch must be >= 1")
    if self.communication.max_splits < 1 or self.communication.num_communicators < 1:
      raise ValueError("communication.max_splits / num_communicators must be >= 1")
    if self.optimizer.num_apply_group < 1:
      raise ValueError("optimizer.num_apply_group must be >= 1")

  def to_dict(self) -> Dict[str, Any]:
    return {"%s.%s" % (g, a): v for g, grp in self._groups.items() for a, v in grp.items()}

  @staticmethod
  def describe() -> Iterator[Tuple[str, Any, str]]:
    """Yield ``(key, default, help)`` for every option (drives the docs test)."""
    for g, attrs in _SCHEMA.items():
      for a, (default, doc) in attrs.items():
        yield "%s.%s" % (g, a), default, doc

  def __repr__(self) -> str:
    return "Config {\n%s}" % "".join("  " + repr(g).replace("\n", "\n  ") + "\n" for g in self._groups.values())
