"""Config system: defaults, dict > env > default precedence, strict typing, unknown keys, docs sync.
Mirrors the reference's config_test.py / config_env_test.py."""
import os

import pytest

import easyparallellibrary_b200 as epl
from easyparallellibrary_b200.config import Config


def test_defaults_match_reference_table():
  c = Config()
  assert c.auto.auto_parallel is False
  assert (c.io.drop_last_files, c.io.unbalanced_io_slicing, c.io.slicing) == (False, False, False)
  assert (c.communication.max_splits, c.communication.num_communicators) == (5, 2)
  assert (c.communication.fp16, c.communication.fp16_scale) == (False, 128)
  assert c.communication.gradients_reduce_method == "mean" and c.communication.clip_after_allreduce is False
  assert (c.pipeline.num_stages, c.pipeline.num_micro_batch, c.pipeline.strategy) == (-1, 1, "preferbackward")
  assert (c.gradient_checkpoint.type, c.gradient_checkpoint.end_taskgraph, c.gradient_checkpoint.check_gradients) == ("", -1, False)
  assert c.zero.level == "" and c.offload.level == "" and c.amp.level == "" and c.amp.loss_scale == "dynamic"
  assert c.cluster.device_place_prefer_intra_node is True and c.cluster.colocate_split_and_replicate is False
  assert c.optimizer.num_apply_group == 1


def test_dict_overrides_env_overrides_default(monkeypatch):
  monkeypatch.setenv("EPL_PIPELINE_NUM_MICRO_BATCH", "4")
  monkeypatch.setenv("EPL_COMMUNICATION_FP16", "True")
  monkeypatch.setenv("EPL_ZERO_LEVEL", "v1")
  c = Config()
  assert c.pipeline.num_micro_batch == 4 and c.communication.fp16 is True and c.zero.level == "v1"
  c = Config({"pipeline.num_micro_batch": 8})
  assert c.pipeline.num_micro_batch == 8 and c.communication.fp16 is True


def test_type_checks_and_unknown_keys():
  with pytest.raises(ValueError):
    Config({"pipeline.num_micro_batch": "4"})
  with pytest.raises(ValueError):
    Config({"communication.fp16": 1})
  with pytest.raises(AttributeError):
    Config({"pipeline.unknown": 1})
  with pytest.raises(ValueError):
    Config({"communication.gradients_reduce_method": "max"})
  with pytest.raises(ValueError):
    Config({"zero.level": "v9"})
  c = Config()
  with pytest.raises(AttributeError):
    c.pipeline.bogus = 1
  with pytest.raises(ValueError):
    os.environ["EPL_IO_SLICING"] = "maybe"
    try:
      Config()
    finally:
      del os.environ["EPL_IO_SLICING"]


def test_amp_loss_scale_special_case_and_lowercasing():
  assert Config({"amp.loss_scale": 128}).amp.loss_scale == 128.0
  assert Config({"amp.loss_scale": "dynamic"}).amp.loss_scale == "dynamic"
  assert Config({"amp.level": "O1"}).amp.level == "o1"
  assert Config({"pipeline.strategy": "PreferForward"}).pipeline.strategy == "preferforward"


def test_every_key_is_documented():
  doc = open(os.path.join(os.path.dirname(__file__), "..", "docs", "config.md")).read()
  for key, _default, _help in Config.describe():
    assert key in doc, "config key %s missing from docs/config.md" % key


def test_init_accepts_dict_and_config():
  env = epl.init({"pipeline.num_micro_batch": 2}, init_process_group=False)
  assert env.config.pipeline.num_micro_batch == 2
  env = epl.init(epl.Config({"zero.level": "v0"}), init_process_group=False)
  assert env.config.zero.level == "v0"


def test_documented_paths_exist():
  """docs/coverage.md maps every reference component to a file of this repository: keep the map honest."""
  import os
  import re
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  pkg = os.path.join(root, "easyparallellibrary_b200")
  text = open(os.path.join(root, "docs", "coverage.md")).read()
  missing = []
  for token in re.findall(r"`([^`]+)`", text):
    token = token.split("::")[0].strip()
    if not re.match(r"^[\w./{},\-]+\.(py|cpp|cu|cuh|md|sh)$", token) and not token.endswith("/"):
      continue
    # expand one level of {a,b,c}
    m = re.match(r"^(.*)\{([^}]*)\}(.*)$", token)
    names = [m.group(1) + x + m.group(3) for x in m.group(2).split(",")] if m else [token]
    for n in names:
      cands = [os.path.join(root, n), os.path.join(pkg, n)]
      if not any(os.path.exists(c) for c in cands):
        missing.append(n)
  assert not missing, "paths named in docs/coverage.md do not exist: %s" % missing
