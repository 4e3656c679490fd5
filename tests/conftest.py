import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 through gpurun)")


def pytest_collection_modifyitems(config, items):
  import torch
  if torch.cuda.is_available():
    return
  skip = pytest.mark.skip(reason="no CUDA device")
  for item in items:
    if "gpu" in item.keywords:
      item.add_marker(skip)


@pytest.fixture(autouse=True)
def _fresh_env():
  """Every test starts from a clean Env (the reference needs one process per test file because it
  monkey-patches TF globally; here state is explicit and resettable)."""
  import easyparallellibrary_b200 as epl
  for k in [k for k in os.environ if k.startswith("EPL_")]:
    del os.environ[k]
  epl.Env.get().reset()
  yield
  epl.Env.get().reset()
