"""The 2-layer MLP of the reference's multi-process tests (``tests/dnn_data_parallel.py``) and BASELINE config #1."""
from torch import nn


class MLP(nn.Sequential):
  def __init__(self, in_features: int = 10, hidden: int = 16, out_features: int = 1, layers: int = 2):
    mods = []
    d = in_features
    for _ in range(layers - 1):
      mods += [nn.Linear(d, hidden), nn.ReLU()]
      d = hidden
    mods.append(nn.Linear(d, out_features))
    super().__init__(*mods)
