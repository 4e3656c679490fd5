"""ResNet-50 with a wide classification head (reference workload: ``examples/resnet/resnet_dp.py:26-66`` — slim
``resnet_v1_50`` + a 10 000-class dense head on synthetic 224x224 images; ``resnet_split.py:49-57`` puts the head
and the loss under ``epl.split``).

Convolutions go through cuDNN (a plain library op, channels-last bf16); the hot path this model exercises is the
data-parallel gradient reduction + optimizer (fused reduce-scatter/AdamW/all-gather kernel) and, in the split
variant, the class-parallel head: all-gather -> tcgen05 GEMM and the one-collective softmax cross-entropy.
"""
from __future__ import annotations

import torch
from torch import nn


class Bottleneck(nn.Module):
  expansion = 4

  def __init__(self, inplanes: int, planes: int, stride: int = 1, downsample=None):
    super().__init__()
    self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
    self.bn1 = nn.BatchNorm2d(planes)
    self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)
    self.bn2 = nn.BatchNorm2d(planes)
    self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
    self.bn3 = nn.BatchNorm2d(planes * 4)
    self.relu = nn.ReLU()
    self.downsample = downsample

  def forward(self, x):
    idt = x if self.downsample is None else self.downsample(x)
    out = self.relu(self.bn1(self.conv1(x)))
    out = self.relu(self.bn2(self.conv2(out)))
    out = self.bn3(self.conv3(out))
    return self.relu(out + idt)


class ResNet50Backbone(nn.Module):
  def __init__(self, width: int = 64, layers=(3, 4, 6, 3)):
    super().__init__()
    self.inplanes = width
    self.stem = nn.Sequential(nn.Conv2d(3, width, 7, 2, 3, bias=False), nn.BatchNorm2d(width), nn.ReLU(),
                              nn.MaxPool2d(3, 2, 1))
    self.layer1 = self._make(width, layers[0], 1)
    self.layer2 = self._make(width * 2, layers[1], 2)
    self.layer3 = self._make(width * 4, layers[2], 2)
    self.layer4 = self._make(width * 8, layers[3], 2)
    self.pool = nn.AdaptiveAvgPool2d(1)
    self.out_features = width * 8 * 4

  def _make(self, planes, blocks, stride):
    down = None
    if stride != 1 or self.inplanes != planes * 4:
      down = nn.Sequential(nn.Conv2d(self.inplanes, planes * 4, 1, stride, bias=False), nn.BatchNorm2d(planes * 4))
    mods = [Bottleneck(self.inplanes, planes, stride, down)]
    self.inplanes = planes * 4
    mods += [Bottleneck(self.inplanes, planes) for _ in range(1, blocks)]
    return nn.Sequential(*mods)

  def forward(self, x):
    x = self.stem(x.contiguous(memory_format=torch.channels_last))
    x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
    return torch.flatten(self.pool(x), 1)


class ResNet50(nn.Module):
  """``split_head=False``: plain DP model.  ``split_head=True``: the reference's replicate(N) backbone + split(N) head."""

  def __init__(self, num_classes: int = 10000, split_head: bool = False, width: int = 64, layers=(3, 4, 6, 3)):
    super().__init__()
    import easyparallellibrary_b200 as epl
    self.split_head = split_head
    if split_head:
      from easyparallellibrary_b200.ops.tensor_parallel import DistributedDense
      world = epl.Env.get().cluster.total_gpu_num if epl.Env.get().cluster else 1
      with epl.replicate(device_count=world):
        self.backbone = ResNet50Backbone(width, layers)
      self._split = epl.split(device_count=world)
      with self._split:
        self.head = DistributedDense(self.backbone.out_features, num_classes)
    else:
      self.backbone = ResNet50Backbone(width, layers)
      from easyparallellibrary_b200.ops.linear import Linear
      self.head = Linear(self.backbone.out_features, num_classes)

  def forward(self, images, labels=None):
    feats = self.backbone(images)
    if self.split_head:
      from easyparallellibrary_b200.ops import tensor_parallel as tp
      with self._split:
        logits = self.head(feats)
        if labels is None:
          return logits
        return tp.distributed_sparse_softmax_cross_entropy_with_logits(labels, logits)
    logits = self.head(feats)
    if labels is None:
      return logits
    from easyparallellibrary_b200.ops.cross_entropy import softmax_cross_entropy
    return softmax_cross_entropy(logits, labels)
