"""ResNet-50 + 10 000-class head: data parallel, or replicate(N) backbone + split(N) head
(the reference's examples/resnet/resnet_dp.py and resnet_split.py; batch 32 per GPU, synthetic 224x224 images).

  torchrun --nproc-per-node 8 examples/train_resnet_split.py            # DP
  torchrun --nproc-per-node 8 examples/train_resnet_split.py --split    # DP backbone + class-parallel head
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import easyparallellibrary_b200 as epl
from easyparallellibrary_b200.models.resnet import ResNet50

ap = argparse.ArgumentParser()
ap.add_argument("--split", action="store_true")
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--steps", type=int, default=20)
args = ap.parse_args()
epl.init(epl.Config({"amp.level": "bf16", "cluster.colocate_split_and_replicate": args.split}))
if args.split:
  model = ResNet50(num_classes=10000, split_head=True)
else:
  with epl.replicate(device_count=1):
    model = ResNet50(num_classes=10000)
trainer = epl.Trainer(model, "adamw", lr=1e-3)
for step in range(args.steps):
  t0 = time.time()
  out = trainer.step(torch.randn(args.batch, 3, 224, 224), torch.randint(0, 10000, (args.batch,)))
  if int(os.environ.get("RANK", 0)) == 0:
    print("step %d loss %.4f (%.3f s)" % (step, out.item(), time.time() - t0), flush=True)
epl.shutdown()
