"""ResNet-50 + 10 000-class head with the head under ``split`` (reference ``examples/resnet/resnet_split.py:49-57``):
``replicate(N)`` backbone + ``split(N)`` dense head and class-parallel softmax cross-entropy on the same N GPUs
(``cluster.colocate_split_and_replicate``); batch 32 per GPU, synthetic 224x224 images.

  torchrun --nproc-per-node 8 examples/resnet/resnet_split.py           # DP backbone + class-parallel head
  torchrun --nproc-per-node 8 examples/resnet/resnet_split.py --dp      # plain DP for comparison (== resnet_dp.py)
  python examples/resnet/resnet_split.py --width 8 --classes 32 --image 32 --batch 4 --steps 2   # CPU smoke run
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import easyparallellibrary_b200 as epl
from easyparallellibrary_b200.models.resnet import ResNet50

ap = argparse.ArgumentParser()
ap.add_argument("--dp", action="store_true", help="replicate the head too (plain data parallelism)")
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--classes", type=int, default=10000)
ap.add_argument("--image", type=int, default=224)
ap.add_argument("--width", type=int, default=64, help="64 = ResNet-50; smaller values give a toy model for CPU smoke runs")
args = ap.parse_args()
args.split = not args.dp
epl.init(epl.Config({"amp.level": "bf16" if torch.cuda.is_available() else "", "cluster.colocate_split_and_replicate": args.split}))
layers = (3, 4, 6, 3) if args.width >= 64 else (1, 1, 1, 1)
if args.split:
  model = ResNet50(num_classes=args.classes, split_head=True, width=args.width, layers=layers)
else:
  with epl.replicate(device_count=1):
    model = ResNet50(num_classes=args.classes, width=args.width, layers=layers)
trainer = epl.Trainer(model, "adamw", lr=1e-3)
for step in range(args.steps):
  t0 = time.time()
  out = trainer.step(torch.randn(args.batch, 3, args.image, args.image), torch.randint(0, args.classes, (args.batch,)))
  if int(os.environ.get("RANK", 0)) == 0:
    print("step %d loss %.4f (%.3f s)" % (step, out.item(), time.time() - t0), flush=True)
epl.shutdown()
