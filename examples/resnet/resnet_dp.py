"""ResNet-50 data parallel (reference ``examples/resnet/resnet_dp.py:26-92``): slim ResNet-50 + a wide dense head on synthetic
224x224 images, with the reference's toggles for gradient checkpointing, mixed precision and ZeRO.

  python examples/resnet/resnet_dp.py --steps 20                                   # 1 GPU
  torchrun --nproc-per-node 8 examples/resnet/resnet_dp.py --batch 128 --amp bf16  # DP8: fused NVLink reduce-scatter+AdamW+all-gather
  torchrun --nproc-per-node 8 examples/resnet/resnet_dp.py --gc auto --zero v1
  torchrun --nproc-per-node 8 examples/resnet/resnet_dp.py --split_head            # replicate(N) backbone + split(N) 10k-class head
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
ap.add_argument("--batch", type=int, default=32, help="images per GPU per step")
ap.add_argument("--classes", type=int, default=10000)
ap.add_argument("--image", type=int, default=224)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--width", type=int, default=64, help="64 = ResNet-50; smaller values give a toy model for CPU smoke runs")
ap.add_argument("--amp", default="bf16")
ap.add_argument("--gc", default="")
ap.add_argument("--zero", default="")
ap.add_argument("--split_head", action="store_true")
ap.add_argument("--optimizer", default="adamw")
args = ap.parse_args()

cuda = torch.cuda.is_available()
epl.init(epl.Config({"amp.level": args.amp if cuda else "", "gradient_checkpoint.type": args.gc, "zero.level": args.zero,
                     "cluster.colocate_split_and_replicate": args.split_head}))
rank = int(os.environ.get("RANK", 0))
layers = (3, 4, 6, 3) if args.width >= 64 else (1, 1, 1, 1)
if args.split_head:
  model = ResNet50(args.classes, split_head=True, width=args.width, layers=layers)
else:
  with epl.replicate(device_count=1):
    model = ResNet50(args.classes, width=args.width, layers=layers)
trainer = epl.Trainer(model, args.optimizer, lr=1e-3)
g = torch.Generator().manual_seed(rank)
t0 = time.time()
for step in range(args.steps):
  images = torch.randn(args.batch, 3, args.image, args.image, generator=g)
  labels = torch.randint(0, args.classes, (args.batch,), generator=g)
  out = trainer.step(images, labels)
  if rank == 0:
    print("step %d loss %.4f" % (step, out.item()), flush=True)
if rank == 0:
  dt = (time.time() - t0) / args.steps
  print("%.1f images/s per process (host clock, incl. data generation)" % (args.batch / dt), flush=True)
epl.shutdown()
