"""MoE transformer with expert parallelism (the reference's examples/moe: 8 experts, top-2 gating, capacity 1.25).

  torchrun --nproc-per-node 8 examples/train_moe.py --experts 8
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import easyparallellibrary_b200 as epl
from easyparallellibrary_b200.models.moe_transformer import MoEConfig, MoETransformer

ap = argparse.ArgumentParser()
ap.add_argument("--experts", type=int, default=8)
ap.add_argument("--gating", default="top2")
ap.add_argument("--batch", type=int, default=2)
ap.add_argument("--seq", type=int, default=512)
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
epl.init(epl.Config({"amp.level": "bf16", "cluster.colocate_split_and_replicate": True}))
world = epl.Env.get().cluster.total_gpu_num
epl.set_default_strategy(epl.replicate(device_count=world))       # reference trainer.py:165-170
model = MoETransformer(MoEConfig(num_experts=args.experts, gating=args.gating), expert_parallel=world)
trainer = epl.Trainer(model, "adamw", lr=1e-4)
g = torch.Generator().manual_seed(int(os.environ.get("RANK", 0)))
for step in range(args.steps):
  tok = torch.randint(0, 32000, (args.batch, args.seq), generator=g)
  out = trainer.step(tok, tok)
  if int(os.environ.get("RANK", 0)) == 0:
    print("step %d loss %.4f" % (step, out.item()), flush=True)
epl.shutdown()
