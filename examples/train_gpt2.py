"""GPT-2 training: data parallel, pipeline x data parallel, ZeRO / recompute / offload — all from the config.

  torchrun --nproc-per-node 8 examples/train_gpt2.py --model xl                         # DP8, fused NVLink optimizer path
  torchrun --nproc-per-node 8 examples/train_gpt2.py --model xl --stages 2 --micro 8   # 2-stage 1F1B pipeline x DP4
  torchrun --nproc-per-node 8 examples/train_gpt2.py --model xl --zero v3 --gc auto --offload v0
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import easyparallellibrary_b200 as epl
from easyparallellibrary_b200.models.gpt2 import GPT2, GPT2Config, lm_loss

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="small")
ap.add_argument("--stages", type=int, default=1)
ap.add_argument("--micro", type=int, default=1)
ap.add_argument("--zero", default="")
ap.add_argument("--gc", default="")
ap.add_argument("--offload", default="")
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--seq", type=int, default=1024)
ap.add_argument("--steps", type=int, default=20)
args = ap.parse_args()

epl.init(epl.Config({"amp.level": "bf16", "pipeline.num_micro_batch": args.micro, "zero.level": args.zero,
                     "gradient_checkpoint.type": args.gc, "offload.level": args.offload}))
cfg = GPT2Config.named(args.model, num_pipeline_stages=args.stages, tie_embeddings=args.stages == 1)
if args.stages == 1:
  with epl.replicate(device_count=1):
    model = GPT2(cfg)
else:
  model = GPT2(cfg)                         # the model cuts its own stages with epl.set_default_strategy
trainer = epl.Trainer(model, "adamw", lr=1e-4, loss_fn=lm_loss if args.stages > 1 else None)
rank = int(os.environ.get("RANK", 0))
g = torch.Generator().manual_seed(rank)
for step in range(args.steps):
  tokens = torch.randint(0, cfg.vocab_size, (args.batch * args.micro, args.seq), generator=g)
  out = trainer.step(tokens, tokens)
  if rank == 0:
    print("step %d loss %.4f" % (step, out.item()), flush=True)
epl.shutdown()
