"""MoE transformer training with expert parallelism (reference ``examples/moe/trainer.py`` + ``moe_ffn.py``): the experts of
every MoE layer are sharded over ``epl.split(N)``, tokens travel to their experts and back through the all-to-all kernels
(K5 / K5b: dispatch fused into the transfer), expert FFNs run on the tcgen05 GEMM, the auxiliary load-balancing loss is added
to the language-model loss, and the run reports tokens/s and the expert-capacity overflow it observed.

  python examples/moe/train_t5_moe.py --size tiny --steps 5
  torchrun --nproc-per-node 8 examples/moe/train_t5_moe.py --size small --experts 8 --batch 8 --seq 512
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import easyparallellibrary_b200 as epl
from easyparallellibrary_b200.models.moe_transformer import MoETransformer
from examples.moe.model_config.t5 import t5_moe

ap = argparse.ArgumentParser()
ap.add_argument("--size", default="small")
ap.add_argument("--experts", type=int, default=8)
ap.add_argument("--gating", default="top2", help="top2 | switch")
ap.add_argument("--capacity_factor", type=float, default=1.25)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--seq", type=int, default=512)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--lr", type=float, default=1e-4)
args = ap.parse_args()

epl.init(epl.Config({"amp.level": "bf16" if torch.cuda.is_available() else "", "cluster.colocate_split_and_replicate": True}))
world = epl.Env.get().cluster.total_gpu_num
rank = int(os.environ.get("RANK", 0))
epl.set_default_strategy(epl.replicate(device_count=world))       # reference trainer.py:165-170
cfg = t5_moe(args.size, num_experts=max(args.experts, world), gating=args.gating, capacity_factor=args.capacity_factor)
model = MoETransformer(cfg, expert_parallel=world)
trainer = epl.Trainer(model, "adamw", lr=args.lr)
g = torch.Generator().manual_seed(rank)
seq = min(args.seq, cfg.n_positions)
t0 = None
for step in range(args.steps):
  tok = torch.randint(0, cfg.vocab_size, (args.batch, seq), generator=g)
  out = trainer.step(tok, tok)
  if step == 1:
    t0 = time.time()
  if rank == 0:
    aux = sum(float(b.ffn.aux_loss) for b in model.blocks if b.use_moe and b.ffn.aux_loss is not None)
    print("step %d loss %.4f aux %.4f" % (step, out.item(), aux), flush=True)
if rank == 0 and t0 is not None and args.steps > 2:
  dt = (time.time() - t0) / (args.steps - 2)
  print("%.0f tokens/s (whole job, host clock)" % (args.batch * seq * world / dt), flush=True)
epl.shutdown()
