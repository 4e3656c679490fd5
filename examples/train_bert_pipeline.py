"""BERT-large 2-stage pipeline (the reference's examples/bert/scripts/train_bert_large_pipe.sh: batch 20, 10 micro-batches,
sequence 384) and the tensor-parallel variant (BASELINE config: BERT-large epl.split(8)).

  torchrun --nproc-per-node 2 examples/train_bert_pipeline.py --stages 2 --micro 10
  torchrun --nproc-per-node 8 examples/train_bert_pipeline.py --tp 8
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import easyparallellibrary_b200 as epl
from easyparallellibrary_b200.models.bert import Bert, BertConfig, squad_loss

ap = argparse.ArgumentParser()
ap.add_argument("--size", default="large")
ap.add_argument("--stages", type=int, default=1)
ap.add_argument("--micro", type=int, default=1)
ap.add_argument("--tp", type=int, default=1)
ap.add_argument("--batch", type=int, default=20)
ap.add_argument("--seq", type=int, default=384)
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()

epl.init(epl.Config({"amp.level": "bf16", "pipeline.num_micro_batch": args.micro,
                     "cluster.colocate_split_and_replicate": args.tp > 1}))
if args.tp > 1:
  epl.set_default_strategy(epl.replicate(device_count=1))
bcfg = BertConfig.named(args.size, num_pipeline_stages=args.stages, tensor_parallel=args.tp)
model = Bert(bcfg)
loss_fn = squad_loss if args.stages > 1 else None
trainer = epl.Trainer(model, "adamw", lr=3e-5, loss_fn=loss_fn)
g = torch.Generator().manual_seed(0 if args.tp > 1 else int(os.environ.get("RANK", 0)))
for step in range(args.steps):
  ids = torch.randint(0, min(30000, bcfg.vocab_size), (args.batch, args.seq), generator=g)
  start, end = torch.randint(0, args.seq, (args.batch,), generator=g), torch.randint(0, args.seq, (args.batch,), generator=g)
  out = trainer.step(ids, start, end)
  if int(os.environ.get("RANK", 0)) == 0:
    print("step %d loss %s" % (step, out.loss), flush=True)
epl.shutdown()
