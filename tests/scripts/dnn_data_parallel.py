"""2-layer MLP data-parallel job for the launcher tests (reference: tests/dnn_data_parallel.py run through
``epl.utils.launcher`` by tests/Makefile:12-13).  ``--amp`` adds dynamic loss scaling with gradient accumulation and a
learning rate large enough to overflow fp16, and asserts that an overflowing step is skipped on every rank."""
import argparse
import sys

import torch

import easyparallellibrary_b200 as epl


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--steps", type=int, default=10)
  ap.add_argument("--amp", action="store_true")
  args = ap.parse_args()
  conf = {}
  if args.amp:
    conf = {"amp.level": "O1", "amp.loss_scale": "dynamic", "pipeline.num_micro_batch": 3}
  env = epl.init(epl.Config(conf))
  rank, world = env.cluster.worker_index, env.cluster.worker_num
  torch.manual_seed(0)
  with epl.replicate(device_count=1):
    model = torch.nn.Sequential(torch.nn.Linear(10, 16), torch.nn.ReLU(), torch.nn.Linear(16, 2))
  tr = epl.Trainer(model, "sgd", lr=(1000.0 if args.amp else 0.05), loss_fn=lambda y, t: torch.nn.functional.cross_entropy(y.float(), t))
  gen = torch.Generator().manual_seed(100 + rank)
  skipped, losses = 0, []
  for _ in range(args.steps):
    x = torch.randn(12, 10, generator=gen) * (50.0 if args.amp else 1.0)
    y = torch.randint(0, 2, (12,), generator=gen)
    out = tr.step(x, y)
    skipped += int(bool(out.skipped))
    losses.append(float(out.loss))
  flat = torch.cat([p.detach().float().flatten() for p in model.parameters()])
  import torch.distributed as dist
  if dist.is_initialized() and world > 1:
    ref = flat.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(ref, flat), "replicas diverged"
    t = torch.tensor([skipped])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    lo = int(t.item())
    dist.all_reduce(t.fill_(skipped), op=dist.ReduceOp.MAX)
    assert lo == int(t.item()), "ranks disagree on skipped steps"
  if args.amp:
    assert skipped >= 1, "no step was skipped although the loss scale must overflow"
  else:
    assert skipped == 0 and losses[-1] < losses[0] * 1.5
  print("rank %d/%d ok: skipped=%d loss %.4f -> %.4f" % (rank, world, skipped, losses[0], losses[-1]), flush=True)
  epl.shutdown()
  return 0


if __name__ == "__main__":
  sys.exit(main())
