#!/usr/bin/env python
"""Headline benchmark: GPT-2-XL training throughput (tokens/s, whole job) on N B200s of one node.

  python bench.py --gpus 1 --steps 20 --warmup 5
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 \
      bench.py --gpus 8 --steps 20 --warmup 5

Metric/config follow BASELINE.json: GPT-2-XL (48 layers, d=1600, 25 heads), sequence 1024, bf16
compute with fp32 master weights, synthetic tokens, random-init weights, full training step
(forward + backward + gradient reduction + AdamW) through the public ``epl.Trainer`` API.

Timed region: exactly K steps bracketed by barrier + ``torch.cuda.synchronize()``, CUDA events on
the launching stream, max over ranks.  ``value`` uses device-resident inputs; ``e2e`` repeats the
measurement with every step's tokens copied from pinned host memory and the loss read back.
``--impl reference`` reports that the TF-1.15 reference cannot be installed in this image.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def parse():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=10)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="epl", choices=["epl", "reference", "baseline"])
  ap.add_argument("--model", default="xl")
  ap.add_argument("--batch", type=int, default=8, help="sequences per GPU per step")
  ap.add_argument("--seq", type=int, default=1024)
  ap.add_argument("--parallelism", default="auto", help="auto | dp | pp2 (2-stage pipeline x DP) | tp")
  ap.add_argument("--micro-batches", type=int, default=1)
  ap.add_argument("--zero", default="")
  ap.add_argument("--gc", default="")
  ap.add_argument("--no-e2e", action="store_true")
  ap.add_argument("--profile", default="", help="write a per-kernel GPU time table of 2 extra steps to this file")
  return ap.parse_args()


class ClockSampler(threading.Thread):
  """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

  def __init__(self, index: int):
    super().__init__(daemon=True)
    self.index, self.rows, self._halt = index, [], threading.Event()

  def run(self):
    q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    while not self._halt.is_set():
      try:
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if out:
          self.rows.append([c.strip() for c in out.split(",")])
      except Exception:
        pass
      self._halt.wait(0.2)

  def stop(self):
    self._halt.set()
    self.join(timeout=3)
    sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
    reasons = []
    for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
      if any(len(r) > col and r[col].lower().startswith("active") for r in self.rows):
        reasons.append(name)
    smax = max((float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()), default=0.0)
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax or None, "reasons": reasons, "samples": len(self.rows)}


def main():
  args = parse()
  if os.environ.get("EPL_HANG_DUMP"):          # debugging aid: dump every Python stack if the run stalls
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ["EPL_HANG_DUMP"]), exit=True)
  if args.impl == "reference":
    print(json.dumps({"impl": "reference", "unavailable":
                      "reference is TensorFlow-1.15/Python<=3.8 only (setup.py imports tensorflow; csrc links TF libs); "
                      "offline pip install into baseline/_ref fails with ModuleNotFoundError: tensorflow (see DESIGN.md)"}))
    return 0

  if args.parallelism.startswith("pp"):
    # pipeline: a first-time kernel launch must never block behind a posted NCCL receive (see parallel/pipeline.py), so load
    # every CUDA module up front instead of lazily; must be set before the CUDA context exists
    os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
  import torch
  import torch.distributed as dist
  import easyparallellibrary_b200 as epl
  from easyparallellibrary_b200.models.gpt2 import GPT2, GPT2Config
  from easyparallellibrary_b200.ops import _lib

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  if world != args.gpus and world == 1 and args.gpus > 1:
    # convenience: re-launch ourselves under torchrun
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000)] + sys.argv
    return subprocess.call(cmd)
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)

  par = args.parallelism
  if par == "auto":
    par = "dp"
  conf = {"amp.level": "bf16", "zero.level": args.zero, "gradient_checkpoint.type": args.gc,
          "pipeline.num_micro_batch": args.micro_batches}
  stages = 1
  if par.startswith("pp"):
    stages = int(par[2:] or 2)
    if conf["pipeline.num_micro_batch"] == 1:
      conf["pipeline.num_micro_batch"] = 8
  if args.impl == "baseline":
    conf["communication.fused_kernels"] = False
  epl.init(epl.Config(conf))
  cfg = GPT2Config.named(args.model, num_pipeline_stages=stages, tie_embeddings=(stages == 1), n_positions=max(1024, args.seq))
  torch.manual_seed(1234)
  with torch.device(dev):                      # random-init weights directly on the GPU
    if stages == 1:
      with epl.replicate(device_count=1):
        model = GPT2(cfg)
    else:
      model = GPT2(cfg)
  from easyparallellibrary_b200.models.gpt2 import lm_loss
  trainer = epl.Trainer(model, "adamw", lr=1e-4, weight_decay=0.01, baseline=(args.impl == "baseline"),
                        loss_fn=lm_loss if stages > 1 else None)
  trainer.build()
  M = conf["pipeline.num_micro_batch"]
  B = args.batch * (M if stages > 1 else 1)
  gen = torch.Generator(device="cpu").manual_seed(rank)
  n_host = 4
  host = [torch.randint(0, 50257, (B, args.seq), generator=gen).pin_memory() for _ in range(n_host)]
  dev_tokens = [h.to(dev) for h in host]

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def run(n, e2e):
    sink = 0.0
    for i in range(n):
      if e2e:
        tok = host[i % n_host].to(dev, non_blocking=True)
        out = trainer.step(tok, tok)
        sink += out.loss.item()                 # device -> host read of the step's result
      else:
        tok = dev_tokens[i % n_host]
        out = trainer.step(tok, tok)
    return out, sink

  phase = None
  try:                                         # exposed reduce + optimizer phase per step (events on the main stream)
    from easyparallellibrary_b200.utils.metric import PhaseTimer
    phase = PhaseTimer(trainer, "_reduce_and_apply", use_cuda=True)
  except Exception:
    phase = None
  apply_ms = [None]

  def timed(n, e2e):
    barrier()
    if phase is not None:
      phase.reset()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
      sampler.start()
    l0 = _lib.launches
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    out, _ = run(n, e2e)
    t1.record()
    barrier()
    ms = t0.elapsed_time(t1)
    clocks = sampler.stop() if sampler else None
    try:
      pm = phase.total_ms() / n if (phase is not None and not e2e) else -1.0
    except Exception:
      pm = -1.0
    t = torch.tensor([ms, pm], device=dev, dtype=torch.float64)
    if world > 1:
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if not e2e:
      apply_ms[0] = float(t[1].item()) if float(t[1].item()) >= 0 else None
    return float(t[0].item()), out, clocks, _lib.launches - l0

  run(max(args.warmup, 3), False)
  if args.profile:                               # kernel timeline of 2 steps (CUPTI via torch.profiler); never a bench value
    from torch.profiler import profile, ProfilerActivity
    barrier()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
      run(2, False)
      torch.cuda.synchronize()
    if rank == 0:
      from easyparallellibrary_b200.profiler.timeline import kernel_table
      text, _ = kernel_table(prof.events())
      evs = [e for e in prof.events() if e.device_type.name == "CUDA"]
      with open(args.profile, "w") as f:
        f.write(text)
        # GEMM durations in launch order (first profiled step): 48 x [qkv, proj, fc1, fc2] forward, lm_head, then backward
        gem = sorted((e.time_range.start, e.time_range.elapsed_us()) for e in evs if "gemm" in e.name)
        gem = gem[:len(gem) // 2]
        f.write("\nGEMM kernel durations (us) in launch order, step 1:\n")
        for i in range(0, len(gem), 12):
          f.write(" ".join("%6.0f" % d for _, d in gem[i:i + 12]) + "\n")
    barrier()
  ms, out, clocks, launches = timed(args.steps, False)
  dp_replicas = trainer.plan.num_replicas
  tokens_per_step = args.batch * args.seq * (M if stages > 1 else 1) * dp_replicas
  value = tokens_per_step * args.steps / (ms / 1e3)
  e2e = None
  if not args.no_e2e:
    run(2, True)
    ms_e, _, _, _ = timed(args.steps, True)
    e2e = {"value": tokens_per_step * args.steps / (ms_e / 1e3), "unit": "tokens/s",
           "h2d_bytes_per_step": int(host[0].numel() * host[0].element_size()), "d2h_bytes_per_step": 4,
           "ms_per_step": ms_e / args.steps}
  if rank == 0:
    peaks = {}
    try:
      peaks = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")))
    except Exception:
      pass
    flops = cfg.flops_per_token(args.seq) * value / world
    line = {
        "metric": "tokens/sec (whole job, device-timed, max over ranks) GPT-2-XL training step",
        "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic tokens, random-init weights", "impl": args.impl,
        "config": {"model": "gpt2-" + args.model, "params": cfg.num_params, "global_batch": args.batch * (M if stages > 1 else 1) * dp_replicas,
                   "seq_len": args.seq, "parallelism": ("dp%d" % dp_replicas) + ("xpp%d" % stages if stages > 1 else ""),
                   "micro_batches": M, "zero": args.zero or ("fused-rs-adam-ag" if trainer.fused is not None else "none"),
                   "optimizer": "adamw fp32 master", "l2": "working set (3 GB bf16 weights + activations) >> 126 MB L2; no flush needed"},
        "model_tflops_per_gpu": flops / 1e12,
        "mfu_of_measured_bf16_sustained": (flops / 1e12) / peaks["bf16_tflops_sustained"] if peaks.get("bf16_tflops_sustained") else None,
        "loss": float(out.loss), "clocks": clocks, "gpu_launches": launches, "e2e": e2e,
    }
    try:      # gradient reduction + optimizer phase: measured (exposed: it runs after backward) vs its roofline
      if apply_ms[0] is not None and stages == 1:
        from easyparallellibrary_b200.utils.metric import fused_dp_roofline_ms
        roof = fused_dp_roofline_ms(cfg.num_params, world, hbm_gbs=float(peaks.get("hbm_gbs", 6400.0)))
        line["reduce_apply"] = {"ms_per_step": apply_ms[0], "roofline_ms": roof, "pct_of_roofline": 100.0 * roof / apply_ms[0],
                                "what": ("fused reduce-scatter + AdamW + all-gather kernels (NVLink peer memory)" if trainer.fused is not None
                                         else ("bucketed NCCL all-reduce + Adam" if world > 1 else "fused AdamW")),
                                "roofline": "2 x remote gradient/weight bytes over 900 GB/s NVLink + 24 B/param optimizer state of the 1/W shard "
                                            "over measured HBM bandwidth (1 GPU: 30 B/param)"}
    except Exception:
      pass
    print(json.dumps(line))
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()
  return 0


if __name__ == "__main__":
  sys.exit(main())
