"""Multi-GPU checks, run under torchrun on 2+ GPUs:
   native NCCL communicator verbs, symmetric memory (barrier, peer copy bandwidth), and the fused
   reduce-scatter+AdamW+all-gather kernel vs the NCCL library path (parameter parity + timing)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import easyparallellibrary_b200 as epl


def log(*a):
  if dist.get_rank() == 0:
    print(*a, flush=True)


def stage(name):
  print("[rank %d] %s" % (dist.get_rank(), name), flush=True)


def check_native(dev, world, rank):
  from easyparallellibrary_b200.communicators.native import NativeBackend
  stage("native: create")
  be = NativeBackend(list(range(world)), dev)
  stage("native: verbs")
  t = torch.full((1024,), float(rank + 1), device=dev)
  be.all_reduce(t)
  assert t[0].item() == world * (world + 1) / 2
  g = be.all_gather(torch.full((2, 3), float(rank), device=dev))
  assert g.shape == (2 * world, 3) and g[2 * (world - 1), 0].item() == world - 1
  rs = be.reduce_scatter(torch.arange(4 * world, device=dev, dtype=torch.float32))
  assert torch.equal(rs, torch.arange(4 * rank, 4 * rank + 4, device=dev, dtype=torch.float32) * world)
  b = torch.full((5,), float(rank), device=dev)
  be.broadcast(b, world - 1)
  assert b[0].item() == world - 1
  r = be.reduce(torch.ones(3, device=dev), 0)
  if rank == 0:
    assert r[0].item() == world
  a2a = be.all_to_all(torch.arange(world, device=dev, dtype=torch.float32) + 10 * rank)
  assert a2a.tolist() == [rank + 10.0 * p for p in range(world)]
  gv, cnt = be.all_gatherv(torch.full((rank + 1, 2), float(rank), device=dev))
  assert gv.shape[0] == world * (world + 1) // 2 and cnt.tolist() == list(range(1, world + 1))
  rows = torch.arange(world * 2, device=dev, dtype=torch.float32).view(-1, 1) + 100 * rank
  out, rc = be.all_to_allv(rows, torch.full((world,), 2))
  assert out.shape[0] == 2 * world and out[0, 0].item() == 2 * rank
  stage("native: p2p")
  if rank == 0:
    be.send(torch.full((4,), 7.0, device=dev), 1); be.wait()
  elif rank == 1:
    x = torch.zeros(4, device=dev); be.recv(x, 0); be.wait(); torch.cuda.synchronize(); assert x[0].item() == 7.0
  torch.cuda.synchronize()
  # bandwidth of the native all-reduce on 256 MiB
  big = torch.ones(64 << 20, device=dev)
  for _ in range(3):
    be.all_reduce(big)
  torch.cuda.synchronize(); dist.barrier()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(5):
    be.all_reduce(big)
  e1.record(); torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / 5
  log("native communicator ok; all-reduce 256 MiB: %.3f ms, busbw %.1f GB/s" % (ms, 2 * (world - 1) / world * big.numel() * 4 / ms / 1e6))
  be.close()


def check_symm(dev, world, rank):
  from easyparallellibrary_b200.runtime.symmetric import SignalPad, SymmetricBuffer, _sym_lib
  from easyparallellibrary_b200.ops import _lib
  stage("symm: alloc")
  n = 256 << 20
  buf = SymmetricBuffer(n, list(range(world)), dev)
  mine = buf.tensor(torch.float32, n // 4)
  mine.fill_(float(rank))
  pad = SignalPad(4, list(range(world)), dev)
  pad.barrier(0)
  peer = (rank + 1) % world
  src_ptr = buf.peer_ptrs[peer]
  dst = torch.empty(n // 4, device=dev)
  lib = _sym_lib()
  for blocks in (32, 148, 296):
    for _ in range(2):
      lib.epl_peer_copy(src_ptr, dst.data_ptr(), n, blocks, _lib.stream())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
      lib.epl_peer_copy(src_ptr, dst.data_ptr(), n, blocks, _lib.stream())
    e1.record(); torch.cuda.synchronize()
    log("peer copy (kernel ld over NVLink), %3d CTAs: %.1f GB/s" % (blocks, n * 5 / e0.elapsed_time(e1) / 1e6))
  assert dst[777].item() == float(peer)
  pad.barrier(0)
  torch.cuda.synchronize()
  log("symmetric memory ok")


def train(dev, world, rank, fused, steps=6, model_name="tiny", batch=4, seq=128, lr=1e-3, eps=1e-8):
  from easyparallellibrary_b200.models.gpt2 import GPT2, GPT2Config
  epl.init(epl.Config({"amp.level": "bf16", "communication.fused_kernels": fused}))
  torch.manual_seed(0)
  with epl.replicate(1):
    cfg = GPT2Config.named(model_name) if model_name != "tiny" else GPT2Config.named("tiny", n_embd=256, n_head=4, vocab_size=2048)
    model = GPT2(cfg)
  tr = epl.Trainer(model, "adamw", lr=lr, eps=eps).build()
  g = torch.Generator().manual_seed(100 + rank)
  toks = [torch.randint(0, cfg.vocab_size, (batch, seq), generator=g).to(dev) for _ in range(steps)]
  losses = []
  for t in toks:
    losses.append(tr.step(t, t).item())
  torch.cuda.synchronize()
  flat = torch.cat([b.flat_param.float().flatten() for b in tr.flats[0].buckets])
  return tr, losses, flat.clone()


def check_fused_nvls(dev, world, rank):
  """K1 over NVLS (multimem.ld_reduce / multimem.st on multicast-mapped buckets) vs the NCCL path, then its step time."""
  os.environ["EPL_K1"] = "nvls"
  try:
    tr_b, loss_b, p_b = train(dev, world, rank, fused=False, lr=1e-2, eps=1.0)
    tr_f, loss_f, p_f = train(dev, world, rank, fused=True, lr=1e-2, eps=1.0)
    assert tr_f.fused is not None and tr_f.fused.kernel == "nvls"
    bufs = [b for b in tr_f._symm_buffers.values()]
    nv = all(getattr(b, "multicast_ptr", 0) for b in bufs)
    diff = (p_b - p_f).abs().max().item()
    gathered = [torch.zeros_like(p_f) for _ in range(world)]
    dist.all_gather(gathered, p_f)
    same = max((gathered[0] - g).abs().max().item() for g in gathered)
    log("K1 over NVLS (multicast buckets: %s) vs NCCL path (eps=1): max |dparam| = %.3e, replica divergence %.1e, losses %s vs %s" % (
        nv, diff, same, loss_f[-2:], loss_b[-2:]))
    assert same == 0.0 and diff < 8e-3 and abs(loss_f[-1] - loss_b[-1]) < 0.03
  finally:
    os.environ["EPL_K1"] = "v2"


def check_fused(dev, world, rank):
  tr_b, loss_b, p_b = train(dev, world, rank, fused=False)
  assert tr_b.fused is None
  tr_f, loss_f, p_f = train(dev, world, rank, fused=True)
  assert tr_f.fused is not None, "fused data-parallel path did not engage"
  diff = (p_b - p_f).abs().max().item()
  gathered = [torch.zeros_like(p_f) for _ in range(world)]
  dist.all_gather(gathered, p_f)
  same = max((gathered[0] - g).abs().max().item() for g in gathered)
  log("fused vs NCCL path: max |dparam| = %.3e, replica divergence = %.3e, losses %s vs %s" % (diff, same, loss_f[-2:], loss_b[-2:]))
  # default AdamW (eps 1e-8) is sign-like in the first steps: a bf16-rounding difference in a near-zero gradient moves a
  # weight by 2 x lr, so only coarse agreement can be asserted here ...
  assert same == 0.0 and diff < 2.5e-2 and abs(loss_f[-1] - loss_b[-1]) < 0.05 * abs(loss_b[-1]) and loss_f[-1] < loss_f[0]
  # ... the strict comparison uses a large eps (update ~ lr * m / eps, linear in the gradient, no sign amplification)
  tr_b, loss_b, p_b = train(dev, world, rank, fused=False, lr=1e-2, eps=1.0)
  tr_f, loss_f, p_f = train(dev, world, rank, fused=True, lr=1e-2, eps=1.0)
  diff = (p_b - p_f).abs().max().item()
  log("fused vs NCCL path (eps=1): max |dparam| = %.3e, losses %s vs %s" % (diff, loss_f[-2:], loss_b[-2:]))
  # parameters are compared in bf16: one ulp is 3.9e-3 for the LayerNorm gains near 1
  assert diff < 8e-3 and abs(loss_f[-1] - loss_b[-1]) < 0.03          # 8 ranks: 0.016 observed (bf16 ring all-reduce vs fp32 sum of bf16 partials)


def _k1_setup(dev, world, rank, n_bucket, dtype, seed):
  """One gradient bucket of ``n_bucket`` elements in symmetric memory + this rank's fp32 optimizer shard."""
  from easyparallellibrary_b200.runtime.symmetric import SignalPad, SymmetricBuffer
  ranks = list(range(world))
  gbuf = SymmetricBuffer(n_bucket * 2, ranks, dev)
  pbuf = SymmetricBuffer(n_bucket * 2, ranks, dev)
  pad = SignalPad(1, ranks, dev)
  g = gbuf.tensor(dtype, n_bucket)
  w = pbuf.tensor(dtype, n_bucket)
  gen = torch.Generator(device=dev).manual_seed(seed + rank)
  g.copy_((torch.randn(n_bucket, device=dev, generator=gen) * 0.02).to(dtype))
  gen0 = torch.Generator(device=dev).manual_seed(seed + 1000)               # identical weights / state on every rank
  master_full = torch.randn(n_bucket, device=dev, generator=gen0) * 0.05
  w.copy_(master_full.to(dtype))
  shard = n_bucket // world
  lo = rank * shard
  st = {"master": master_full[lo:lo + shard].clone(), "m": torch.randn(shard, device=dev, generator=gen0) * 1e-3,
        "v": torch.rand(shard, device=dev, generator=gen0) * 1e-5,
        "mask": (torch.rand(shard, device=dev, generator=gen0) > 0.1).float()}
  return gbuf, pbuf, pad, g, w, st, lo, shard


def _k1_launch(lib, gbuf, pbuf, pad, sync, st, lo, shard, rank, world, dtype, dyn, hyper, blocks, use_mask=True):
  from easyparallellibrary_b200.ops import _lib
  b1, b2, eps, wd = hyper
  rc = lib.epl_fused_rs_adam_ag_v2(gbuf.peer_table(0), pbuf.peer_table(0), pad.slot_table(0), sync.data_ptr(),
                                   st["master"].data_ptr(), st["m"].data_ptr(), st["v"].data_ptr(),
                                   st["mask"].data_ptr() if use_mask else None, lo, shard, rank, world, 0, _lib.dtype_code(dtype),
                                   dyn.data_ptr(), b1, b2, eps, wd, blocks, _lib.stream())
  _lib.check(rc, "fused_rs_adam_ag_v2")


def check_k1(dev, world, rank):
  """K1 v2 against fp32 ground truth: gradients of all ranks are gathered, summed in fp32 in rank order (the kernel's
  order, so the reduction must match BIT FOR BIT), AdamW is evaluated in fp32 with the same formula; master / m / v must
  agree to fp32 rounding of a handful of operations and the bf16/fp16 weights every rank ends up with must be the
  rounding of the new master — and identical on every rank.  Also: a ragged bucket (last piece partial), several
  consecutive launches (epoch protocol, grid size changing between launches), fp16 + loss scale."""
  from easyparallellibrary_b200.runtime.symmetric import _sym_lib
  lib = _sym_lib()
  hyper = (0.9, 0.999, 1e-8, 0.01)
  for dtype, n_bucket, scale in ((torch.bfloat16, world * 8 * 40961, 1.0), (torch.float16, world * 8 * 1031, 1.0 / 128),
                                 (torch.bfloat16, world * 8 * 5, 0.5)):
    gbuf, pbuf, pad, g, w, st, lo, shard = _k1_setup(dev, world, rank, n_bucket, dtype, 7)
    sync = torch.zeros(4, dtype=torch.int32, device=dev)
    ref = {k: v.clone() for k, v in st.items()}
    for step, blocks in enumerate((148, 6, 33)):
      lr, t = 1e-3 * (step + 1), step + 1
      inv_c1, inv_c2 = 1.0 / (1 - hyper[0] ** t), 1.0 / (1 - hyper[1] ** t)
      dyn = torch.tensor([lr, inv_c1, inv_c2, scale / world], dtype=torch.float32, device=dev)
      # ---- fp32 truth
      allg = [torch.empty_like(g) for _ in range(world)]
      dist.all_gather(allg, g)
      acc = torch.zeros(shard, device=dev)
      for r in range(world):
        acc += allg[r][lo:lo + shard].float()                                # rank order, fp32: exact
      gr = acc * torch.tensor(scale / world, dtype=torch.float32, device=dev)
      b1, b2, eps, wd = hyper
      f32 = lambda x: torch.tensor(x, dtype=torch.float32)
      omb1, omb2 = float(f32(1.0) - f32(b1)), float(f32(1.0) - f32(b2))      # the kernel forms 1 - beta in fp32
      ref["m"] = float(f32(b1)) * ref["m"] + omb1 * gr
      ref["v"] = float(f32(b2)) * ref["v"] + omb2 * gr * gr
      upd = (ref["m"] * inv_c1) / ((ref["v"] * inv_c2).sqrt() + eps) + wd * ref["mask"] * ref["master"]
      ref["master"] = ref["master"] - lr * upd
      # ---- kernel
      dist.barrier()
      _k1_launch(lib, gbuf, pbuf, pad, sync, st, lo, shard, rank, world, dtype, dyn, hyper, blocks)
      torch.cuda.synchronize(); dist.barrier()
      dm = (st["m"] - ref["m"]).abs().max().item() / (ref["m"].abs().max().item() + 1e-30)
      dv = (st["v"] - ref["v"]).abs().max().item() / (ref["v"].abs().max().item() + 1e-30)
      dp = (st["master"] - ref["master"]).abs().max().item()
      # weights: every rank's full buffer == rounding of the owners' new masters
      mine = st["master"].to(dtype)
      gathered = [torch.empty_like(mine) for _ in range(world)]
      dist.all_gather(gathered, mine)
      expect = torch.cat(gathered)
      dw = (w.float() - expect.float()).abs().max().item()
      log("k1 exact %s n=%d step %d blocks=%d: rel|dm|=%.1e rel|dv|=%.1e |dmaster|=%.1e |dweights vs rounded master|=%.1e" % (
          str(dtype).split(".")[-1], n_bucket, step, blocks, dm, dv, dp, dw))
      # m, v: two fp32 FMAs vs mul+add (<= 2 ulp = 2.4e-7 relative); master: + one division / sqrt (fast-math free build)
      assert dm < 1e-6 and dv < 1e-6 and dp < 2e-6 * max(1.0, lr * 1e3) and dw == 0.0, (dm, dv, dp, dw)
      for k in ("master", "m", "v"):                                        # continue from the kernel's state: errors do not pile up
        ref[k] = st[k].clone()
      g.copy_((g.float() * 0.5 + 0.01).to(dtype))                           # new gradients for the next step


def bench_k1(dev, world, rank):
  """K1 v1 vs v2 on GPT-2-XL-sized buckets: whole GPU (exposed tail) and a few CTAs (overlapped with backward)."""
  from easyparallellibrary_b200.runtime.symmetric import _sym_lib
  from easyparallellibrary_b200.ops import _lib
  lib = _sym_lib()
  hyper = (0.9, 0.999, 1e-8, 0.01)
  res = {}
  for n_bucket in (1557611200 // 16 // (world * 8) * (world * 8), 1557611200 // 5 // (world * 8) * (world * 8)):
    gbuf, pbuf, pad, g, w, st, lo, shard = _k1_setup(dev, world, rank, n_bucket, torch.bfloat16, 3)
    sync = torch.zeros(4, dtype=torch.int32, device=dev)
    sync1 = torch.zeros(4, dtype=torch.int32, device=dev)
    from easyparallellibrary_b200.runtime.symmetric import SignalPad
    pad1 = SignalPad(1, list(range(world)), dev)
    dyn = torch.tensor([1e-4, 1.0, 1.0, 1.0 / world], dtype=torch.float32, device=dev)
    ep = [0]

    def v1():
      ep[0] += 1
      rc = lib.epl_fused_rs_adam_ag(gbuf.peer_table(0), pbuf.peer_table(0), pad1.slot_table(0), sync1.data_ptr(),
                                    st["master"].data_ptr(), st["m"].data_ptr(), st["v"].data_ptr(), st["mask"].data_ptr(), lo, shard,
                                    rank, world, ep[0], _lib.dtype_code(torch.bfloat16), 1e-4, 0.9, 0.999, 1e-8, 0.01, 1.0 / world,
                                    1.0, 1.0, 148, _lib.stream())
      _lib.check(rc, "k1v1")
    cases = [("v1 148 CTAs", v1)] + [("v2 %3d CTAs" % b, (lambda b=b: _k1_launch(lib, gbuf, pbuf, pad, sync, st, lo, shard, rank, world,
                                                                             torch.bfloat16, dyn, hyper, b))) for b in (148, 74, 32, 16, 8, 4)]
    for name, fn in cases:
      for _ in range(2):
        fn()
      torch.cuda.synchronize(); dist.barrier()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      for _ in range(5):
        fn()
      e1.record(); torch.cuda.synchronize()
      t = torch.tensor([e0.elapsed_time(e1) / 5], device=dev)
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
      link = (world - 1) / world * n_bucket * 2 * 2            # bytes per direction: peers' grads in + weights in (= grads out + weights out)
      hbm = n_bucket / world * 32 + n_bucket * 2 * 2
      floor = max(link / 770e9, hbm / 6.4e12) * 1e3
      res["%d:%s" % (n_bucket, name)] = t.item()
      log("k1 bucket %.0f M params, %s: %.3f ms  (per-direction link bytes %.0f MB -> %.0f GB/s; floor max(link@770GB/s, HBM) = %.3f ms -> %.0f %%)" % (
          n_bucket / 1e6, name, t.item(), link / 1e6, link / t.item() / 1e6, floor, 100 * floor / t.item()))
    del gbuf, pbuf, pad, pad1
  if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/k1_bench_w%d.json" % world, "w"), indent=1)


def check_tp(dev, world, rank):
  """Fused all-gather->GEMM and GEMM->reduce-scatter kernels vs the NCCL + separate GEMM path."""
  from easyparallellibrary_b200.ops import tensor_parallel as tp
  from easyparallellibrary_b200.ops import tp_fused
  epl.init(epl.Config({"cluster.colocate_split_and_replicate": True}))
  with epl.split(world):
    group = tp.current_tp_group()
  torch.manual_seed(7)
  results = {}
  for (T, K, N) in ((1024 * world, 1024, 4096 // world * world), (8192, 1600, 6400 // world // 8 * 8)):
    T = T // (128 * world) * (128 * world)
    xs = (torch.randn(T // world, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    b = torch.randn(N, device=dev).bfloat16()
    a = (torch.randn(T, K, device=dev) * 0.5).bfloat16()
    outs = {}
    for fused in (False, True):
      tp_fused.USE_FUSED = fused
      y, pre, xf = tp_fused.ag_gemm(xs, w, group, bias=b, gelu=True)
      z = tp_fused.gemm_rs(a, w, group)
      torch.cuda.synchronize()
      outs[fused] = (y.float(), z.float(), xf.float())
    # K2: the gathered operand is the (ZeRO-3 sharded) weight
    from easyparallellibrary_b200.ops import tp_kernels
    wsh = w[rank * (N // world):(rank + 1) * (N // world)].contiguous()
    y2, _, wfull = tp_kernels.ag_weight_gemm(a, wsh, group, bias=b, gelu=False)
    torch.cuda.synchronize()
    wref = group.comm.allgather(wsh)
    yref = torch.nn.functional.linear(a.float(), wref.float(), b.float())
    dw2 = (wfull.float() - wref.float()).abs().max().item()
    dy2 = (y2.float() - yref).abs().max().item()
    log("   K2 weight-gather GEMM: |w_full diff|=%.1e |y diff|=%.3e (ref max %.2f)" % (dw2, dy2, yref.abs().max().item()))
    assert dw2 == 0.0 and dy2 < 0.02 * yref.abs().max().item() + 0.1
    dy = (outs[True][0] - outs[False][0]).abs().max().item()
    dz = (outs[True][1] - outs[False][1]).abs().max().item()
    dx = (outs[True][2] - outs[False][2]).abs().max().item()
    ref = outs[False][1].abs().max().item()
    log("tp T=%d K=%d N=%d: |ag_gemm diff|=%.3e |x_full diff|=%.1e |gemm_rs diff|=%.3e (ref max %.2f)" % (T, K, N, dy, dx, dz, ref))
    assert dx == 0.0 and dy < 0.1 and dz < 0.02 * ref + 0.1
    for fused in (False, True):
      tp_fused.USE_FUSED = fused
      from easyparallellibrary_b200.ops import linear as LL
      k2 = (lambda: tp_kernels.ag_weight_gemm(a, wsh, group, bias=b)) if fused else (lambda: LL.gemm(a, group.comm.allgather(wsh), bias=b, epilogue=LL.EPI_BIAS))
      for name, fn in (("ag_gemm", lambda: tp_fused.ag_gemm(xs, w, group, bias=b, gelu=True)), ("gemm_rs", lambda: tp_fused.gemm_rs(a, w, group)),
                       ("agw_gemm", k2)):
        for _ in range(3):
          fn()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
          fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        results[(T, K, N, name, fused)] = t.item()
    for name in ("ag_gemm", "gemm_rs", "agw_gemm"):
      fl = 2.0 * T * K * N
      nv = (world - 1) / world * (T * K if name == "ag_gemm" else (T * N if name == "gemm_rs" else N * K)) * 2
      log("  %-8s NCCL+GEMM %.3f ms | fused %.3f ms | speedup %.2fx | roofline max(gemm %.3f ms @1.4PF, nvlink %.3f ms @770GB/s)" % (
          name, results[(T, K, N, name, False)], results[(T, K, N, name, True)],
          results[(T, K, N, name, False)] / results[(T, K, N, name, True)], fl / 1.4e12, nv / 770e6))
  tp_fused.USE_FUSED = True
  if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump({"%d_%d_%d_%s_%s" % k: v for k, v in results.items()}, open("gpurun_out/tp_fused_bench_w%d.json" % world, "w"))


def check_tp_train(dev, world, rank):
  """Tensor-parallel BERT layers trained for a few steps with the fused collective GEMMs (K3a / K3b forward AND their
  backward pairing in ops/tp_fused.py) vs the NCCL + GEMM path: losses and weights must agree."""
  from easyparallellibrary_b200.models.bert import Bert, BertConfig
  from easyparallellibrary_b200.ops import tp_fused
  res = {}
  for fused in (False, True):
    tp_fused.USE_FUSED = fused
    epl.init(epl.Config({"amp.level": "bf16", "cluster.colocate_split_and_replicate": True}))
    epl.set_default_strategy(epl.replicate(device_count=1))
    torch.manual_seed(0)
    heads = max(4, world)                                                # at least one 64-wide head per tensor-parallel rank
    model = Bert(BertConfig.named("tiny", hidden_size=64 * heads, num_attention_heads=heads, intermediate_size=256 * heads,
                                  tensor_parallel=world))
    tr = epl.Trainer(model, "adamw", lr=1e-3, eps=1.0).build()          # eps=1: no sign amplification of rounding differences
    g = torch.Generator().manual_seed(3)                                 # the whole TP group sees the same batch
    losses = []
    for _ in range(4):
      ids = torch.randint(0, 1000, (8, 128), generator=g).to(dev)
      s, e = torch.randint(0, 128, (8,), generator=g).to(dev), torch.randint(0, 128, (8,), generator=g).to(dev)
      losses.append(float(tr.step(ids, s, e).loss))
    torch.cuda.synchronize()
    res[fused] = (losses, torch.cat([p.detach().float().flatten() for p in model.parameters()]))
  tp_fused.USE_FUSED = True
  dl = max(abs(a - b) for a, b in zip(res[True][0], res[False][0]))
  dp = (res[True][1] - res[False][1]).abs().max().item()
  log("TP training, fused vs NCCL+GEMM: losses %s vs %s (max diff %.2e), max |dparam| %.2e" % (res[True][0], res[False][0], dl, dp))
  assert dl < 0.02 and dp < 8e-3


def check_moe(dev, world, rank):
  """K5: the peer-store all-to-all (csrc/symm.cu alltoall_p2p_kernel) vs the NCCL all-to-all, values and timing; then the
  expert-parallel MoE layer forward/backward with either transport."""
  from easyparallellibrary_b200.ops import moe
  from easyparallellibrary_b200.ops import tensor_parallel as tp
  epl.init(epl.Config({"cluster.colocate_split_and_replicate": True}))
  with epl.split(world):
    group = tp.current_tp_group()
  torch.manual_seed(11 + rank)
  for rows in (world * 64, world * 4096):
    t = torch.randn(rows, 512, device=dev).bfloat16()
    outs = {}
    for p2p in (False, True):
      moe.USE_P2P_KERNEL = p2p
      for _ in range(3):                                  # repeated calls exercise the epoch / buffer-reuse protocol
        outs[p2p] = moe.expert_all_to_all_raw(t, group)
      torch.cuda.synchronize()
    diff = (outs[True].float() - outs[False].float()).abs().max().item()
    times = {}
    for p2p in (False, True):
      moe.USE_P2P_KERNEL = p2p
      torch.cuda.synchronize(); dist.barrier()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      for _ in range(10):
        moe.expert_all_to_all_raw(t, group)
      e1.record(); torch.cuda.synchronize()
      times[p2p] = e0.elapsed_time(e1) / 10
    log("all-to-all %d x 512 bf16: |p2p - nccl| = %.1e ; nccl %.3f ms, p2p kernel %.3f ms" % (rows, diff, times[False], times[True]))
    assert diff == 0.0
  moe.USE_P2P_KERNEL = True
  torch.manual_seed(5)
  with epl.split(world):
    layer = moe.MoEFFN(256, 512, 2 * world, capacity_factor=2.0).to(dev)
  res = {}
  for p2p in (False, True):
    moe.USE_P2P_KERNEL = p2p
    torch.manual_seed(100 + rank)
    x = torch.randn(2, 64, 256, device=dev, requires_grad=True)
    y = layer(x)
    layer.zero_grad()
    (y.square().sum() + layer.aux_loss).backward()
    res[p2p] = (y.detach().clone(), x.grad.clone(), layer.wi.grad.clone())
  d = max((a - b).abs().max().item() for a, b in zip(res[True], res[False]))
  log("MoE layer fwd/bwd with p2p vs nccl transport: max diff %.2e" % d)
  assert d < 1e-5
  moe.USE_P2P_KERNEL = True


def check_clip(dev, world, rank):
  """K1 with gradient clipping (clip-then-reduce: the local coefficient is applied to the bucket before the kernel reads it)
  vs the NCCL + separate optimizer path with the same clipping."""
  res = {}
  for fused in (False, True):
    from easyparallellibrary_b200.models.gpt2 import GPT2, GPT2Config
    epl.init(epl.Config({"amp.level": "bf16", "communication.fused_kernels": fused}))
    torch.manual_seed(0)
    with epl.replicate(1):
      cfg = GPT2Config.named("tiny", n_embd=256, n_head=4, vocab_size=2048)
      model = GPT2(cfg)
    tr = epl.Trainer(model, "adamw", lr=1e-2, eps=1.0, max_grad_norm=0.05).build()
    assert (tr.fused is not None) == fused
    g = torch.Generator().manual_seed(100 + rank)
    out = None
    for _ in range(4):
      t = torch.randint(0, cfg.vocab_size, (4, 128), generator=g).to(dev)
      out = tr.step(t, t)
    torch.cuda.synchronize()
    res[fused] = (torch.cat([b.flat_param.float().flatten() for b in tr.flats[0].buckets]).clone(), float(out.grad_norm), out.item())
  diff = (res[True][0] - res[False][0]).abs().max().item()
  log("clip: fused vs NCCL path: max |dparam| = %.3e, grad norm %.4f vs %.4f, loss %.4f vs %.4f" % (
      diff, res[True][1], res[False][1], res[True][2], res[False][2]))
  assert res[True][1] > 0.05 and diff < 8e-3 and abs(res[True][2] - res[False][2]) < 0.03


def check_zero3(dev, world, rank):
  """ZeRO-3 (+ prefetch, + recompute, + weight/optimizer offload) on GPUs vs plain data parallelism: same trajectory, and the
  persistent device bytes the engine reports are what the allocator holds between steps."""
  from easyparallellibrary_b200.models.gpt2 import GPT2, GPT2Config
  res = {}
  for name, conf in (("dp", {"communication.fused_kernels": False}), ("zero3", {"zero.level": "v3"}),
                     ("zero3+gc+offload", {"zero.level": "v3", "gradient_checkpoint.type": "auto", "offload.level": "v0"})):
    conf = dict(conf)
    conf["amp.level"] = "bf16"
    epl.init(epl.Config(conf))
    torch.manual_seed(0)
    with epl.replicate(1):
      cfg = GPT2Config.named("tiny", n_embd=256, n_head=4, vocab_size=2048, n_layer=4)
      model = GPT2(cfg)
    tr = epl.Trainer(model, "adamw", lr=1e-2, eps=1.0).build()
    g = torch.Generator().manual_seed(100 + rank)
    losses = []
    for _ in range(4):
      t = torch.randint(0, cfg.vocab_size, (4, 128), generator=g).to(dev)
      losses.append(tr.step(t, t).item())
    torch.cuda.synchronize()
    res[name] = losses
    if tr.zero3:
      z = next(iter(tr.zero3.values()))
      log("  %s: persistent device bytes per rank %.2f MB (model %.2f M params), units %d, prefetch %s" % (
          name, z.persistent_bytes() / 1e6, cfg.num_params / 1e6, len(z.units), z.prefetch))
    del tr, model
  log("zero3: losses dp %s | zero3 %s | zero3+gc+offload %s" % tuple([round(v, 4) for v in res[k]] for k in ("dp", "zero3", "zero3+gc+offload")))
  for k in ("zero3", "zero3+gc+offload"):
    assert max(abs(a - b) for a, b in zip(res["dp"], res[k])) < 0.03, (k, res)


def check_nvls(dev, world, rank):
  """NVLS / multimem: in-switch all-reduce (csrc/symm.cu::nvls_allreduce_kernel) vs NCCL, values and time — when the platform
  exposes multicast; otherwise the probe's verdict is reported and the check passes vacuously."""
  from easyparallellibrary_b200.runtime import nvls
  buf = nvls.probe(dev)
  if buf is None or not buf.supported:
    log("nvls: multicast NOT available on this platform (symmetric rendezvous %s) -> peer-pointer kernels remain the path" % (
        "ok" if buf is not None else "failed"))
    return
  for dtype, n in ((torch.bfloat16, 1 << 16), (torch.float32, 1 << 14), (torch.bfloat16, 64 << 20)):
    nb = n * (2 if dtype == torch.bfloat16 else 4)
    big = nvls.MulticastBuffer(nb, dev) if nb > buf.nbytes else buf
    t = big.tensor(dtype, n)
    gen = torch.Generator(device=dev).manual_seed(5 + rank)
    src = (torch.randn(n, device=dev, generator=gen) * 0.1).to(dtype)
    ref = src.float().clone()
    dist.all_reduce(ref)
    t.copy_(src)
    torch.cuda.synchronize(); dist.barrier()
    big.all_reduce_(dtype, n, blocks=148 if nb > (1 << 22) else 16)
    torch.cuda.synchronize(); dist.barrier()
    err = (t.float() - ref).abs().max().item()
    tol = 1e-6 if dtype == torch.float32 else 2e-2 * ref.abs().max().item()
    times = {}
    nccl_t = src.clone()
    for name, fn in (("nvls", lambda: big.all_reduce_(dtype, n, blocks=148 if nb > (1 << 22) else 16)), ("nccl", lambda: dist.all_reduce(nccl_t))):
      for _ in range(3):
        fn()
      torch.cuda.synchronize(); dist.barrier()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      for _ in range(10):
        fn()
      e1.record(); torch.cuda.synchronize()
      tt = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
      dist.all_reduce(tt, op=dist.ReduceOp.MAX)
      times[name] = tt.item()
    log("nvls all-reduce %s x %d: max err %.2e ; multimem kernel %.4f ms vs NCCL %.4f ms (busbw %.0f GB/s)" % (
        str(dtype).split(".")[-1], n, err, times["nvls"], times["nccl"], 2 * (world - 1) / world * nb / times["nvls"] / 1e6))
    assert err <= tol, (err, tol)


def main():
  dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
  rank, world = dist.get_rank(), dist.get_world_size()
  dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
  torch.cuda.set_device(dev)
  what = sys.argv[1:] or ["native", "symm", "fused"]
  table = [("native", check_native), ("symm", check_symm), ("k1", check_k1), ("k1bench", bench_k1), ("fused", check_fused),
           ("clip", check_clip), ("tp", check_tp), ("tptrain", check_tp_train), ("moe", check_moe), ("zero3", check_zero3), ("nvls", check_nvls), ("k1nvls", check_fused_nvls)]
  if "all" in what:
    # the NVLS checks have only been validated at 2 ranks: beyond that they must be asked for by name (a hang inside "all"
    # would take the asserted checks of tests/test_multigpu.py down with it)
    skip = {"k1bench"} | ({"nvls", "k1nvls"} if world > 2 and os.environ.get("EPL_CHECK_NVLS", "0") != "1" else set())
    what = [n for n, _ in table if n not in skip]
  failed = []
  for name, fn in table:
    if name in what:
      try:                                             # a failing check (same exception on every rank) must not hide the later ones
        fn(dev, world, rank)
        torch.cuda.synchronize()
        ok = 1
      except Exception:
        import traceback
        traceback.print_exc()
        ok = 0
      t = torch.tensor([ok], device=dev)
      dist.all_reduce(t, op=dist.ReduceOp.MIN)
      if int(t.item()):
        log("CHECK %s PASSED (world %d)" % (name, world))
      else:
        failed.append(name)
        log("CHECK %s FAILED (world %d)" % (name, world))
  dist.barrier()
  if failed:
    log("MGPU CHECK FAILED: %s" % ", ".join(failed))
    dist.destroy_process_group()
    sys.exit(1)
  log("MGPU CHECK PASSED")
  dist.destroy_process_group()


if __name__ == "__main__":
  main()
