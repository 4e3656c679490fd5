"""Multi-GPU checks (need >= 2 CUDA devices), each compared against ground truth computed in fp32 / with library calls:

* ``native``  in-tree NCCL communicator verbs          * ``symm``   NVLink symmetric memory + device barrier
* ``k1``      fused reduce-scatter+AdamW+all-gather kernel vs an fp32 reduction in rank order + fp32 AdamW: the reduction
              and the rounded weights must match BIT FOR BIT, m / v / master to fp32 rounding; bf16, fp16 + loss scale,
              ragged buckets, grid size changing between launches
* ``fused``   GPT-2 training through K1 (captured in a CUDA graph, bucket kernels overlapped with backward) vs NCCL path
* ``clip``    K1 + gradient clipping vs the NCCL path    * ``tp``     fused all-gather->GEMM / GEMM->reduce-scatter / weight-gather GEMM
* ``tptrain`` tensor-parallel BERT training, fused vs NCCL+GEMM   * ``moe``  K5 / K5b all-to-all kernels and the MoE layer
* ``zero3``   ZeRO-3 (+ recompute + offload) vs data parallelism
* pipelines:  2 stages, and 2 stages x 2 replicas, through ``bench.py``

One ``torchrun`` per world size runs every check in one process group (``tools/mgpu_check.py all``) and prints a
``CHECK <name> PASSED`` line per check; the tests below assert on those lines, so an 8-GPU box pays the start-up cost three
times (8 ranks, then 4 and 2 ranks concurrently on disjoint GPUs), not once per test.
"""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NGPU = torch.cuda.device_count() if torch.cuda.is_available() else 0
WORLDS = [w for w in (2, 4, 8) if w <= NGPU]
CHECKS = ["native", "symm", "k1", "fused", "clip", "tp", "tptrain", "moe", "zero3"]


def _torchrun(script_args, nproc, port, env=None, timeout=900, wait=True):
  cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
         "--master-port", str(port)] + script_args
  e = dict(os.environ)
  e.update(env or {})
  if wait:
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout, env=e)
  return subprocess.Popen(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=e)


@pytest.fixture(scope="session")
def mgpu_logs():
  """world size -> combined output of ``mgpu_check.py all``."""
  logs = {}
  env = {"EPL_CUDA_GRAPH": "1", "EPL_FUSED_OVERLAP_MIN_WORLD": "2"}
  if 8 in WORLDS:
    r = _torchrun(["tools/mgpu_check.py", "all"], 8, 29611, env)
    logs[8] = r.stdout + r.stderr
  rest = [w for w in WORLDS if w != 8]
  procs, first = [], 0
  for w in sorted(rest, reverse=True):                 # 4 and 2 ranks side by side on disjoint GPUs when the box has them
    if first + w > NGPU:
      for ww, p in procs:
        logs[ww] = p.communicate(timeout=900)[0]
      procs, first = [], 0
    e = dict(env, CUDA_VISIBLE_DEVICES=",".join(str(i) for i in range(first, first + w)))
    procs.append((w, _torchrun(["tools/mgpu_check.py", "all"], w, 29620 + w, e, wait=False)))
    first += w
  for ww, p in procs:
    logs[ww] = p.communicate(timeout=900)[0]
  os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
  for w, text in logs.items():
    with open(os.path.join(ROOT, "gpurun_out", "pytest_mgpu_check_w%d.log" % w), "w") as f:
      f.write(text)
  return logs


@pytest.mark.skipif(NGPU < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("check", CHECKS)
def test_mgpu_check(mgpu_logs, check, world):
  if world not in mgpu_logs:
    pytest.skip("needs %d GPUs" % world)
  text = mgpu_logs[world]
  assert "CHECK %s PASSED (world %d)" % (check, world) in text, "\n".join(
      l for l in text.splitlines() if not l.startswith(("W0", "***")) and "OMP_NUM" not in l)[-3000:]


@pytest.mark.skipif(NGPU < 2, reason="needs 2 GPUs")
def test_pipeline_two_stages_gpu():
  r = _torchrun(["bench.py", "--gpus", "2", "--steps", "2", "--warmup", "3", "--parallelism", "pp2", "--model", "small", "--batch", "2",
                 "--seq", "256", "--no-e2e"], 2, 29615)
  assert r.returncode == 0 and '"parallelism": "dp1xpp2"' in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.skipif(NGPU < 4, reason="needs 4 GPUs")
def test_pipeline_times_data_parallel_gpu():
  """2 stages x 2 replicas: the symmetric buckets of each stage's data-parallel group are exchanged among a rank SUBSET."""
  r = _torchrun(["bench.py", "--gpus", "4", "--steps", "2", "--warmup", "3", "--parallelism", "pp2", "--model", "small", "--batch", "2",
                 "--seq", "256", "--no-e2e"], 4, 29616)
  assert r.returncode == 0 and '"parallelism": "dp2xpp2"' in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
