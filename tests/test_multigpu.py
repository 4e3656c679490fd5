"""Multi-GPU checks (need >= 2 CUDA devices): native NCCL communicator, NVLink symmetric memory, the fused
reduce-scatter+AdamW+all-gather kernel, the fused tensor-parallel GEMM kernels, and a 2-stage pipeline step.
Each runs under torchrun in a subprocess (one process per GPU)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(script_args, nproc, port, timeout=600):
  cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
         "--master-port", str(port)] + script_args
  return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("what,port", [("native", 29611), ("symm", 29612), ("fused", 29613), ("tp", 29614)])
def test_mgpu_check(what, port):
  # the fused data-parallel kernel was validated at 2 and 8 ranks; the other checks at 2 (round 1)
  n = min(torch.cuda.device_count(), 8) if what == "fused" else 2
  r = _torchrun(["tools/mgpu_check.py", what], n, port)
  assert r.returncode == 0 and "MGPU CHECK PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_pipeline_two_stages_gpu():
  r = _torchrun(["bench.py", "--gpus", "2", "--steps", "2", "--warmup", "3", "--parallelism", "pp2", "--model", "small", "--batch", "2",
                 "--seq", "256", "--no-e2e"], 2, 29615)
  assert r.returncode == 0 and '"parallelism": "dp1xpp2"' in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
