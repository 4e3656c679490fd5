"""``epl-launch``: spawn one process per GPU (reference ``epl/utils/launcher.py``).

  epl-launch --num_workers 2 --gpu_per_worker 4 [--machine_list ip:port,... --machine_rank i] script.py [args]

The reference starts one process per *worker* with ``TF_CONFIG`` and ``CUDA_VISIBLE_DEVICES``; here every GPU gets
its own process with the ``torch.distributed`` environment (RANK, LOCAL_RANK, WORLD_SIZE, LOCAL_WORLD_SIZE,
MASTER_ADDR, MASTER_PORT) *and* an EPL-style ``TF_CONFIG`` for scripts that still read it.  stderr of every rank is
written to ``<log_dir>/stderr_<rank>.log`` like the reference; a failed rank tears the job down (by PID, never by
pattern).
"""
from __future__ import annotations

import argparse
import json
import os
import signal
import socket
import subprocess
import sys
import time
from typing import List


def _free_port() -> int:
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  port = s.getsockname()[1]
  s.close()
  return port


def parse(argv=None):
  ap = argparse.ArgumentParser(prog="epl-launch")
  ap.add_argument("--num_workers", type=int, default=1, help="workers (nodes / launch groups) in the whole job")
  ap.add_argument("--gpu_per_worker", type=int, default=1)
  ap.add_argument("--machine_list", default="", help="ip:port of every machine, comma separated (multi-node)")
  ap.add_argument("--machine_rank", type=int, default=0)
  ap.add_argument("--debug", action="store_true")
  ap.add_argument("--log_dir", default=".")
  ap.add_argument("--backend", default="", help="force 'gloo' to run the plumbing on CPUs")
  ap.add_argument("--max_restarts", type=int, default=0,
                  help="relaunch the whole local group this many times after a failed attempt (EPL_RESTART_COUNT tells the script "
                       "which attempt it is; scripts resume from their last checkpoint, e.g. examples/bert/run_squad.py --resume)")
  ap.add_argument("script")
  ap.add_argument("script_args", nargs=argparse.REMAINDER)
  return ap.parse_args(argv)


def build_commands(args) -> List[dict]:
  machines = [m for m in args.machine_list.split(",") if m]
  if machines:
    master_addr, master_port = machines[0].rsplit(":", 1)
    workers_here = max(args.num_workers // len(machines), 1)
    first_worker = args.machine_rank * workers_here
  else:
    master_addr, master_port = "127.0.0.1", str(_free_port())
    workers_here, first_worker = args.num_workers, 0
  world = args.num_workers * args.gpu_per_worker
  worker_hosts = ["%s:%d" % (master_addr if not machines else machines[min(w * len(machines) // max(args.num_workers, 1), len(machines) - 1)].split(":")[0],
                             20000 + w) for w in range(args.num_workers)]
  cmds = []
  for w in range(first_worker, first_worker + workers_here):
    for g in range(args.gpu_per_worker):
      rank = w * args.gpu_per_worker + g
      local = (w - first_worker) * args.gpu_per_worker + g
      env = dict(os.environ)
      env.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(local), LOCAL_WORLD_SIZE=str(workers_here * args.gpu_per_worker),
                 GROUP_RANK=str(w), MASTER_ADDR=master_addr, MASTER_PORT=str(master_port),
                 TF_CONFIG=json.dumps({"cluster": {"worker": worker_hosts}, "task": {"type": "worker", "index": w}}))
      if args.backend == "gloo":
        env["CUDA_VISIBLE_DEVICES"] = ""
      runner = [sys.executable] if args.script.endswith(".py") else ["bash"]
      cmds.append({"rank": rank, "env": env, "argv": runner + [args.script] + list(args.script_args)})
  return cmds


def main(argv=None) -> int:
  """Run the job; after a failed attempt tear every process down (by exact PID) and, with ``--max_restarts``, start over with a
  fresh rendezvous port — the reference's ``run_script`` retry loop (``utils/launcher.py:168-188``) with recovery from the last
  checkpoint instead of a blind re-run."""
  args = parse(argv)
  rc = 0
  for attempt in range(max(args.max_restarts, 0) + 1):
    rc = _run_once(args, attempt)
    if rc == 0:
      break
    if attempt < args.max_restarts:
      print("[epl-launch] attempt %d failed with exit code %d; restarting (%d left)" % (attempt, rc, args.max_restarts - attempt),
            file=sys.stderr, flush=True)
      time.sleep(1.0)
  return rc


def _run_once(args, attempt: int) -> int:
  os.makedirs(args.log_dir, exist_ok=True)
  procs = []
  for c in build_commands(args):
    c["env"]["EPL_RESTART_COUNT"] = str(attempt)
    err = open(os.path.join(args.log_dir, "stderr_%d.log" % c["rank"]), "a" if attempt else "w")
    procs.append((c["rank"], subprocess.Popen(c["argv"], env=c["env"], stderr=err if not args.debug else None), err))
  rc = 0
  try:
    alive = {r for r, _, _ in procs}
    while alive:
      for r, p, _ in procs:
        if r in alive and p.poll() is not None:
          alive.discard(r)
          if p.returncode != 0:
            rc = rc or p.returncode       # report the first failure, not the SIGTERM of the ranks torn down because of it
            for _, q, _ in procs:          # tear down by exact PID
              if q.poll() is None:
                q.send_signal(signal.SIGTERM)
      time.sleep(0.2)
  finally:
    for _, p, err in procs:
      if p.poll() is None:
        p.kill()
      err.close()
  return rc


if __name__ == "__main__":
  sys.exit(main())
