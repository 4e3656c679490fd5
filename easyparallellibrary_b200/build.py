"""In-tree native build: every ``csrc/*.cu`` -> ``lib/libepl_kernels.so`` (sm_100a only),
every ``csrc/*.cpp`` -> ``lib/libepl_runtime.so``.

The kernels expose a plain C ABI (raw pointers + stream) and are loaded with
``ctypes`` (``ops/_lib.py``): no torch headers, so a full rebuild is seconds and
nvcc cross-compiles on a GPU-less box.  The built ``.so`` files are git-ignored
but travel with the source tree.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib")
OBJ = os.path.join(LIB, "obj")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
CXX_FLAGS = ["-O3", "-std=c++17", "-fPIC", "-Wall", "-Werror=return-type", "-fstack-protector"]


def _nvcc() -> str:
  for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
    if cand and (os.path.sep not in cand or os.path.exists(cand)):
      return cand
  return "nvcc"


def _digest(path: str, extra: str) -> str:
  h = hashlib.sha1(extra.encode())
  with open(path, "rb") as f:
    h.update(f.read())
  for hdr in sorted(os.listdir(CSRC)):
    if hdr.endswith((".cuh", ".h", ".hpp")):
      with open(os.path.join(CSRC, hdr), "rb") as f:
        h.update(f.read())
  return h.hexdigest()


def _compile_one(src: str, verbose: bool) -> str:
  name = os.path.splitext(os.path.basename(src))[0]
  obj = os.path.join(OBJ, name + ".o")
  stamp = obj + ".sha1"
  is_cu = src.endswith(".cu")
  cmd = ([_nvcc()] + NVCC_FLAGS if is_cu else ["g++"] + CXX_FLAGS + ["-I/usr/local/cuda/include"]) + ["-I", CSRC, "-c", src, "-o", obj]
  dig = _digest(src, " ".join(cmd))
  if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
    return obj
  if verbose:
    print("[epl build]", " ".join(cmd), file=sys.stderr)
  subprocess.run(cmd, check=True)
  with open(stamp, "w") as f:
    f.write(dig)
  return obj


def build_all(verbose: bool = False, force: bool = False) -> dict:
  os.makedirs(OBJ, exist_ok=True)
  if force:
    for f in os.listdir(OBJ):
      os.remove(os.path.join(OBJ, f))
  cus = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
  cpps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cpp"))
  with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
    cu_objs = list(ex.map(lambda s: _compile_one(s, verbose), cus))
    cpp_objs = list(ex.map(lambda s: _compile_one(s, verbose), cpps))
  out = {}
  if cu_objs:
    so = os.path.join(LIB, "libepl_kernels.so")
    if force or not os.path.exists(so) or any(os.path.getmtime(o) > os.path.getmtime(so) for o in cu_objs):
      subprocess.run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", so] + cu_objs + ["-lcudart"], check=True)
    out["kernels"] = so
  if cpp_objs:
    so = os.path.join(LIB, "libepl_runtime.so")
    if force or not os.path.exists(so) or any(os.path.getmtime(o) > os.path.getmtime(so) for o in cpp_objs):
      subprocess.run(["g++", "-shared", "-o", so] + cpp_objs + ["-L/usr/local/cuda/lib64", "-lcudart", "-ldl", "-lpthread"], check=True)
    out["runtime"] = so
  return out


if __name__ == "__main__":
  print(build_all(verbose=True, force="--force" in sys.argv))
