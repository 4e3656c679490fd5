"""Per-kernel counts of the Blackwell-native SASS mnemonics in the built kernel library (runs on a CPU-only box):

  python tools/sass_mnemonics.py > profiles/r2_sass_mnemonics.txt

UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMAREDG = TMA tensor load / reduce-store, UBLKCP = cp.async.bulk,
LDGMC = multimem.ld_reduce (multimem.st assembles to a plain STG.E.128.STRONG.SYS on the multicast address), SYNCS = mbarrier,
UTCBAR = tcgen05.commit.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "easyparallellibrary_b200", "lib", "libepl_kernels.so")
PAT = re.compile(r"\b(UTC[A-Z]*MMA(?:\.2CTA)?|UTCBAR(?:\.2CTA)?|LDTM|STTM|UTMALDG|UTMAREDG|UTMASTG|UBLKCP|LDGMC|STGMC|SYNCS)\b")


def main():
  sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
  names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
  counts, order, cur, k = {}, [], None, 0
  for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
      cur = names[k]
      k += 1
      counts[cur] = collections.Counter()
      order.append(cur)
      continue
    if cur is None:
      continue
    for op in PAT.findall(line.split("/*")[1] if line.strip().startswith("/*") and line.count("/*") > 1 else line):
      counts[cur][op] += 1
  print("Blackwell-native SASS mnemonics per kernel in lib/libepl_kernels.so (round 2; cuobjdump -sass, counts of instructions)")
  print("UTC*MMA = tcgen05.mma (UTCHMMA f16/bf16, UTCQMMA fp8), LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMAREDG = TMA tensor load / reduce-store,")
  print("UBLKCP = cp.async.bulk (linear bulk copies: K1 v2 ring, K3 copy CTAs), LDGMC = multimem.ld_reduce (NVLS; multimem.st assembles to STG.E.128.STRONG.SYS on the multicast address), SYNCS = mbarrier\n")
  rows = [(n, c) for n, c in counts.items() if c]
  rows.sort(key=lambda r: -sum(r[1].values()))
  for n, c in rows:
    print("%-112s %s" % (n[:112], "  ".join("%s=%d" % (op, v) for op, v in sorted(c.items()))))
  print("\n%d of %d kernels use at least one of these instructions" % (len(rows), len(counts)))


if __name__ == "__main__":
  sys.exit(main())
