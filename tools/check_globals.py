"""Static check for names that do not exist: every LOAD_GLOBAL of every code object of every module of the package (plus bench.py,
tools/, examples/) must resolve to a module global or a builtin.  GPU-only branches never run in the CPU test tiers; a misspelt
helper or a forgotten import in one of them would only show up on a B200.  (pyflakes / pylint are not in the image.)

  python tools/check_globals.py          # exit code 1 and one line per unresolved name
"""
import builtins
import dis
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def code_objects(code):
  yield code
  for c in code.co_consts:
    if isinstance(c, types.CodeType):
      yield from code_objects(c)


def check_source(path, module_globals=None):
  src = open(path).read()
  code = compile(src, path, "exec")
  names = set(module_globals or ())
  # names bound anywhere at module level (imports, defs, assignments), also inside if/try blocks
  for c in [code]:
    for ins in dis.get_instructions(c):
      if ins.opname in ("STORE_NAME", "STORE_GLOBAL", "IMPORT_NAME", "IMPORT_FROM"):
        names.add(ins.argval.split(".")[0] if ins.opname == "IMPORT_NAME" else ins.argval)
  for c in code_objects(code):
    for ins in dis.get_instructions(c):
      if ins.opname == "STORE_GLOBAL":
        names.add(ins.argval)
  bad = []
  for c in code_objects(code):
    local = {i.argval for i in dis.get_instructions(c) if i.opname == "STORE_NAME"}       # class bodies bind with STORE_NAME
    for ins in dis.get_instructions(c):
      if ins.opname in ("LOAD_GLOBAL", "LOAD_NAME") and ins.argval not in names and ins.argval not in local and not hasattr(builtins, ins.argval):
        if ins.argval in ("__file__", "__name__", "__doc__", "__spec__", "__builtins__", "__package__", "__path__", "__class__"):
          continue
        bad.append((path, c.co_name, ins.positions.lineno if ins.positions else c.co_firstlineno, ins.argval))
  return bad


def check_native_symbols(files):
  """Every ``epl_*`` function a Python file calls through ctypes must be exported by one of the in-tree libraries."""
  import re
  import subprocess
  exported = set()
  libdir = os.path.join(ROOT, "easyparallellibrary_b200", "lib")
  for lib in ("libepl_kernels.so", "libepl_runtime.so"):
    path = os.path.join(libdir, lib)
    if not os.path.exists(path):
      print("(%s not built: native symbol check skipped)" % lib)
      return []
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True).stdout
    exported |= {l.split()[-1] for l in out.splitlines() if l.strip()}
  bad = []
  pat = re.compile(r"(?:lib|_lib|self\.lib|L|self\._lib)\(?\)?\.(epl_[a-z0-9_]+)")
  for f in files:
    for n, line in enumerate(open(f), 1):
      for name in pat.findall(line):
        if name not in exported:
          bad.append((f, n, name))
  return bad


def main():
  files = []
  for base in ("easyparallellibrary_b200", "tools", "examples", "baseline", "epl"):
    for d, _, fs in os.walk(os.path.join(ROOT, base)):
      if "_ref" in d or "__pycache__" in d:
        continue
      files += [os.path.join(d, f) for f in fs if f.endswith(".py")]
  files += [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
  bad = []
  for f in sorted(files):
    bad += check_source(f)
  for path, fn, line, name in bad:
    print("%s:%d: in %s: name %r is not defined at module level" % (os.path.relpath(path, ROOT), line, fn, name))
  missing = check_native_symbols(files)
  for path, line, name in missing:
    print("%s:%d: native function %r is not exported by lib/libepl_*.so" % (os.path.relpath(path, ROOT), line, name))
  print("%d files checked, %d unresolved names, %d missing native symbols" % (len(files), len(bad), len(missing)))
  return 1 if (bad or missing) else 0


if __name__ == "__main__":
  sys.exit(main())
