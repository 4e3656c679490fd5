"""The code blocks of docs/api_examples.md marked ``<!-- runnable -->`` run as written (single CPU process)."""
import os
import re

import pytest

import easyparallellibrary_b200 as epl

DOC = os.path.join(os.path.dirname(__file__), "..", "docs", "api_examples.md")


def _blocks():
  text = open(DOC).read()
  out = []
  for m in re.finditer(r"<!-- runnable -->\s*```python\n(.*?)```", text, re.S):
    title = text[:m.start()].rsplit("\n## ", 1)[-1].split("\n", 1)[0].strip()
    out.append(pytest.param(m.group(1), id=re.sub(r"[^a-z0-9]+", "-", title.lower()).strip("-")))
  return out


def test_doc_has_runnable_blocks():
  assert len(_blocks()) >= 5


@pytest.mark.parametrize("code", _blocks())
def test_api_example_runs(code):
  try:
    exec(compile(code, DOC, "exec"), {"__name__": "__docs__"})
  finally:
    epl.shutdown()
