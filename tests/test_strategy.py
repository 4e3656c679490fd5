"""Strategy scopes: nesting rules, default strategy, taskgraph creation (reference strategy_test.py:46-80)."""
import pytest
import torch
from torch import nn

import easyparallellibrary_b200 as epl


def setup_function(_):
  epl.init(init_process_group=False)


def test_nesting_rules():
  with pytest.raises(RuntimeError):
    with epl.replicate(1):
      with epl.replicate(1):
        pass
  with pytest.raises(RuntimeError):
    with epl.split(2):
      with epl.replicate(1):
        pass
  with pytest.raises(RuntimeError):
    with epl.replicate(1):
      with epl.split(2):
        pass
  with pytest.raises(RuntimeError):
    with epl.split(2):
      with epl.split(2):
        pass


def test_scopes_create_taskgraphs_in_order():
  with epl.replicate(1, name="s0"):
    a = nn.Linear(4, 4)
  with epl.replicate(1, name="s1"):
    b = nn.Linear(4, 4)
  with epl.split(2, name="head"):
    c = nn.Linear(4, 8)
  g = epl.Graph.get()
  assert len(g.taskgraphs) == 3
  assert g.taskgraph_of(a).index == 0 and g.taskgraph_of(b).index == 1 and g.taskgraph_of(c).index == 2
  assert g.taskgraphs[2].is_split and g.taskgraphs[2].num_device_per_replica == 2
  assert g.taskgraph_of(a.weight).index == 0
  assert g.num_stages == 2 and not g.pipeline_enabled
  assert "Taskgraph 1" in g.format()


def test_default_strategy_and_restaging():
  epl.set_default_strategy(epl.replicate(1))
  a = nn.Linear(2, 2)
  epl.set_default_strategy(epl.replicate(1))      # BERT example idiom: next stage
  b = nn.Linear(2, 2)
  with epl.split(2):
    c = nn.Linear(2, 2)                           # explicit scope suspends the default
  d = nn.Linear(2, 2)
  g = epl.Graph.get()
  assert [g.taskgraph_of(m).index for m in (a, b, c, d)] == [0, 1, 2, 1]
  with pytest.raises(ValueError):
    epl.set_default_strategy(epl.split(2))


def test_tied_parameters_stay_in_first_taskgraph():
  with epl.replicate(1):
    emb = nn.Embedding(10, 4)
  with epl.replicate(1):
    head = nn.Linear(4, 10, bias=False)
    head.weight = emb.weight
  g = epl.Graph.get()
  assert g.taskgraph_of(emb.weight).index == 0
  assert len(g.taskgraphs[1].parameters) == 1     # only the Linear's original weight was tagged there


def test_pipeline_enabled_needs_micro_batches():
  epl.init({"pipeline.num_micro_batch": 4}, init_process_group=False)
  with epl.replicate(1):
    nn.Linear(2, 2)
  with epl.replicate(1):
    nn.Linear(2, 2)
  assert epl.Graph.get().pipeline_enabled


def test_annotations_forbidden_under_auto_parallel():
  epl.init({"auto.auto_parallel": True}, init_process_group=False)
  with pytest.raises(RuntimeError):
    with epl.replicate(1):
      pass


def test_collections_api():
  t = torch.ones(3)
  epl.add_to_collection(t, epl.GraphKeys.GLOBAL_MEAN_OBJECTS)
  epl.add_to_collection([t, t], epl.GraphKeys.LOCAL_SUM_OBJECTS)
  assert len(epl.get_collection(epl.GraphKeys.GLOBAL_MEAN_OBJECTS)) == 1
  assert len(epl.get_all_collections()) == 3
  with pytest.raises(ValueError):
    epl.add_to_collection(t, "nope")
