"""Cluster layouts with fabricated world sizes (the reference mocks Cluster.available_gpus and TF_CONFIG,
tests/cluster_test.py:32-248)."""
import json

import pytest

import easyparallellibrary_b200 as epl
from easyparallellibrary_b200.cluster import Cluster


def hosts(n):
  return ",".join("127.0.0.1:%d" % (8000 + i) for i in range(n))


def test_all_layout_every_gpu_is_a_replica():
  c = Cluster(worker_hosts=hosts(2), worker_index=1, gpus_per_worker=4, layout="all")
  vd = c.virtual_devices[0]
  assert vd.num_replicas == 8 and vd.num_devices_per_replica == 1
  assert [d.rank for d in vd.all_devices] == list(range(8))
  assert [d.rank for d in vd.local_devices] == [4, 5, 6, 7]
  assert c.total_gpu_num == 8 and c.worker_num == 2 and c.gpu_num_per_worker == 4


def test_auto_layout_two_stages_four_gpus_row_major():
  # reference strategy_new_test.py:35-73: 4 GPUs, 2 stages -> [[GPU0],[GPU2]], [[GPU1],[GPU3]]
  c = Cluster(worker_hosts=hosts(1), worker_index=0, gpus_per_worker=4, layout={"auto": [1, 1]})
  assert c.virtual_devices[0].ranks() == [[0], [2]]
  assert c.virtual_devices[1].ranks() == [[1], [3]]


def test_auto_layout_column_major_spreads_replica_over_workers():
  c = Cluster(worker_hosts=hosts(2), worker_index=0, gpus_per_worker=2, layout=None, prefer_intra_node=False)
  vds = c.generate_virtual_devices("auto", [1, 1])
  assert vds[0].ranks() == [[0], [1]] and vds[1].ranks() == [[2], [3]]
  c.set_prefer_intra_node(True)
  vds = c.generate_virtual_devices("auto", [1, 1])
  assert vds[0].ranks() == [[0], [2]] and vds[1].ranks() == [[1], [3]]


def test_auto_layout_not_divisible_raises():
  c = Cluster(worker_hosts=hosts(1), worker_index=0, gpus_per_worker=4)
  with pytest.raises(RuntimeError):
    c.generate_virtual_devices("auto", [3])


def test_multi_device_replicas():
  c = Cluster(worker_hosts=hosts(1), worker_index=0, gpus_per_worker=8, layout={"auto": [2, 2]})
  assert c.virtual_devices[0].ranks() == [[0, 1], [4, 5]]
  assert c.virtual_devices[1].ranks() == [[2, 3], [6, 7]]


def test_specific_layout():
  spec = [[["/job:worker/replica:0/task:0/device:GPU:0"], ["/job:worker/replica:0/task:1/device:GPU:0"]]]
  c = Cluster(worker_hosts=hosts(2), worker_index=0, gpus_per_worker=1, layout={"specific": spec})
  assert c.virtual_devices[0].ranks() == [[0], [1]]


def test_aware_row_layout_groups_hosts_by_machine():
  h = "10.0.0.1:1,10.0.0.2:1,10.0.0.1:2,10.0.0.2:2"
  c = Cluster(worker_hosts=h, worker_index=2, gpus_per_worker=1, layout={"aware_row": 2})
  assert c.hosts == "10.0.0.1:1,10.0.0.1:2,10.0.0.2:1,10.0.0.2:2"
  assert c.worker_index == 1
  assert [[d.rank for d in s] for s in c.virtual_devices[0].slice_devices] == [[0], [1]]
  assert len(c.virtual_devices) == 2
  with pytest.raises(RuntimeError):
    Cluster(worker_hosts=h, worker_index=0, gpus_per_worker=2, layout={"aware_row": 2})


def test_multiple_layouts_rejected():
  with pytest.raises(ValueError):
    Cluster(worker_hosts=hosts(1), gpus_per_worker=1, layout={"all": True, "auto": [1]})


def test_tf_config_with_chief(monkeypatch):
  cfg = {"cluster": {"chief": ["a:1"], "worker": ["b:1", "c:1"]}, "task": {"type": "worker", "index": 1}}
  monkeypatch.setenv("TF_CONFIG", json.dumps(cfg))
  c = Cluster(gpus_per_worker=1)
  assert c.worker_num == 3 and c.worker_index == 2 and c.hosts == "a:1,b:1,c:1"
  cfg["task"] = {"type": "chief", "index": 0}
  monkeypatch.setenv("TF_CONFIG", json.dumps(cfg))
  assert Cluster(gpus_per_worker=1).worker_index == 0


def test_torchrun_env(monkeypatch):
  monkeypatch.setenv("WORLD_SIZE", "8")
  monkeypatch.setenv("RANK", "5")
  monkeypatch.setenv("LOCAL_RANK", "5")
  monkeypatch.setenv("LOCAL_WORLD_SIZE", "8")
  c = Cluster()
  assert c.total_gpu_num == 8 and c.rank == 5 and c.worker_num == 1


def test_available_gpus_is_mockable(monkeypatch):
  monkeypatch.setattr(Cluster, "available_gpus", staticmethod(lambda: 6))
  c = Cluster(worker_hosts=hosts(2), worker_index=0)
  assert c.total_gpu_num == 12


def test_cluster_as_context_manager():
  epl.init(init_process_group=False)
  outer = epl.Env.get().cluster
  with Cluster(worker_hosts=hosts(1), gpus_per_worker=2, layout="all") as c:
    assert epl.Env.get().cluster is c
  assert epl.Env.get().cluster is outer
