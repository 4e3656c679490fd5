"""``import epl`` — drop-in spelling for users of the reference; everything lives in ``easyparallellibrary_b200``."""
from easyparallellibrary_b200 import *  # noqa: F401,F403
from easyparallellibrary_b200 import (Cluster, Config, Env, Graph, GraphKeys, ModelPhase, VERSION, add_to_collection,  # noqa: F401
                                      get_all_collections, get_collection, init, replicate, set_default_strategy, split)
import easyparallellibrary_b200 as _impl


def __getattr__(name):
  return getattr(_impl, name)
