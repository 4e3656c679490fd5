from easyparallellibrary_b200.ir.graph import Graph, GraphKeys, add_to_collection, get_collection, get_all_collections
from easyparallellibrary_b200.ir.taskgraph import Taskgraph
from easyparallellibrary_b200.ir.node import Node, TensorMeta
from easyparallellibrary_b200.ir.phase import ModelPhase, current_phase, phase_scope
