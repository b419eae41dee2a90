"""tsp_gnn_b200 -- B200-native message-passing hot path of TSP-GNN behind the reference's
build_network / GraphNN / Mlp surface.  Compute lives in libtspgnn.so (hand-written sm_100a
CUDA, include/tspgnn.h); importing ``engine`` / running a Session loads it and fails loudly
if it is missing."""
from .mlp import Mlp
from .graphnn import GraphNN, LSTMStateTuple
from .model import build_network, Session, global_variables_initializer
from .instances import InstanceLoader, Incidence, create_batch, create_graph, read_graph, write_graph
from .params import load_weights, save_weights, init_params, param_spec

__all__ = ["Mlp", "GraphNN", "LSTMStateTuple", "build_network", "Session", "global_variables_initializer",
           "InstanceLoader", "Incidence", "create_batch", "create_graph", "read_graph", "write_graph",
           "load_weights", "save_weights", "init_params", "param_spec"]
