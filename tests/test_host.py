"""Host-side logic: parameter inventory, declarative checks of GraphNN/Mlp, data plane,
sharding (incl. a world_size-2 gloo run) and the C-ABI symbol table.  CPU only."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import tspgnn_oracle as orc
import tsp_gnn_b200 as tg
from tsp_gnn_b200 import instances as inst
from tsp_gnn_b200 import params as P
from tsp_gnn_b200 import sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_param_spec_identical_to_oracle_inventory():
    assert [(n, tuple(s), k) for n, s, k in P.param_spec(64)] == [(n, tuple(s), k) for n, s, k in orc.param_spec(64)]
    table, total = P.param_offsets(64)
    assert total == 115529


def test_flatten_roundtrip_and_validation():
    p = P.init_params(64, seed=3)
    blob = P.flatten(p)
    q = P.unflatten(blob)
    assert all(np.array_equal(p[k], q[k]) for k in p)
    bad = dict(p); bad["V_init"] = np.zeros((1, 32), np.float32)
    with pytest.raises(ValueError):
        P.flatten(bad)
    del bad["V_init"]
    with pytest.raises(KeyError):
        P.flatten(bad)


def test_checkpoint_roundtrip_and_missing_path(tmp_path):
    p = P.init_params(64, seed=5)
    P.save_weights(p, str(tmp_path / "epoch=100"))
    q = P.load_weights(str(tmp_path / "epoch=100"))
    assert set(q) == set(p) and all(np.array_equal(p[k], q[k]) for k in p)
    with pytest.raises(Exception, match="Path does not exist!"):     # util.py:20
        P.load_weights(str(tmp_path / "nope"))


def test_mlp_layer_list_semantics():
    m = tg.Mlp(layer_sizes=[64 / 8, 64 / 4, 64 / 2], activations=["relu"] * 3, output_size=64, name="E_init_MLP")
    assert m.layer_sizes() == [8, 16, 32, 64]                       # model.py:34 float sizes cast
    assert [l["activation"] for l in m.layers] == ["relu", "relu", "relu", None]
    assert m.layers[0]["name"] == "E_init_MLP_MLP_layer_1"          # mlp.py:36-38
    m2 = tg.Mlp(layer_sizes=[4, 4], activations="relu", name="x")   # mlp.py:26-28
    assert [l["activation"] for l in m2.layers] == ["relu", "relu"]


def _tsp_args(d=64):
    return ({"V": d, "E": d}, {"EV": ("E", "V")}, {"V_msg_E": ("V", "E"), "E_msg_V": ("E", "V")},
            {"V": [{"mat": "EV", "msg": "E_msg_V", "transpose?": True, "var": "E"}],
             "E": [{"mat": "EV", "msg": "V_msg_E", "var": "V"}]})


def test_graphnn_check_model_messages():
    var, mat, msg, loop = _tsp_args()
    g = tg.GraphNN(var, mat, msg, loop, name="TSP")
    assert set(g.variable_names()) == {n for n, _, _ in P.param_spec(64) if n.startswith("TSP/")}
    with pytest.raises(Warning, match="Variable E is not updated anywhere"):          # graphnn.py:76
        tg.GraphNN(var, mat, msg, {"V": loop["V"]})
    with pytest.raises(Exception, match="Updating variable X, which has not been declared"):
        tg.GraphNN(var, mat, msg, dict(loop, X=[]))
    with pytest.raises(Exception, match="Matrix EV definition depends on undeclared variable Q"):
        tg.GraphNN(var, {"EV": ("Q", "V")}, msg, loop)
    with pytest.raises(Exception, match="Message V_msg_E maps to undeclared variable Z"):
        tg.GraphNN(var, mat, dict(msg, V_msg_E=("V", "Z")), loop)


def test_graphnn_other_topologies_take_the_generic_path():
    """Only the TSP wiring maps onto the fused kernels; everything else is executed op by op on the generic CUDA
    building blocks (graphnn.py:142-173) -- never rejected, never a CPU path."""
    var, mat, msg, loop = _tsp_args()
    assert tg.GraphNN(var, mat, msg, loop)._kernel_roles is not None
    assert tg.GraphNN(var, mat, msg, loop, MLP_depth=2)._kernel_roles is None
    assert tg.GraphNN({"V": 32, "E": 32}, mat, msg, loop)._kernel_roles is None
    loop2 = dict(loop, V=[{"mat": "EV", "msg": "E_msg_V", "var": "E"}])     # missing transpose
    assert tg.GraphNN(var, mat, msg, loop2)._kernel_roles is None
    # a model with a transfer function, a matrix-only input of integer width and two terms per variable
    g = tg.GraphNN({"A": 8, "B": 16, "C": 8}, {"M_AB": ("A", "B"), "M_CB": ("C", "B"), "F_C": ("C", 4)},
                   {"B2A": ("B", "A"), "A2B": ("A", "B"), "B2C": ("B", "C")},
                   {"A": [{"mat": "M_AB", "msg": "B2A", "var": "B"}, {"var": "A", "fun": lambda y: y * y}],
                    "B": [{"mat": "M_AB", "transpose?": True, "msg": "A2B", "var": "A"}],
                    "C": [{"mat": "M_CB", "msg": "B2C", "var": "B"}, {"mat": "F_C"}]}, name="TOY")
    assert g._kernel_roles is None
    assert (g.input_width("A"), g.input_width("B"), g.input_width("C")) == (16, 16, 12)
    p = g.init_parameters(seed=1)
    assert sorted(p) == sorted(g.variable_names())
    assert p["TOY/A_cell/layer_norm_basic_lstm_cell/kernel"].shape == (24, 32)
    assert p["TOY/C_cell/layer_norm_basic_lstm_cell/kernel"].shape == (20, 32)
    assert p["TOY/B2A_MLP_layer_4/kernel"].shape == (16, 8) and p["TOY/B2A_MLP_layer_4/bias"].shape == (8,)
    with pytest.raises(RuntimeError, match="uninitialized"):
        tg.Mlp([4, 4], name="m")(np.zeros((2, 3), dtype=np.float32))


def test_graphnn_check_run_shapes():
    var, mat, msg, loop = _tsp_args()
    g = tg.GraphNN(var, mat, msg, loop, name="TSP")
    EV = np.zeros((6, 4)); V0 = np.zeros((4, 64)); E0 = np.zeros((6, 64))
    g.check_run({"EV": EV}, {"V": V0, "E": E0}, 3, {})
    with pytest.raises(ValueError, match="Initial embedding of variable V doesn't have the same dimensionality 64"):
        g.check_run({"EV": EV}, {"V": np.zeros((4, 32)), "E": E0}, 3, {})
    with pytest.raises(ValueError, match="Matrix EV doesn't have the same number of nodes"):
        g.check_run({"EV": np.zeros((5, 4))}, {"V": V0, "E": E0}, 3, {})
    with pytest.raises(ValueError, match="LSTM's initial hidden state"):
        g.check_run({"EV": EV}, {"V": V0, "E": E0}, 3, {"V": np.zeros((3, 64))})


def test_build_network_key_set():
    GNN = tg.build_network(64)
    for k in ("gnn", "route_exists", "n_vertices", "n_edges", "EV", "W", "C", "time_steps", "last_states",
              "predictions", "TP", "FP", "TN", "FN", "acc", "loss", "train_step"):       # model.py:97-104,123,147-167
        assert k in GNN
    # other embedding sizes (train.py:108 -d) build too: they run on the generic CUDA path behind Session.run
    assert tg.build_network(32)["gnn"]._kernel_roles is None
    assert tg.build_network(64)["gnn"]._kernel_roles is not None
    with pytest.raises(ValueError):
        tg.build_network(4)


def test_graph_file_roundtrip_and_loader(tmp_path):
    d = tmp_path / "instances"
    inst.create_dataset(str(d), 5, 8, conn_min=0.5, conn_max=1.0, samples=6, seed=1)
    files = sorted(os.listdir(d))
    assert len(files) == 6
    Ma, Mw, route = inst.read_graph(str(d / "0.graph"))
    n = Ma.shape[0]
    assert np.all(np.tril(Ma) == 0) and sorted(route) == list(range(n))
    text = open(d / "0.graph").read()
    for kw in ("DIMENSION", "EDGE_DATA_SECTION", "EDGE_WEIGHT_SECTION", "TOUR_SECTION", "EOF"):
        assert kw in text
    loader = tg.InstanceLoader(str(d))
    batches = list(loader.get_batches(2, 0.02))
    assert len(batches) == 3
    EV, W, C, y, nv, ne = batches[0]
    assert list(y) == [0, 1, 0, 1] and nv[0] == nv[1] and nv[2] == nv[3]        # instance_loader.py:21-23,50
    e0 = int(ne[0])
    assert np.allclose(C[:e0] * (1 + 0.02) / (1 - 0.02), C[e0:2 * e0])           # -dev / +dev copies


def test_incidence_from_dense_validation():
    EV = inst.synth_batch([5, 4], seed=1)[0]
    back = tg.Incidence.from_dense(EV.toarray())
    assert np.array_equal(back.src, EV.src) and np.array_equal(back.dst, EV.dst) and back.shape == EV.shape
    bad = EV.toarray(); bad[0, :] = 0
    with pytest.raises(ValueError):
        tg.Incidence.from_dense(bad)


def test_partition_is_balanced_and_complete():
    rng = np.random.RandomState(0)
    n = rng.randint(20, 61, size=512)
    ne = n * (n - 1) // 2
    parts = sharding.partition_instances(ne, 8)
    allidx = np.sort(np.concatenate(parts))
    assert np.array_equal(allidx, np.arange(512))
    loads = np.array([ne[p].sum() for p in parts])
    assert loads.max() / loads.mean() < 1.01


def test_take_instances_builds_local_ids():
    EV, W, C, y, nv, ne = inst.synth_batch([5, 6, 4, 7], seed=3)
    src, dst, w, c, nvl, nel = sharding.take_instances(np.array([1, 3]), EV.src, EV.dst, W, C, nv, ne)
    assert list(nvl) == [6, 7] and len(src) == 15 + 21
    assert src[:15].max() < 6 and dst[15:].min() >= 6 and dst.max() == 12
    full = sharding.scatter_logits([1.0, 2.0], [1, 3], 4)
    assert list(full) == [0, 1, 0, 2]


_GLOO_WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, {root!r})
import torch.distributed as dist
from tsp_gnn_b200 import sharding
from tsp_gnn_b200 import instances as inst
from oracle import tspgnn_oracle as orc
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
EV, W, C, y, nv, ne = inst.synth_batch([5, 7, 6, 8, 5], seed=3)
params = orc.init_params(64, seed=2)
parts = sharding.partition_instances(ne, world)
src, dst, w, c, nvl, nel = sharding.take_instances(parts[rank], EV.src, EV.dst, W, C, nv, ne)
# the oracle stands in for the per-rank GPU engine: this test covers the sharding logic only
local = orc.forward(params, src, dst, w, c, nvl, nel, 3)["logits"] if len(parts[rank]) else np.zeros(0)
full = sharding.all_reduce_logits(local, parts[rank], len(ne)).numpy()
ref = orc.forward(params, EV.src, EV.dst, W, C, nv, ne, 3)["logits"]
assert np.abs(full - ref).max() < 1e-6, (full, ref)
if rank == 0:
    print("GLOO_OK", world)
dist.destroy_process_group()
"""


def test_world_size_2_gloo_logit_allreduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                         capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "GLOO_OK 2" in res.stdout


_GLOO_GRAD_WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, {root!r})
import torch
import torch.distributed as dist
from tsp_gnn_b200 import sharding, params as P
from tsp_gnn_b200 import instances as inst
from oracle import tspgnn_oracle as orc, tspgnn_oracle_grad as og
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
EV, W, C, y, nv, ne = inst.synth_batch([5, 7, 6, 8, 5], seed=3)
params = orc.init_params(64, seed=2)
parts = sharding.partition_instances(ne, world)
src, dst, w, c, nvl, nel = sharding.take_instances(parts[rank], EV.src, EV.dst, W, C, nv, ne)
# the oracle stands in for the per-rank GPU engine: this test covers the sharding / reduction logic
r = og.forward_backward(params, src, dst, w, c, nvl, nel, np.asarray(y)[parts[rank]], 3, global_batch=len(ne))
blob = torch.from_numpy(P.flatten({{k: v.astype(np.float32) for k, v in r["grads"].items()}}))
loss = torch.tensor([r["loss"]], dtype=torch.float32)
sharding.all_reduce_gradients(blob, loss)
full = og.forward_backward(params, EV.src, EV.dst, W, C, nv, ne, y, 3)
ref = P.flatten({{k: v.astype(np.float32) for k, v in full["grads"].items()}})
assert abs(float(loss[0]) - full["loss"]) < 1e-6
assert np.abs(blob.numpy() - ref).max() <= 1e-5 * np.abs(ref).max()
# every rank applies the same reduced gradient: replicas stay identical
new, _ = og.apply_gradients(params, P.unflatten(blob.numpy()), og.new_optimizer_state(params))
mine = torch.from_numpy(P.flatten({{k: v.astype(np.float32) for k, v in new.items()}}))
other = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(other, mine)
assert all(torch.equal(o, mine) for o in other)
if rank == 0:
    print("GLOO_GRAD_OK", world)
dist.destroy_process_group()
"""


def test_world_size_2_gloo_gradient_allreduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_GRAD_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29733", str(script)],
                         capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "GLOO_GRAD_OK 2" in res.stdout


def test_binary_search_bounds_follow_the_reference():
    """experiments/binary_search.py:23-34 on a 4-city instance (values worked out by hand)."""
    from tsp_gnn_b200 import experiments
    Mw = np.array([[0, 1, 2, 3], [1, 0, 4, 5], [2, 4, 0, 6], [3, 5, 6, 0]], dtype=float)
    lo, hi = experiments.cost_bounds(Mw)
    # triu/tril each hold 10 zeros and {1,2,3,4,5,6}: the 4 lightest entries are zeros, the 4 heaviest 3+4+5+6
    assert lo == 0.0 and hi == 18.0 / 4


def test_library_exports_every_declared_symbol():
    from tsp_gnn_b200 import _lib
    header = open(os.path.join(ROOT, "include", "tspgnn.h")).read()
    declared = set(re.findall(r"\b(tspgnn_[a-z_0-9]+)\s*\(", header))
    declared.discard("tspgnn_ctx")
    assert len(declared) >= 15
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == {n for n, _, _ in _lib.SIGNATURES}
    assert lib.tspgnn_version() == 100
    _lib.lib.tspgnn_param_count.restype = ctypes.c_int64
    assert _lib.lib.tspgnn_param_count(64) == 115529


def test_dense_ev_helper_runs_on_host():
    from tsp_gnn_b200.engine import dense_ev_to_coo
    from tsp_gnn_b200._lib import TspGnnError
    EV = inst.synth_batch([6, 5], seed=2)[0]
    for dt in (np.float64, np.float32):
        s, d = dense_ev_to_coo(EV.toarray(dt))
        assert np.array_equal(s, EV.src) and np.array_equal(d, EV.dst)
    bad = EV.toarray(); bad[3, :] = 1
    with pytest.raises(TspGnnError, match="exactly 2"):
        dense_ev_to_coo(bad)
