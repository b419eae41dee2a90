"""The oracle against the reference's OWN graph code executed on the TF1 stand-in (oracle/tf1_shim.py:
numpy for the forward pass, its float64 torch backend for the training graph; fixtures made by
tests/golden/make_reference_golden.py).  CPU only.

The fixtures are outputs of /root/reference/model.py::build_network + graphnn.py + mlp.py +
instance_loader.py::create_batch, unmodified; see tf1_shim.py for what that pins and what it does not."""
import os
import sys

import numpy as np
import pytest

from oracle import tspgnn_oracle as orc
from tsp_gnn_b200 import instances as inst
from tsp_gnn_b200 import params as P

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import make_reference_golden as mrg  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_shim_forward.npz"))


def _inputs(name):
    sizes, iseed, pseed, T, conn = mrg.CASES[name]
    EV, W, C, y, nv, ne = inst.synth_batch(sizes, seed=iseed, connectivity=conn)
    params = orc.init_params(64, seed=pseed, perturb_ln=True)
    return EV, W, C, y, nv, ne, params, T


@pytest.mark.parametrize("name", sorted(mrg.CASES))
def test_oracle_matches_reference_graph_code(name):
    EV, W, C, y, nv, ne, params, T = _inputs(name)
    # the product's vectorised batch builder against the reference's loops (instance_loader.py:29-80)
    assert np.array_equal(W.reshape(-1), GOLD[name + "/W"]) and np.allclose(C.reshape(-1), GOLD[name + "/C"], rtol=0, atol=1e-15)
    out = orc.forward(params, EV.src, EV.dst, W, C, nv, ne, T, dtype=np.float64)
    assert np.abs(out["predictions"] - GOLD[name + "/predictions"]).max() < 1e-12
    assert np.abs(out["V_h"] - GOLD[name + "/V_h"]).max() < 1e-10
    assert np.abs(out["V_c"] - GOLD[name + "/V_c"]).max() < 1e-10
    assert np.abs(out["E_h"][:48] - GOLD[name + "/E_h_head"]).max() < 1e-10
    assert np.abs(out["E_c"][:48] - GOLD[name + "/E_c_head"]).max() < 1e-10
    assert np.abs(out["E_h"].sum(axis=1) - GOLD[name + "/E_h_rowsum"]).max() < 1e-9
    m = orc.metrics(out["logits"], y)
    assert abs(m["loss"] - GOLD[name + "/loss"]) < 1e-12 and m["acc"] == GOLD[name + "/acc"]
    assert [m["TP"], m["FP"], m["TN"], m["FN"]] == list(GOLD[name + "/confusion"])


TRAIN = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_shim_train.npz"))


def train_inputs(name):
    sizes, iseed, pseed, T, conn, n_steps = mrg.TRAIN_CASES[name]
    EV, W, C, y, nv, ne = inst.synth_batch(sizes, seed=iseed, connectivity=conn)
    params = orc.init_params(64, seed=pseed, perturb_ln=True)
    return EV, W, C, y, nv, ne, params, T, n_steps


def train_fixture(name, key):
    """{variable name: array} of one group of the training fixture."""
    prefix = "%s/%s|" % (name, key)
    return {k[len(prefix):].replace("|", "/"): TRAIN[k] for k in TRAIN.files if k.startswith(prefix)}


@pytest.mark.parametrize("name", sorted(mrg.TRAIN_CASES))
def test_training_oracle_matches_reference_training_graph(name):
    """model.py:157-167 as the reference builds it (loss + 1e-10 * sum l2_loss over tf.trainable_variables(),
    tf.gradients, clip_by_global_norm 0.65, AdamOptimizer 2e-5) executed on the shim's torch backend, against
    the hand-derived oracle: gradients incl. the L2 term, global norm, loss, and the variables after each step."""
    from oracle import tspgnn_oracle_grad as og
    EV, W, C, y, nv, ne, params, T, n_steps = train_inputs(name)
    assert np.allclose(W.reshape(-1), TRAIN[name + "/W"], rtol=0, atol=1e-15)
    cur = {k: v.astype(np.float64) for k, v in params.items()}
    st = og.new_optimizer_state(cur)
    for step in range(n_steps):
        ref = og.forward_backward(cur, EV.src, EV.dst, W, C, nv, ne, y, T)
        assert abs(ref["loss"] - TRAIN["%s/loss_%d" % (name, step)]) < 1e-12
        assert np.abs(ref["predictions"] - TRAIN["%s/predictions_%d" % (name, step)]).max() < 1e-12
        if step == 0:
            gold = train_fixture(name, "grad_0")
            assert sorted(gold) == sorted(cur)                      # every trainable variable gets a gradient
            for k, g in gold.items():
                mine = ref["grads"][k] + og.L2NORM_SCALING * cur[k]   # tf.gradients(loss + l2 * vars_cost)
                assert np.abs(mine - g).max() <= 2e-7 * np.abs(g).max() + 1e-12, k   # fixture stored as float32
        cur, gnorm = og.apply_gradients(cur, ref["grads"], st)
        assert abs(gnorm - TRAIN["%s/global_norm_%d" % (name, step)]) < 1e-10
        gold = train_fixture(name, "dvar_over_lr_%d" % step)
        for k, dv in gold.items():
            mine = (cur[k] - params[k].astype(np.float64)) / og.LEARNING_RATE
            assert np.abs(mine - dv.astype(np.float64)).max() < 2e-3 * (step + 1), k   # float16 storage, units of lr
    assert st["step"] == int(TRAIN[name + "/adam_step"])


def test_variable_names_are_the_ones_the_reference_scopes_produce():
    assert sorted(GOLD["variable_names"]) == sorted(n for n, _, _ in P.param_spec(64))


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference sources only exist in the build container")
def test_fixtures_regenerate_from_the_reference_sources():
    model, loader = mrg.load_reference_modules()
    tr = mrg.run_train_case("ref_train_sparse", model, loader)
    assert tr["loss_0"] == TRAIN["ref_train_sparse/loss_0"] and tr["global_norm_0"] == TRAIN["ref_train_sparse/global_norm_0"]
    k = "grad_0/TSP/E_cell/layer_norm_basic_lstm_cell/kernel"
    assert np.array_equal(tr[k], TRAIN["ref_train_sparse/" + k.replace("/", "|")])
    from oracle import tf1_shim as _shim
    _shim.reset()
    out = mrg.run_case("ref_tiny", model, loader)
    assert np.array_equal(out["predictions"], GOLD["ref_tiny/predictions"])
    assert np.array_equal(out["V_c"], GOLD["ref_tiny/V_c"])
    # the shim raises the reference's own run-time assertion (graphnn.py:234-243) on a bad matrix
    from oracle import tf1_shim
    tf1_shim.reset()
    GNN = model.build_network(64)
    with pytest.raises(tf1_shim.errors.InvalidArgumentError, match="Matrix EV"):
        tf1_shim.Session().run(GNN["predictions"], feed_dict={
            GNN["EV"]: np.zeros((3, 4)), GNN["W"]: np.zeros((2, 1)), GNN["C"]: np.zeros((2, 1)), GNN["time_steps"]: 1,
            GNN["route_exists"]: np.zeros(1), GNN["n_vertices"]: np.array([4]), GNN["n_edges"]: np.array([2])})
