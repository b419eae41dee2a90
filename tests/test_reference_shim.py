"""The oracle against the reference's OWN graph code executed on the numpy TF1 stand-in
(oracle/tf1_shim.py; fixtures made by tests/golden/make_reference_golden.py).  CPU only.

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


def test_variable_names_are_the_ones_the_reference_scopes_produce():
    assert sorted(GOLD["variable_names"]) == sorted(n for n, _, _ in P.param_spec(64))


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference sources only exist in the build container")
def test_fixtures_regenerate_from_the_reference_sources():
    model, loader = mrg.load_reference_modules()
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
