"""Runs the reference's OWN graph code (/root/reference/model.py, graphnn.py, mlp.py, instance_loader.py,
unmodified, imported in place) on the numpy TF1 stand-in oracle/tf1_shim.py and stores its outputs as
golden vectors.  Only possible where /root/reference exists (the build container); the fixtures travel.

    python tests/golden/make_reference_golden.py

What this pins / does not pin is stated in oracle/tf1_shim.py: the reference's Python (wiring, op order,
batch layout, read-out, metrics, variable names) is executed for real; TensorFlow's internals
(LayerNormBasicLSTMCell, layer_norm, Dense) are the shim's restatement.
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path.insert(0, ROOT)
from oracle import tf1_shim, tspgnn_oracle as orc      # noqa: E402
from tsp_gnn_b200 import instances as inst             # noqa: E402

CASES = {
    # name: (sizes, instance seed, param seed, time_steps, connectivity)
    "ref_tiny": ([5, 6, 7, 8], 11, 7, 4, 1.0),
    "ref_sparse": ([9, 12, 10], 5, 3, 6, 0.5),
    "ref_config1": ([20] * 16, 42, 0, 32, 1.0),
}


def load_reference_modules():
    tf1_shim.install()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    for m in ("mlp", "graphnn", "model", "instance_loader"):
        sys.modules.pop(m, None)
    model = importlib.import_module("model")
    loader = importlib.import_module("instance_loader")
    assert os.path.dirname(model.__file__) == REF, model.__file__
    return model, loader


def run_case(name, model, loader):
    sizes, iseed, pseed, T, conn = CASES[name]
    instances = inst.synth_instances(sizes, seed=iseed, connectivity=conn)
    # the reference's own batch builder: dense EV (instance_loader.py:29-80)
    EV, W, C, route_exists, n_vertices, n_edges = loader.InstanceLoader.create_batch(instances, dev=0.02)
    params = orc.init_params(64, seed=pseed, perturb_ln=True)
    tf1_shim.reset(dtype=np.float64, seed=0)
    GNN = model.build_network(64)
    sess = tf1_shim.Session()
    feed = {GNN["EV"]: EV, GNN["W"]: W, GNN["C"]: C, GNN["time_steps"]: T, GNN["route_exists"]: route_exists,
            GNN["n_vertices"]: n_vertices, GNN["n_edges"]: n_edges}
    sess.run(GNN["predictions"], feed_dict=feed)            # first run creates every variable (random init)
    names = tf1_shim.variable_names()
    assert sorted(names) == sorted(params), (sorted(set(names) ^ set(params)))
    tf1_shim.set_variables(params)                          # same seeded parameters as the oracle, by TF name
    fetch = [GNN[k] for k in ("predictions", "loss", "acc", "TP", "FP", "TN", "FN")]
    preds, loss, acc, TP, FP, TN, FN = sess.run(fetch, feed_dict=feed)
    st = sess.run(GNN["last_states"], feed_dict=feed)
    return dict(predictions=preds, loss=loss, acc=acc, confusion=np.array([TP, FP, TN, FN]),
                E_h=st["E"].h, E_c=st["E"].c, V_h=st["V"].h, V_c=st["V"].c,
                W=W.reshape(-1), C=C.reshape(-1), n_edges=n_edges, names=np.array(sorted(names)))


TRAIN_CASES = {
    # name: (sizes, instance seed, param seed, time_steps, connectivity, optimizer steps)
    "ref_train_tiny": ([6, 9, 7, 8], 13, 9, 3, 1.0, 2),
    "ref_train_sparse": ([10, 12, 9, 11], 4, 2, 5, 0.6, 1),
}


def run_train_case(name, model, loader):
    """The reference's training graph (model.py:157-167: loss + 1e-10 * sum l2_loss, tf.gradients,
    clip_by_global_norm(0.65), AdamOptimizer(2e-5).apply_gradients) executed on the torch backend of the
    shim: ``sess.run([train_step, loss, predictions])`` exactly as train.py:36-42 issues it."""
    sizes, iseed, pseed, T, conn, n_steps = TRAIN_CASES[name]
    instances = inst.synth_instances(sizes, seed=iseed, connectivity=conn)
    EV, W, C, route_exists, n_vertices, n_edges = loader.InstanceLoader.create_batch(instances, dev=0.02)
    params = orc.init_params(64, seed=pseed, perturb_ln=True)
    tf1_shim.reset(dtype=np.float64, seed=0, backend="torch")
    GNN = model.build_network(64)
    sess = tf1_shim.Session()
    feed = {GNN["EV"]: EV, GNN["W"]: W, GNN["C"]: C, GNN["time_steps"]: T, GNN["route_exists"]: route_exists,
            GNN["n_vertices"]: n_vertices, GNN["n_edges"]: n_edges}
    sess.run(GNN["predictions"], feed_dict=feed)            # creates every variable
    tf1_shim.set_variables(params)
    out = dict(W=W.reshape(-1), C=C.reshape(-1), route_exists=np.asarray(route_exists, dtype=np.float64))
    for step in range(n_steps):
        _, loss, preds = sess.run([GNN["train_step"], GNN["loss"], GNN["predictions"]], feed_dict=feed)
        out["loss_%d" % step] = np.float64(loss)
        out["predictions_%d" % step] = preds
        out["global_norm_%d" % step] = np.float64(tf1_shim.last_global_norm())
        # Compact storage: gradients of the first step as float32; the variables as their change since the
        # initial values in units of the learning rate (Adam moves a variable by at most ~lr per step) as
        # float16, i.e. to 1e-3 of a step (an update of 2e-5 would vanish in the rounding of the variable itself)
        if step == 0:
            for k, v in tf1_shim.last_gradients().items():
                out["grad_0/%s" % k] = v.astype(np.float32)               # d(loss + l2)/d var, before the clip
        for k, v in tf1_shim.get_variables().items():
            out["dvar_over_lr_%d/%s" % (step, k)] = ((v - params[k]) / 2e-5).astype(np.float16)
    out["adam_step"] = np.int64(tf1_shim.optimizer_slots()["step"])
    return out


if __name__ == "__main__":
    model, loader = load_reference_modules()
    tstore = {}
    for name in TRAIN_CASES:
        out = run_train_case(name, model, loader)
        for k, v in out.items():
            tstore[name + "/" + k.replace("/", "|")] = v
        print(name, "loss", [float(out["loss_%d" % i]) for i in range(TRAIN_CASES[name][5])],
              "global norm %.6f" % out["global_norm_0"])
    tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_shim_train.npz")
    np.savez_compressed(tpath, **tstore)
    print("wrote", tpath)
    tf1_shim.reset()
    store = {}
    for name in CASES:
        out = run_case(name, model, loader)
        for k in ("predictions", "loss", "acc", "confusion", "V_h", "V_c", "W", "C"):
            store[name + "/" + k] = out[k]
        store[name + "/E_h_head"] = out["E_h"][:48]
        store[name + "/E_c_head"] = out["E_c"][:48]
        store[name + "/E_h_rowsum"] = out["E_h"].sum(axis=1)
        store["variable_names"] = out["names"]
        print(name, "predictions", out["predictions"][:4], "loss %.6f" % out["loss"])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_shim_forward.npz")
    np.savez_compressed(path, **store)
    print("wrote", path)
