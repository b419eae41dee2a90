"""Golden vectors for the GENERIC GraphNN features the TSP model does not use (graphnn.py:142-173,90,
244-255): a transfer function ``fun``, a matrix-only input with an integer second dimension, several
concatenated update terms per variable, a transposed matrix.  The reference's own graphnn.py / mlp.py run
unmodified on the numpy TF1 stand-in (oracle/tf1_shim.py); only possible where /root/reference exists.

    python tests/golden/make_reference_generic_golden.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path.insert(0, ROOT)
from oracle import tf1_shim      # noqa: E402

VAR = {"A": 8, "B": 16, "C": 8}
MAT = {"M_AB": ("A", "B"), "M_CB": ("C", "B"), "F_C": ("C", 4)}
MSG = {"B2A": ("B", "A"), "A2B": ("A", "B"), "B2C": ("B", "C")}
N = {"A": 7, "B": 5, "C": 6}
T_STEPS = 3


def toy_loop(square):
    """The update rules; ``square`` is the transfer function in the caller's tensor library."""
    return {
        "A": [{"mat": "M_AB", "msg": "B2A", "var": "B"}, {"var": "A", "fun": square}],
        "B": [{"mat": "M_AB", "transpose?": True, "msg": "A2B", "var": "A"}],
        "C": [{"mat": "M_CB", "msg": "B2C", "var": "B"}, {"mat": "F_C"}],
    }


def toy_inputs(seed=3):
    rng = np.random.RandomState(seed)
    mats = {"M_AB": (rng.rand(N["A"], N["B"]) < 0.5).astype(np.float64),
            "M_CB": (rng.rand(N["C"], N["B"]) < 0.4).astype(np.float64),
            "F_C": rng.normal(size=(N["C"], 4))}
    init = {v: np.abs(rng.normal(size=(N[v], d))) for v, d in VAR.items()}
    c0 = {"B": rng.normal(size=(N["B"], VAR["B"]))}          # LSTM_initial_states for one variable only
    return mats, init, c0


if __name__ == "__main__":
    tf1_shim.install()
    import tensorflow as tf      # the shim
    sys.path.insert(0, REF)
    for m in ("mlp", "graphnn"):
        sys.modules.pop(m, None)
    graphnn = importlib.import_module("graphnn")
    assert os.path.dirname(graphnn.__file__) == REF
    tf1_shim.reset(dtype=np.float64, seed=5)
    gnn = graphnn.GraphNN(VAR, MAT, MSG, toy_loop(lambda y: tf.multiply(y, y)), name="TOY")
    mats, init, c0 = toy_inputs()
    ph_m = {k: tf.placeholder(tf.float32, shape=(None, None), name=k) for k in mats}
    ph_i = {k: tf.placeholder(tf.float32, shape=(None, None), name="init_" + k) for k in init}
    ph_c = {k: tf.placeholder(tf.float32, shape=(None, None), name="c0_" + k) for k in c0}
    ts = tf.placeholder(tf.int32, shape=(), name="time_steps")
    last = gnn(ph_m, ph_i, ts, LSTM_initial_states=ph_c)
    feed = {ts: T_STEPS}
    feed.update({ph_m[k]: v for k, v in mats.items()})
    feed.update({ph_i[k]: v for k, v in init.items()})
    feed.update({ph_c[k]: v for k, v in c0.items()})
    sess = tf1_shim.Session()
    st = sess.run(last, feed_dict=feed)
    store = {}
    for v in VAR:
        store["out/%s/h" % v] = np.asarray(st[v].h)
        store["out/%s/c" % v] = np.asarray(st[v].c)
    for k, val in tf1_shim.get_variables().items():
        store["var/" + k.replace("/", "|")] = np.asarray(val)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_shim_generic.npz")
    np.savez_compressed(path, **store)
    print("wrote", path, sorted(k for k in store if k.startswith("var/"))[:6], "...", len(store), "arrays")
    print({v: float(np.abs(store["out/%s/h" % v]).mean()) for v in VAR})
