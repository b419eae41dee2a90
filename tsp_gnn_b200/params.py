"""Trainable-variable inventory of ``build_network(d)`` keyed by the TF variable names.

Names follow the scopes the reference creates (mlp.py:36-38, graphnn.py:65-66,129,168,
model.py:37,47,93,111); see SURVEY.md 8f-3.  The flat float32 blob in this order is what
``tspgnn_set_params`` (include/tspgnn.h) consumes.
"""
import os
import numpy as np

GATE_SCOPES = ("input", "transform", "forget", "output", "state")


def param_spec(d=64):
    """[(name, shape, init_kind)] in canonical blob order."""
    spec = []
    sizes = [2, int(d / 8), int(d / 4), int(d / 2), d]       # model.py:34 (float sizes cast by Dense)
    for i in range(4):
        spec.append(("E_init_MLP_MLP_layer_%d/kernel" % (i + 1), (sizes[i], sizes[i + 1]), "xavier"))
        spec.append(("E_init_MLP_MLP_layer_%d/bias" % (i + 1), (sizes[i + 1],), "zeros"))   # model.py:40
    spec.append(("V_init", (1, d), "normal"))                                               # model.py:47
    for msg in ("V_msg_E", "E_msg_V"):
        for i in range(4):
            spec.append(("TSP/%s_MLP_layer_%d/kernel" % (msg, i + 1), (d, d), "xavier"))
            # graphnn.py:121 passes the *weight* initializer as bias_initializer
            spec.append(("TSP/%s_MLP_layer_%d/bias" % (msg, i + 1), (d,), "xavier_bias"))
    for v in ("V", "E"):
        base = "TSP/%s_cell/layer_norm_basic_lstm_cell" % v
        spec.append((base + "/kernel", (2 * d, 4 * d), "xavier"))
        for g in GATE_SCOPES:
            spec.append(("%s/%s/gamma" % (base, g), (d,), "ones"))
            spec.append(("%s/%s/beta" % (base, g), (d,), "zeros"))
    vs = [d, d, d, d, 1]
    for i in range(4):
        spec.append(("E_vote_MLP_layer_%d/kernel" % (i + 1), (vs[i], vs[i + 1]), "xavier"))
        spec.append(("E_vote_MLP_layer_%d/bias" % (i + 1), (vs[i + 1],), "zeros"))         # model.py:114
    return spec


def param_offsets(d=64):
    """name -> (offset, size, shape) into the flat blob; plus total length."""
    off, table = 0, {}
    for name, shape, _ in param_spec(d):
        n = int(np.prod(shape))
        table[name] = (off, n, shape)
        off += n
    return table, off


def init_params(d=64, seed=None):
    """Reference initialisers: xavier-uniform kernels (and message-MLP biases), zero
    E_init/E_vote biases, glorot-uniform LSTM kernels, gamma=1, beta=0, V_init~N(0,1)."""
    rng = np.random.RandomState(seed)
    out = {}
    for name, shape, kind in param_spec(d):
        if kind == "xavier":
            lim = np.sqrt(6.0 / (shape[0] + shape[1]))
            a = rng.uniform(-lim, lim, size=shape)
        elif kind == "xavier_bias":
            lim = np.sqrt(6.0 / (2 * shape[0]))
            a = rng.uniform(-lim, lim, size=shape)
        elif kind == "normal":
            a = rng.normal(size=shape)
        elif kind == "ones":
            a = np.ones(shape)
        else:
            a = np.zeros(shape)
        out[name] = a.astype(np.float32)
    return out


def flatten(params, d=64):
    table, total = param_offsets(d)
    blob = np.empty(total, dtype=np.float32)
    for name, (off, n, shape) in table.items():
        if name not in params:
            raise KeyError("missing variable %r" % name)
        a = np.asarray(params[name], dtype=np.float32)
        if tuple(a.shape) != tuple(shape):
            raise ValueError("variable %r has shape %r, expected %r" % (name, a.shape, shape))
        blob[off:off + n] = a.reshape(-1)
    return blob


def unflatten(blob, d=64):
    table, total = param_offsets(d)
    blob = np.asarray(blob, dtype=np.float32)
    if blob.shape != (total,):
        raise ValueError("blob has %r elements, expected %d" % (blob.shape, total))
    return {name: blob[off:off + n].reshape(shape).copy() for name, (off, n, shape) in table.items()}


def save_weights(params, path):
    """util.py:24-37 analogue: ``<path>/model.npz`` keyed by TF variable names."""
    os.makedirs(path, exist_ok=True)
    np.savez(os.path.join(path, "model.npz"), **{k.replace("/", "|"): v for k, v in params.items()})


def load_weights(path):
    """util.py:5-22 analogue; raises like the reference when the path is missing."""
    if not os.path.exists(path):
        raise Exception("Path does not exist!")
    with np.load(os.path.join(path, "model.npz")) as z:
        return {k.replace("|", "/"): z[k] for k in z.files}
