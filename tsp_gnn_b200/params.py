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


# tf.train.Saver() without a var_list stores every GLOBAL variable (util.py:13,31): the trainable ones plus
# AdamOptimizer's slots "<var>/Adam" (m), "<var>/Adam_1" (v) and the scalars beta1_power / beta2_power.
ADAM_M, ADAM_V = "/Adam", "/Adam_1"


def optimizer_state_to_named(state, d=64, beta1=0.9, beta2=0.999):
    """{'m','v','step'} (flat blobs, Engine.get_optimizer_state) -> TF-named slot arrays."""
    out = {}
    for name, a in unflatten(state["m"], d).items():
        out[name + ADAM_M] = a
    for name, a in unflatten(state["v"], d).items():
        out[name + ADAM_V] = a
    out["beta1_power"] = np.float32(beta1 ** (int(state["step"]) + 1))     # TF keeps beta^(t+1) after t updates
    out["beta2_power"] = np.float32(beta2 ** (int(state["step"]) + 1))
    out["_adam_step"] = np.int64(state["step"])
    return out


def named_to_optimizer_state(named, d=64):
    """Inverse of optimizer_state_to_named; None when the checkpoint holds no Adam slots."""
    table, _ = param_offsets(d)
    if not all((n + ADAM_M) in named and (n + ADAM_V) in named for n in table):
        return None
    m = flatten({n: named[n + ADAM_M] for n in table}, d)
    v = flatten({n: named[n + ADAM_V] for n in table}, d)
    if "_adam_step" in named:
        step = int(named["_adam_step"])
    elif "beta1_power" in named:                          # a converted TF checkpoint: beta1_power = 0.9^(t+1)
        step = max(0, int(round(np.log(float(named["beta1_power"])) / np.log(0.9))) - 1)
    else:
        step = 0
    return {"m": m, "v": v, "step": step}


def save_weights(params, path, optimizer_state=None):
    """util.py:24-37 analogue: ``<path>/model.npz`` keyed by TF variable names; with ``optimizer_state``
    also the Adam slots and beta powers, like tf.train.Saver() (all global variables)."""
    os.makedirs(path, exist_ok=True)
    store = dict(params)
    if optimizer_state is not None:
        d = int(np.asarray(params["V_init"]).shape[-1])
        store.update(optimizer_state_to_named(optimizer_state, d))
    np.savez(os.path.join(path, "model.npz"), **{k.replace("/", "|"): v for k, v in store.items()})


def load_checkpoint(path):
    """Every array of the checkpoint, TF names -> numpy (variables, Adam slots, beta powers)."""
    if not os.path.exists(path):
        raise Exception("Path does not exist!")
    with np.load(os.path.join(path, "model.npz")) as z:
        return {k.replace("|", "/"): z[k] for k in z.files}


def load_weights(path):
    """util.py:5-22 analogue; raises like the reference when the path is missing.  Returns the trainable
    variables only (see load_checkpoint / named_to_optimizer_state for the optimizer slots)."""
    allv = load_checkpoint(path)
    return {k: v for k, v in allv.items()
            if not (k.endswith(ADAM_M) or k.endswith(ADAM_V) or k in ("beta1_power", "beta2_power", "_adam_step"))}
