"""CPU restatement of the TSP-GNN forward pass.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module; the product
package (``tsp_gnn_b200``) never does.

PARITY UNPINNED: the reference (machine-reasoning-ufrgs/TSP-GNN) ships no tests,
golden vectors or checkpoints, and its arithmetic lives in TensorFlow 1.x
(``tf.contrib.rnn.LayerNormBasicLSTMCell``, ``tf.contrib.layers.layer_norm``,
``tf.layers.Dense``), which is neither vendored under /root/reference nor
installable here (version unpinned; ``tf.contrib`` => TF <= 1.15).  The TF-internal
semantics below (marked "TF:") are restated from the published TF 1.x source of those
ops; everything else follows the reference file:line cited next to it.

Three formulations of the same forward pass:
  * ``forward(..., dtype=np.float64)``            ground truth for error reports
  * ``forward(..., dtype=np.float32, dense=True)``  multiplies the dense block-diagonal
        EV matrix exactly like graphnn.py:156-160 (the "reference CPU path" stand-in)
  * ``forward(..., dtype=np.float32, dense=False)`` gather / segment-sum form
"""
import numpy as np

D_DEFAULT = 64
LN_EPS = 1e-12       # TF: tf.contrib.layers.layer_norm -> variance_epsilon = 1e-12
FORGET_BIAS = 1.0    # TF: LayerNormBasicLSTMCell(forget_bias=1.0) default (graphnn.py:107-112)
GATE_SCOPES = ("input", "transform", "forget", "output", "state")


# ----------------------------------------------------------------------------
# parameter inventory (names follow the TF variable scopes the reference creates)
# ----------------------------------------------------------------------------
def param_spec(d=D_DEFAULT):
    """[(tf_variable_name, shape, initializer)] in canonical order.

    mlp.py:36-38   layer names  <name>_MLP_layer_<i>
    model.py:33-41 E_init_MLP   sizes d/8,d/4,d/2 -> d, zero biases
    model.py:47    V_init       random_normal (1,d)
    graphnn.py:114-125  message MLPs: 3 hidden + output, all d wide; biases use the
                        *weight* initializer (xavier), graphnn.py:121
    graphnn.py:107-112  LayerNormBasicLSTMCell(d, activation=relu)
    model.py:107-115    E_vote  d,d,d -> 1, zero biases
    """
    spec = []
    sizes = [2, int(d / 8), int(d / 4), int(d / 2), d]
    for i in range(4):
        spec.append(("E_init_MLP_MLP_layer_%d/kernel" % (i + 1), (sizes[i], sizes[i + 1]), "xavier"))
        spec.append(("E_init_MLP_MLP_layer_%d/bias" % (i + 1), (sizes[i + 1],), "zeros"))
    spec.append(("V_init", (1, d), "normal"))
    for msg in ("V_msg_E", "E_msg_V"):
        for i in range(4):
            spec.append(("TSP/%s_MLP_layer_%d/kernel" % (msg, i + 1), (d, d), "xavier"))
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
        spec.append(("E_vote_MLP_layer_%d/bias" % (i + 1), (vs[i + 1],), "zeros"))
    return spec


def init_params(d=D_DEFAULT, seed=0, perturb_ln=False):
    """Seeded parameters with the reference initialisers.

    TF: xavier_initializer() (uniform) = U(+-sqrt(6/(fan_in+fan_out))); for a 1-D bias
    of length n TF's fan computation gives fan_in = fan_out = n.  LSTM kernel uses the
    variable-scope default glorot_uniform.  gamma=1, beta=0.  V_init ~ N(0,1).
    ``perturb_ln`` moves gamma/beta/zero-biases off their trivial initial values so
    that tests exercise them (a trained model has non-trivial values there).
    """
    rng = np.random.RandomState(seed)
    out = {}
    for name, shape, kind in param_spec(d):
        if kind == "xavier":
            lim = np.sqrt(6.0 / (shape[0] + shape[1]))
            a = rng.uniform(-lim, lim, size=shape)
        elif kind == "xavier_bias":
            lim = np.sqrt(6.0 / (shape[0] + shape[0]))
            a = rng.uniform(-lim, lim, size=shape)
        elif kind == "normal":
            a = rng.normal(size=shape)
        elif kind == "ones":
            a = np.ones(shape)
            if perturb_ln:
                a = a + 0.1 * rng.normal(size=shape)
        elif kind == "zeros":
            a = np.zeros(shape)
            if perturb_ln:
                a = a + 0.05 * rng.normal(size=shape)
        else:
            raise ValueError(kind)
        out[name] = a.astype(np.float32)
    return out


# ----------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------
def spread_params(params, kernel_scale, logit_shift):
    """Test parameter set whose predictions spread over (0, 1): with the reference initialisers every
    prediction of a batch sits within ~1e-4 of the others, so a kernel that ignored most of the graph would
    still pass a 1e-4 gate.  Scaling every kernel makes the network less contractive (instances end in
    visibly different states) and the shift of the last vote bias centres the logits around zero."""
    out = {}
    for k, v in params.items():
        out[k] = (v * np.float32(kernel_scale)).astype(np.float32) if k.endswith("kernel") else v.copy()
    out["E_vote_MLP_layer_4/bias"] = (out["E_vote_MLP_layer_4/bias"] - np.float32(logit_shift)).astype(np.float32)
    return out


def sigmoid(x):
    """Stable logistic; same value as 1/(1+exp(-x)) wherever that does not overflow."""
    e = np.exp(-np.abs(x))
    return np.where(x >= 0, 1.0 / (1.0 + e), e / (1.0 + e)).astype(x.dtype)


def relu(x):
    return np.maximum(x, 0)


def dense(x, kernel, bias):
    """TF: tf.layers.Dense -> x @ kernel[in,out] + bias   (mlp.py:39-52)."""
    return x @ kernel + bias


def mlp(x, params, prefix, n_layers=4):
    """mlp.py:57-63 with ReLU on all but the last layer (graphnn.py:116-119, model.py:33-41,107-115)."""
    for i in range(n_layers):
        x = dense(x, params["%s_MLP_layer_%d/kernel" % (prefix, i + 1)],
                  params["%s_MLP_layer_%d/bias" % (prefix, i + 1)])
        if i < n_layers - 1:
            x = relu(x)
    return x


def layer_norm(u, gamma, beta):
    """TF: layers.layer_norm(begin_norm_axis=1): nn.moments (two-pass, biased variance)
    then nn.batch_normalization: inv = rsqrt(var+eps)*gamma; u*inv + (beta - mean*inv)."""
    mean = u.mean(axis=1, keepdims=True)
    var = np.square(u - mean).mean(axis=1, keepdims=True)
    inv = (1.0 / np.sqrt(var + u.dtype.type(LN_EPS))) * gamma
    return u * inv + (beta - mean * inv)


def lnlstm(x, c, h, params, base):
    """TF: LayerNormBasicLSTMCell.call with activation=relu (graphnn.py:15,110),
    layer_norm=True (no bias), dropout_keep_prob=1.  Returns (c', h')."""
    d = h.shape[1]
    z = np.concatenate([x, h], axis=1) @ params[base + "/kernel"]      # graphnn.py:167-169
    i, j, f, o = z[:, :d], z[:, d:2 * d], z[:, 2 * d:3 * d], z[:, 3 * d:]
    i = layer_norm(i, params[base + "/input/gamma"], params[base + "/input/beta"])
    j = layer_norm(j, params[base + "/transform/gamma"], params[base + "/transform/beta"])
    f = layer_norm(f, params[base + "/forget/gamma"], params[base + "/forget/beta"])
    o = layer_norm(o, params[base + "/output/gamma"], params[base + "/output/beta"])
    g = relu(j)
    new_c = c * sigmoid(f + c.dtype.type(FORGET_BIAS)) + sigmoid(i) * g
    new_c = layer_norm(new_c, params[base + "/state/gamma"], params[base + "/state/beta"])
    new_h = relu(new_c) * sigmoid(o)
    return new_c, new_h


def dense_EV(src, dst, n_vertices_total, dtype):
    """instance_loader.py:45,63-66: EV[e, src(e)] = EV[e, dst(e)] = 1 (global ids)."""
    E = len(src)
    EV = np.zeros((E, n_vertices_total), dtype=dtype)
    EV[np.arange(E), src] = 1
    EV[np.arange(E), dst] = 1
    return EV


def message_passing(P, src, dst, EV, E_c, E_h, V_c, V_h, time_steps, trace=None):
    """The hot loop: ``time_steps`` iterations of while_body (graphnn.py:142-173, driven by
    tf.while_loop graphnn.py:175-179).  ``EV`` dense [sumE,sumV] multiplies exactly like
    graphnn.py:156-160; ``EV=None`` uses the gather / segment-sum form of the same products.
    ``P`` must already be cast to the working dtype.  Returns (E_c, E_h, V_c, V_h)."""
    dtype = E_h.dtype
    nV, d = V_h.shape
    for _ in range(int(time_steps)):
        mE = mlp(E_h, P, "TSP/E_msg_V")                     # graphnn.py:152-154
        mV = mlp(V_h, P, "TSP/V_msg_E")
        if EV is not None:
            xV = EV.T @ mE                                   # graphnn.py:156-160, adjoint_a=True
            xE = EV @ mV
        else:
            xV = np.zeros((nV, d), dtype=dtype)
            np.add.at(xV, src, mE)
            np.add.at(xV, dst, mE)
            xE = mV[src] + mV[dst]
        # both cells read the time-t states (graphnn.py:144-148)
        nVc, nVh = lnlstm(xV, V_c, V_h, P, "TSP/V_cell/layer_norm_basic_lstm_cell")
        nEc, nEh = lnlstm(xE, E_c, E_h, P, "TSP/E_cell/layer_norm_basic_lstm_cell")
        V_c, V_h, E_c, E_h = nVc, nVh, nEc, nEh
        if trace is not None:
            trace.append(dict(mE=mE, mV=mV, xV=xV, xE=xE, V_c=V_c, V_h=V_h, E_c=E_c, E_h=E_h))
    return E_c, E_h, V_c, V_h


# ----------------------------------------------------------------------------
# forward pass
# ----------------------------------------------------------------------------
def forward(params, src, dst, W, C, n_vertices, n_edges, time_steps,
            dtype=np.float64, dense=False, return_trace=False):
    """model.py:33-51,118-147 + graphnn.py:134-179.

    src,dst : int [sumE] global vertex ids of each edge row (src<dst)
    W, C    : float [sumE] or [sumE,1]
    Returns dict(logits, predictions, E_vote, E_h, E_c, V_h, V_c[, trace]).
    """
    P = {k: v.astype(dtype) for k, v in params.items()}
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    n_vertices = np.asarray(n_vertices, dtype=np.int64)
    n_edges = np.asarray(n_edges, dtype=np.int64)
    nV, nE = int(n_vertices.sum()), int(n_edges.sum())
    assert len(src) == nE and len(dst) == nE
    d = P["V_init"].shape[1]
    W = np.asarray(W, dtype=dtype).reshape(nE, 1)
    C = np.asarray(C, dtype=dtype).reshape(nE, 1)

    # model.py:33-43
    E_h = mlp(np.concatenate([W, C], axis=1), P, "E_init_MLP")
    # model.py:46-51  tile(V_init / sqrt(d))
    V_h = np.tile(P["V_init"] / np.sqrt(dtype(d)), (nV, 1)).astype(dtype)
    E_c = np.zeros_like(E_h)        # graphnn.py:137
    V_c = np.zeros_like(V_h)

    EV = dense_EV(src, dst, nV, dtype) if dense else None
    trace = [] if return_trace else None
    E_c, E_h, V_c, V_h = message_passing(P, src, dst, EV, E_c, E_h, V_c, V_h, time_steps, trace)

    E_vote = mlp(E_h, P, "E_vote").reshape(-1)               # model.py:124-128
    off = np.concatenate([[0], np.cumsum(n_edges)])
    logits = np.array([E_vote[off[k]:off[k + 1]].mean() for k in range(len(n_edges))],
                      dtype=dtype)                           # model.py:134-145
    out = dict(logits=logits, predictions=sigmoid(logits), E_vote=E_vote,
               E_h=E_h, E_c=E_c, V_h=V_h, V_c=V_c)
    if return_trace:
        out["trace"] = trace
    return out


def metrics(logits, route_exists):
    """model.py:147-157 (incl. the reference's own TP/FP/TN/FN definitions)."""
    y = np.asarray(route_exists, dtype=logits.dtype)
    pred = sigmoid(logits)
    r = np.round(pred)
    eq = (y == r).astype(logits.dtype)
    ne = 1 - eq
    # TF: sigmoid_cross_entropy_with_logits = max(l,0) - l*y + log1p(exp(-|l|))
    xent = np.maximum(logits, 0) - logits * y + np.log1p(np.exp(-np.abs(logits)))
    return dict(predictions=pred, loss=xent.mean(), acc=eq.mean(),
                TP=(y * eq).sum(), FP=(y * ne).sum(),
                TN=((1 - y) * eq).sum(), FN=((1 - y) * ne).sum())


# ----------------------------------------------------------------------------
# batch layout restated literally (instance_loader.py:29-80), incl. the dense EV
# ----------------------------------------------------------------------------
def create_batch_ref(instances, dev=0.02, target_cost=None):
    """instances: list of (Ma upper-triangular 0/1, Mw, route).  Loops kept as in the
    reference so it can check the product's vectorised builder."""
    n_instances = len(instances)
    n_vertices = np.array([x[0].shape[0] for x in instances])
    n_edges = np.array([len(np.nonzero(x[0])[0]) for x in instances])
    total_vertices, total_edges = sum(n_vertices), sum(n_edges)
    EV = np.zeros((total_edges, total_vertices))
    W = np.zeros((total_edges, 1))
    C = np.zeros((total_edges, 1))
    route_exists = np.array([i % 2 for i in range(n_instances)])
    for i, (Ma, Mw, route) in enumerate(instances):
        n, m = n_vertices[i], n_edges[i]
        n_acc, m_acc = sum(n_vertices[0:i]), sum(n_edges[0:i])
        edges = list(zip(np.nonzero(Ma)[0], np.nonzero(Ma)[1]))
        for e, (x, y) in enumerate(edges):
            EV[m_acc + e, n_acc + x] = 1
            EV[m_acc + e, n_acc + y] = 1
            W[m_acc + e] = Mw[x, y]
        # instance_loader.py:70 (closing edge uses route[1:]+route[1:], kept verbatim)
        cost = sum([Mw[min(x, y), max(x, y)] for (x, y) in zip(route, route[1:] + route[1:])]) / n
        if target_cost is None:
            C[m_acc:m_acc + m, 0] = (1 - dev) * cost if i % 2 == 0 else (1 + dev) * cost
        else:
            C[m_acc:m_acc + m, 0] = target_cost
    return EV, W, C, route_exists, n_vertices, n_edges


def ev_to_coo(EV):
    """Dense EV -> (src, dst) with src<dst per edge row."""
    r, c = np.nonzero(EV)
    assert len(r) == 2 * EV.shape[0] and np.all(r[0::2] == r[1::2])
    return c[0::2].astype(np.int64), c[1::2].astype(np.int64)


# ----------------------------------------------------------------------------
# scalar pure-Python LN-LSTM (independent of the numpy code above; tiny cases only)
# ----------------------------------------------------------------------------
def lnlstm_scalar(x, c, h, K, gammas, betas):
    """One row, Python floats and loops only.  gammas/betas: dict scope->list."""
    import math
    d = len(h)
    xin = list(x) + list(h)
    z = [sum(xin[k] * K[k][n] for k in range(len(xin))) for n in range(4 * d)]

    def ln(u, g, b):
        mu = sum(u) / len(u)
        var = sum((t - mu) ** 2 for t in u) / len(u)
        inv = 1.0 / math.sqrt(var + LN_EPS)
        return [(t - mu) * inv * g[k] + b[k] for k, t in enumerate(u)]

    def sg(t):
        return 1.0 / (1.0 + math.exp(-t))

    i = ln(z[0:d], gammas["input"], betas["input"])
    j = ln(z[d:2 * d], gammas["transform"], betas["transform"])
    f = ln(z[2 * d:3 * d], gammas["forget"], betas["forget"])
    o = ln(z[3 * d:4 * d], gammas["output"], betas["output"])
    nc = [c[k] * sg(f[k] + FORGET_BIAS) + sg(i[k]) * max(j[k], 0.0) for k in range(d)]
    nc = ln(nc, gammas["state"], betas["state"])
    nh = [max(nc[k], 0.0) * sg(o[k]) for k in range(d)]
    return nc, nh
