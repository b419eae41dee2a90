"""CPU restatement of the TSP-GNN training step (loss, gradients, clip, Adam).  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this
module; the product package (``tsp_gnn_b200``) never does.

PARITY UNPINNED beyond the forward pass: the reference obtains its gradients from
``tf.gradients`` and its update from ``tf.train.AdamOptimizer`` (model.py:157-167), i.e. from
TensorFlow 1.x, which cannot be installed here.  Two independent restatements are kept and
checked against each other (tests/test_oracle_grad.py):

  * ``forward_backward``  -- numpy, hand-derived reverse pass over the forward pass of
                             ``tspgnn_oracle.forward`` (the same recurrence the CUDA kernels use)
  * ``torch_forward_backward`` -- the forward pass written with torch ops, float64,
                             differentiated by torch.autograd

``apply_gradients`` restates model.py:160-167: L2 term, clip_by_global_norm, Adam.
"""
import numpy as np

from . import tspgnn_oracle as orc

LEARNING_RATE = 2e-5          # model.py:13
L2NORM_SCALING = 1e-10        # model.py:14
CLIP_RATIO = 0.65             # model.py:15
ADAM_BETA1, ADAM_BETA2, ADAM_EPS = 0.9, 0.999, 1e-8   # TF: tf.train.AdamOptimizer defaults

CELL = {"V": "TSP/V_cell/layer_norm_basic_lstm_cell", "E": "TSP/E_cell/layer_norm_basic_lstm_cell"}


# ----------------------------------------------------------------------------
# pieces: forward with cache, backward
# ----------------------------------------------------------------------------
def _ln_fwd(u, gamma, beta):
    mean = u.mean(axis=1, keepdims=True)
    var = np.square(u - mean).mean(axis=1, keepdims=True)
    r = 1.0 / np.sqrt(var + u.dtype.type(orc.LN_EPS))
    uh = (u - mean) * r
    return uh * gamma + beta, uh, r


def _ln_bwd(dy, uh, r, gamma):
    """Returns (du, dgamma, dbeta) of y = uh*gamma + beta, uh = (u-mean)*r."""
    dgamma = (dy * uh).sum(axis=0)
    dbeta = dy.sum(axis=0)
    duh = dy * gamma
    du = r * (duh - duh.mean(axis=1, keepdims=True) - uh * (duh * uh).mean(axis=1, keepdims=True))
    return du, dgamma, dbeta


def _mlp_fwd(x, P, prefix, n_layers=4):
    acts = [x]
    for i in range(n_layers):
        x = x @ P["%s_MLP_layer_%d/kernel" % (prefix, i + 1)] + P["%s_MLP_layer_%d/bias" % (prefix, i + 1)]
        if i < n_layers - 1:
            x = np.maximum(x, 0)
        acts.append(x)
    return acts           # acts[0] = input, acts[l] = output of layer l (post-ReLU for hidden layers)


def _mlp_bwd(dy, acts, P, prefix, G, n_layers=4):
    """Accumulates kernel/bias gradients into G and returns d(input)."""
    for i in range(n_layers - 1, -1, -1):
        if i < n_layers - 1:
            dy = dy * (acts[i + 1] > 0)
        G["%s_MLP_layer_%d/kernel" % (prefix, i + 1)] += acts[i].T @ dy
        G["%s_MLP_layer_%d/bias" % (prefix, i + 1)] += dy.sum(axis=0)
        dy = dy @ P["%s_MLP_layer_%d/kernel" % (prefix, i + 1)].T
    return dy


def _lstm_bwd(x, c, h, d_cn, d_hn, P, base, G):
    """Reverse of tspgnn_oracle.lnlstm (recomputes its forward).  Returns (dx, dc, dh)."""
    d = h.shape[1]
    one = h.dtype.type(1.0)
    xh = np.concatenate([x, h], axis=1)
    z = xh @ P[base + "/kernel"]
    names = ("input", "transform", "forget", "output")
    ln = [_ln_fwd(z[:, k * d:(k + 1) * d], P["%s/%s/gamma" % (base, n)], P["%s/%s/beta" % (base, n)])
          for k, n in enumerate(names)]
    i, j, f, o = (t[0] for t in ln)
    si, sf, so = orc.sigmoid(i), orc.sigmoid(f + h.dtype.type(orc.FORGET_BIAS)), orc.sigmoid(o)
    g = np.maximum(j, 0)
    ct = c * sf + si * g
    cn, ch, cr = _ln_fwd(ct, P[base + "/state/gamma"], P[base + "/state/beta"])
    # h' = relu(c') * sig(o)
    d_so = d_hn * np.maximum(cn, 0)
    d_cn = d_cn + d_hn * so * (cn > 0)
    d_ct, dg_s, db_s = _ln_bwd(d_cn, ch, cr, P[base + "/state/gamma"])
    G[base + "/state/gamma"] += dg_s
    G[base + "/state/beta"] += db_s
    dc = d_ct * sf
    dgate = [d_ct * g * si * (one - si),         # input
             d_ct * si * (j > 0),                # transform
             d_ct * c * sf * (one - sf),         # forget
             d_so * so * (one - so)]             # output
    dz = np.empty_like(z)
    for k, n in enumerate(names):
        du, dg_, db_ = _ln_bwd(dgate[k], ln[k][1], ln[k][2], P["%s/%s/gamma" % (base, n)])
        G["%s/%s/gamma" % (base, n)] += dg_
        G["%s/%s/beta" % (base, n)] += db_
        dz[:, k * d:(k + 1) * d] = du
    G[base + "/kernel"] += xh.T @ dz
    dxh = dz @ P[base + "/kernel"].T
    return dxh[:, :d], dc, dxh[:, d:]


def loss_and_dlogits(logits, route_exists, global_batch=None):
    """model.py:157: loss = reduce_mean(sigmoid_cross_entropy_with_logits); d loss / d logits.
    ``global_batch`` is the divisor of the mean (the whole batch when instances are sharded)."""
    y = np.asarray(route_exists, dtype=logits.dtype)
    B = len(logits) if global_batch is None else global_batch
    xent = np.maximum(logits, 0) - logits * y + np.log1p(np.exp(-np.abs(logits)))
    return xent.sum() / B, (orc.sigmoid(logits) - y) / B


def forward_backward(params, src, dst, W, C, n_vertices, n_edges, route_exists, time_steps,
                     dtype=np.float64, global_batch=None):
    """Loss of model.py:157 (without the L2 term, added in apply_gradients) and its gradient
    with respect to every trainable variable.  Returns dict(loss, logits, predictions, grads)."""
    P = {k: v.astype(dtype) for k, v in params.items()}
    G = {k: np.zeros_like(v) for k, v in P.items()}
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    n_edges = np.asarray(n_edges, dtype=np.int64)
    nV, nE = int(np.sum(n_vertices)), int(n_edges.sum())
    d = P["V_init"].shape[1]
    W = np.asarray(W, dtype=dtype).reshape(nE, 1)
    C = np.asarray(C, dtype=dtype).reshape(nE, 1)

    init_acts = _mlp_fwd(np.concatenate([W, C], axis=1), P, "E_init_MLP")
    E_h = init_acts[-1]
    V_h = np.tile(P["V_init"] / np.sqrt(dtype(d)), (nV, 1)).astype(dtype)
    E_c, V_c = np.zeros_like(E_h), np.zeros_like(V_h)
    snaps = []
    for _ in range(int(time_steps)):
        mE = orc.mlp(E_h, P, "TSP/E_msg_V")
        mV = orc.mlp(V_h, P, "TSP/V_msg_E")
        xV = np.zeros((nV, d), dtype=dtype)
        np.add.at(xV, src, mE)
        np.add.at(xV, dst, mE)
        xE = mV[src] + mV[dst]
        snaps.append((E_h, E_c, V_h, V_c, xE, xV))
        nVc, nVh = orc.lnlstm(xV, V_c, V_h, P, CELL["V"])
        nEc, nEh = orc.lnlstm(xE, E_c, E_h, P, CELL["E"])
        V_c, V_h, E_c, E_h = nVc, nVh, nEc, nEh
    vote_acts = _mlp_fwd(E_h, P, "E_vote")
    E_vote = vote_acts[-1].reshape(-1)
    off = np.concatenate([[0], np.cumsum(n_edges)])
    logits = np.array([E_vote[off[k]:off[k + 1]].mean() for k in range(len(n_edges))], dtype=dtype)
    loss, dlogits = loss_and_dlogits(logits, route_exists, global_batch)

    # ---- reverse ----
    dvote = np.repeat(dlogits / n_edges, n_edges).reshape(nE, 1)
    gEh = _mlp_bwd(dvote, vote_acts, P, "E_vote", G)
    gEc = np.zeros_like(gEh)
    gVh = np.zeros((nV, d), dtype=dtype)
    gVc = np.zeros_like(gVh)
    for (E_h, E_c, V_h, V_c, xE, xV) in reversed(snaps):
        dxE, gEc, gEh_l = _lstm_bwd(xE, E_c, E_h, gEc, gEh, P, CELL["E"], G)
        dxV, gVc, gVh_l = _lstm_bwd(xV, V_c, V_h, gVc, gVh, P, CELL["V"], G)
        dmV = np.zeros((nV, d), dtype=dtype)          # (EV)^T . dxE
        np.add.at(dmV, src, dxE)
        np.add.at(dmV, dst, dxE)
        dmE = dxV[src] + dxV[dst]                      # EV . dxV
        gEh = gEh_l + _mlp_bwd(dmE, _mlp_fwd(E_h, P, "TSP/E_msg_V"), P, "TSP/E_msg_V", G)
        gVh = gVh_l + _mlp_bwd(dmV, _mlp_fwd(V_h, P, "TSP/V_msg_E"), P, "TSP/V_msg_E", G)
    _mlp_bwd(gEh, init_acts, P, "E_init_MLP", G)
    G["V_init"] += gVh.sum(axis=0, keepdims=True) / np.sqrt(dtype(d))
    return dict(loss=loss, logits=logits, predictions=orc.sigmoid(logits), grads=G,
                E_h=E_h, V_h=V_h)


# ----------------------------------------------------------------------------
# the same forward pass in torch, differentiated by autograd (independent check)
# ----------------------------------------------------------------------------
def torch_forward_backward(params, src, dst, W, C, n_vertices, n_edges, route_exists, time_steps,
                           global_batch=None):
    import torch
    tp = {k: torch.tensor(np.asarray(v, dtype=np.float64), requires_grad=True) for k, v in params.items()}
    src_t = torch.as_tensor(np.asarray(src, dtype=np.int64))
    dst_t = torch.as_tensor(np.asarray(dst, dtype=np.int64))
    n_edges = np.asarray(n_edges, dtype=np.int64)
    nV, nE = int(np.sum(n_vertices)), int(n_edges.sum())
    d = tp["V_init"].shape[1]

    def mlp(x, prefix):
        for i in range(4):
            x = x @ tp["%s_MLP_layer_%d/kernel" % (prefix, i + 1)] + tp["%s_MLP_layer_%d/bias" % (prefix, i + 1)]
            if i < 3:
                x = torch.relu(x)
        return x

    def ln(u, g, b):
        mean = u.mean(dim=1, keepdim=True)
        var = ((u - mean) ** 2).mean(dim=1, keepdim=True)
        inv = torch.rsqrt(var + orc.LN_EPS) * g
        return u * inv + (b - mean * inv)

    def cell(x, c, h, base):
        z = torch.cat([x, h], dim=1) @ tp[base + "/kernel"]
        i, j, f, o = z[:, :d], z[:, d:2 * d], z[:, 2 * d:3 * d], z[:, 3 * d:]
        i = ln(i, tp[base + "/input/gamma"], tp[base + "/input/beta"])
        j = ln(j, tp[base + "/transform/gamma"], tp[base + "/transform/beta"])
        f = ln(f, tp[base + "/forget/gamma"], tp[base + "/forget/beta"])
        o = ln(o, tp[base + "/output/gamma"], tp[base + "/output/beta"])
        nc = c * torch.sigmoid(f + orc.FORGET_BIAS) + torch.sigmoid(i) * torch.relu(j)
        nc = ln(nc, tp[base + "/state/gamma"], tp[base + "/state/beta"])
        return nc, torch.relu(nc) * torch.sigmoid(o)

    Wt = torch.tensor(np.asarray(W, dtype=np.float64).reshape(nE, 1))
    Ct = torch.tensor(np.asarray(C, dtype=np.float64).reshape(nE, 1))
    E_h = mlp(torch.cat([Wt, Ct], dim=1), "E_init_MLP")
    V_h = (tp["V_init"] / np.sqrt(float(d))).repeat(nV, 1)
    E_c, V_c = torch.zeros_like(E_h), torch.zeros_like(V_h)
    for _ in range(int(time_steps)):
        mE, mV = mlp(E_h, "TSP/E_msg_V"), mlp(V_h, "TSP/V_msg_E")
        xV = torch.zeros(nV, d, dtype=torch.float64).index_add(0, src_t, mE).index_add(0, dst_t, mE)
        xE = mV[src_t] + mV[dst_t]
        nVc, nVh = cell(xV, V_c, V_h, CELL["V"])
        nEc, nEh = cell(xE, E_c, E_h, CELL["E"])
        V_c, V_h, E_c, E_h = nVc, nVh, nEc, nEh
    vote = mlp(E_h, "E_vote").reshape(-1)
    seg = torch.as_tensor(np.repeat(np.arange(len(n_edges)), n_edges))
    logits = torch.zeros(len(n_edges), dtype=torch.float64).index_add(0, seg, vote) / torch.as_tensor(
        n_edges.astype(np.float64))
    y = torch.tensor(np.asarray(route_exists, dtype=np.float64))
    B = len(n_edges) if global_batch is None else global_batch
    xent = torch.clamp(logits, min=0) - logits * y + torch.log1p(torch.exp(-logits.abs()))
    loss = xent.sum() / B
    loss.backward()
    grads = {k: (v.grad.numpy() if v.grad is not None else np.zeros(v.shape)) for k, v in tp.items()}
    return dict(loss=float(loss.detach()), logits=logits.detach().numpy(), grads=grads)


# ----------------------------------------------------------------------------
# model.py:160-167: L2 term, clip_by_global_norm, Adam
# ----------------------------------------------------------------------------
def new_optimizer_state(params):
    return dict(step=0, m={k: np.zeros_like(v, dtype=np.float64) for k, v in params.items()},
                v={k: np.zeros_like(v, dtype=np.float64) for k, v in params.items()})


def apply_gradients(params, grads, state, lr=LEARNING_RATE, l2=L2NORM_SCALING, clip=CLIP_RATIO, dtype=np.float64):
    """TF: vars_cost = sum l2_loss(var) = sum(var**2)/2  =>  gradient l2*var is added to every
    variable's gradient (model.py:163,166); clip_by_global_norm scales all gradients by
    clip / max(global_norm, clip); AdamOptimizer: lr_t = lr*sqrt(1-b2^t)/(1-b1^t),
    m = b1 m + (1-b1) g, v = b2 v + (1-b2) g^2, var -= lr_t * m / (sqrt(v) + eps).
    Returns (new_params, global_norm); ``state`` is updated in place."""
    g = {k: np.asarray(grads[k], dtype=dtype) + dtype(l2) * np.asarray(params[k], dtype=dtype) for k in params}
    gnorm = np.sqrt(sum(float(np.square(v).sum()) for v in g.values()))
    scale = clip / max(gnorm, clip)
    state["step"] += 1
    t = state["step"]
    lr_t = lr * np.sqrt(1.0 - ADAM_BETA2 ** t) / (1.0 - ADAM_BETA1 ** t)
    out = {}
    for k in params:
        gk = g[k] * scale
        state["m"][k] = ADAM_BETA1 * state["m"][k] + (1 - ADAM_BETA1) * gk
        state["v"][k] = ADAM_BETA2 * state["v"][k] + (1 - ADAM_BETA2) * gk * gk
        out[k] = (np.asarray(params[k], dtype=dtype) - lr_t * state["m"][k] / (np.sqrt(state["v"][k]) + ADAM_EPS))
    return out, gnorm
