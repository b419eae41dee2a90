"""``build_network(d)`` and a minimal ``Session`` so that train.py / test.py style drivers of
the reference run unchanged against the CUDA engine (model.py:9-171, train.py:17-63).

``build_network`` returns a dict with the reference's keys; the values are light handles
(:class:`Placeholder`, :class:`Fetch`) that ``Session.run(fetches, feed_dict)`` resolves:

    GNN = build_network(64)
    with Session(GNN) as sess:
        sess.run(global_variables_initializer())
        loss, acc, predictions, TP, FP, TN, FN = sess.run(
            [GNN['loss'], GNN['acc'], GNN['predictions'], GNN['TP'], GNN['FP'], GNN['TN'], GNN['FN']],
            feed_dict={GNN['EV']: EV, GNN['W']: W, GNN['C']: C, GNN['time_steps']: 32,
                       GNN['route_exists']: y, GNN['n_vertices']: nv, GNN['n_edges']: ne})

``EV`` may be the reference's dense ``[sumE,sumV]`` array or an ``instances.Incidence``.
Fetching ``train_step`` runs the reverse pass, the L2 term, the global-norm clip and one Adam
step on the device (model.py:157-167); the other fetches of the same ``run`` are the values of
the pre-update variables, as in the reference's single ``sess.run``.
"""
import numpy as np

from .graphnn import GraphNN, LSTMStateTuple
from .mlp import Mlp
from .instances import Incidence
from . import params as _params


class Placeholder(object):
    def __init__(self, name, dtype, shape):
        self.name, self.dtype, self.shape = name, dtype, shape

    def __repr__(self):
        return "<Placeholder %s>" % self.name


class Fetch(object):
    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return "<Fetch %s>" % self.name


class _InitOp(object):
    def __init__(self, seed=None):
        self.seed = seed


def global_variables_initializer(seed=None):
    """tf.global_variables_initializer() analogue (train.py:210): draws every trainable
    variable from the reference's initialisers."""
    return _InitOp(seed)


def build_network(d, mode="bf16x3", device=0):
    # hyper-parameters of model.py:13-15
    learning_rate = 2e-5
    l2norm_scaling = 1e-10
    global_norm_gradient_clipping_ratio = 0.65

    route_exists = Placeholder("route_exists", np.float32, (None,))          # model.py:18
    n_vertices = Placeholder("n_vertices", np.int32, (None,))                # model.py:20
    n_edges = Placeholder("edges", np.int32, (None,))                        # model.py:21
    EV_matrix = Placeholder("EV", np.float32, (None, None))                  # model.py:23
    edge_weight = Placeholder("edge_weight", np.float32, (None, 1))          # model.py:25
    target_cost = Placeholder("target_cost", np.float32, (None, 1))          # model.py:27
    time_steps = Placeholder("time_steps", np.int32, ())                     # model.py:29

    edge_init_MLP = Mlp(layer_sizes=[d / 8, d / 4, d / 2], activations=["relu" for _ in range(3)], output_size=d,
                        name="E_init_MLP", name_internal_layers=True, kernel_initializer="xavier",
                        bias_initializer="zeros")                           # model.py:33-41

    gnn = GraphNN(
        {"V": d, "E": d},
        {"EV": ("E", "V")},
        {"V_msg_E": ("V", "E"), "E_msg_V": ("E", "V")},
        {
            "V": [{"mat": "EV", "msg": "E_msg_V", "transpose?": True, "var": "E"}],   # V(t+1) <- Vu(EV^T x E_msg_V(E(t)))
            "E": [{"mat": "EV", "msg": "V_msg_E", "var": "V"}],                       # E(t+1) <- Eu(EV x V_msg_E(V(t)))
        },
        name="TSP")                                                          # model.py:57-94
    if int(d) != d or d < 8 or int(d / 8) < 1:
        raise ValueError("build_network(d=%r): d must be an integer >= 8 (E_init_MLP has a layer of d/8 units, "
                         "model.py:34)" % (d,))
    # d == 64 (train.py:108, the reference's default) runs on the fused kernels; any other size runs the same graph
    # op by op on the generic CUDA building blocks (inference fetches only, see _GenericRunner)

    E_vote_MLP = Mlp(layer_sizes=[d for _ in range(3)], activations=["relu" for _ in range(3)], output_size=1,
                     name="E_vote", name_internal_layers=True, kernel_initializer="xavier",
                     bias_initializer="zeros")                              # model.py:107-115

    GNN = {}
    GNN["gnn"] = gnn
    GNN["route_exists"] = route_exists
    GNN["n_vertices"] = n_vertices
    GNN["n_edges"] = n_edges
    GNN["EV"] = EV_matrix
    GNN["W"] = edge_weight
    GNN["C"] = target_cost
    GNN["time_steps"] = time_steps
    GNN["last_states"] = Fetch("last_states")                               # model.py:123
    for k in ("predictions", "TP", "FP", "TN", "FN", "acc", "loss", "train_step"):
        GNN[k] = Fetch(k)                                                    # model.py:147-167
    # not part of the reference dict: construction-time settings Session needs
    GNN["_config"] = {"d": d, "mode": mode, "device": device, "E_init_MLP": edge_init_MLP, "E_vote_MLP": E_vote_MLP,
                      "learning_rate": learning_rate, "l2norm_scaling": l2norm_scaling,
                      "clip": global_norm_gradient_clipping_ratio}
    return GNN


def _metrics(logits, predictions, route_exists):
    """model.py:150-157, evaluated on the host from the [B] logits the device returns."""
    y = np.asarray(route_exists, dtype=np.float32)
    l = logits.astype(np.float32)
    r = np.round(predictions)
    eq = (y == r).astype(np.float32)
    ne = 1.0 - eq
    xent = np.maximum(l, 0) - l * y + np.log1p(np.exp(-np.abs(l)))   # sigmoid_cross_entropy_with_logits
    return {
        "TP": np.float32((y * eq).sum()), "FP": np.float32((y * ne).sum()),       # model.py:150-151 (reference's naming)
        "TN": np.float32(((1 - y) * eq).sum()), "FN": np.float32(((1 - y) * ne).sum()),
        "acc": np.float32(eq.mean()), "loss": np.float32(xent.mean()),
    }


class _GenericRunner(object):
    """build_network(d != 64) behind ``Session.run``: model.py:33-51,118-147 executed op by op on the generic
    CUDA building blocks of tsp_gnn_b200/generic.py (dense layer, COO matrix product, LayerNorm-LSTM cell of any
    size) -- the reference's ``-d`` flag (train.py:108) for the inference fetches.  Every arithmetic step is a
    libtspgnn kernel: the vertex tiling is a dense layer on a column of ones, the per-instance vote mean a COO
    product with 1/n_edges entries, the sigmoid a 1x1 dense layer.  The reverse pass exists for d = 64 only."""

    def __init__(self, GNN):
        cfg = GNN["_config"]
        self.d, self.device = int(cfg["d"]), cfg["device"]
        self.gnn, self.e_init, self.e_vote = GNN["gnn"], cfg["E_init_MLP"], cfg["E_vote_MLP"]
        self._vinit = None
        self._plan = None
        self.states = None

    def set_params(self, params):
        import torch
        self.e_init.set_parameters(params)
        self.e_vote.set_parameters(params)
        self.gnn.set_parameters(params)
        dev = torch.device("cuda", self.device)
        v = np.asarray(params["V_init"], dtype=np.float32).reshape(1, self.d) / np.float32(np.sqrt(np.float32(self.d)))
        self._vinit = torch.from_numpy(np.ascontiguousarray(v)).to(dev)          # model.py:46-51
        self._one = torch.ones(1, 1, dtype=torch.float32, device=dev)

    def plan(self, nv, ne, src, dst):
        import torch
        from . import generic
        dev = torch.device("cuda", self.device)
        nE, nV, B = int(ne.sum()), int(nv.sum()), int(ne.shape[0])
        e = np.arange(nE, dtype=np.int32)
        EV = generic.CooMatrix.from_entries(np.concatenate([e, e]), np.concatenate([src, dst]), None, (nE, nV), dev)
        inst_of_edge = np.repeat(np.arange(B, dtype=np.int32), ne)
        mean = generic.CooMatrix.from_entries(inst_of_edge, e, (1.0 / ne.astype(np.float64))[inst_of_edge], (B, nE), dev)
        self._plan = (EV, mean, torch.ones(nV, 1, dtype=torch.float32, device=dev), nE, nV, B)

    def forward(self, W, C, time_steps):
        import torch
        from . import generic
        EV, mean, ones_v, nE, nV, B = self._plan
        dev = ones_v.device
        with torch.cuda.device(dev):
            wc = torch.from_numpy(np.ascontiguousarray(np.stack([W, C], axis=1), dtype=np.float32)).to(dev)
            Eh = self.e_init(wc)                                              # model.py:33-43
            Vh = generic.dense(ones_v, self._vinit)                           # model.py:46-51: tile(V_init / sqrt(d))
            st = self.gnn._call_generic({"EV": EV}, {"V": Vh, "E": Eh}, time_steps, {})
            votes = self.e_vote(st["E"].h)                                    # model.py:124-128
            logits = mean.matmul(votes)                                       # model.py:134-145
            preds = generic.dense(logits, self._one, None, "sigmoid")         # model.py:147
            self.states = st
            return logits.reshape(-1).cpu().numpy(), preds.reshape(-1).cpu().numpy()


class Session(object):
    """tf.Session stand-in bound to one GPU engine."""

    def __init__(self, GNN=None, config=None):
        self._gnn = GNN
        self._engine = None
        self._params = None
        self._params_stale = False      # the device holds newer variables than self._params
        self._plan_key = None
        self._generic = None            # _GenericRunner when the network is not the fused kernels' d = 64
        if GNN is not None and GNN["gnn"]._kernel_roles is None:
            self._generic = _GenericRunner(GNN)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def close(self):
        if self._engine is not None:
            self._engine.close()
            self._engine = None

    # -- variables ----------------------------------------------------------------
    def _ensure_engine(self):
        if self._engine is None:
            from .engine import Engine
            cfg = self._gnn["_config"]
            self._engine = Engine(cfg["d"], cfg["mode"], cfg["device"])
            self._engine.set_hyper(cfg["learning_rate"], cfg["l2norm_scaling"], cfg["clip"])
            self._gnn["gnn"].bind(self._engine)
        return self._engine

    def set_variables(self, params):
        self._params = {k: np.asarray(v, dtype=np.float32) for k, v in params.items()}
        self._params_stale = False
        if self._generic is not None:
            self._generic.set_params(self._params)
        else:
            self._ensure_engine().set_params(self._params)

    def get_variables(self):
        if self._params_stale:
            self._params = _params.unflatten(self._engine.get_params(), self._gnn["_config"]["d"])
            self._params_stale = False
        return dict(self._params)

    def get_optimizer_state(self):
        """Adam slots + step (tf.train.Saver stores them with the variables, util.py:35)."""
        if self._generic is not None:
            return None
        return self._ensure_engine().get_optimizer_state()

    def set_optimizer_state(self, state):
        self._ensure_engine().set_optimizer_state(state)

    def load_weights(self, path):
        """util.load_weights: tf.train.Saver().restore brings back every global variable, i.e. the Adam
        slots and beta powers too (util.py:13-17), so a resumed run continues the same optimizer trajectory."""
        ckpt = _params.load_checkpoint(path)
        self.set_variables(_params.load_weights(path))
        opt = _params.named_to_optimizer_state(ckpt, self._gnn["_config"]["d"])
        if opt is not None and self._generic is None:
            self.set_optimizer_state(opt)

    def save_weights(self, path):
        """util.save_weights: variables + Adam slots + beta powers (tf.train.Saver() default var_list)."""
        opt = self.get_optimizer_state() if (self._engine is not None and self._generic is None) else None
        _params.save_weights(self.get_variables(), path, optimizer_state=opt)

    # -- run ------------------------------------------------------------------------
    def run(self, fetches, feed_dict=None):
        if self._gnn is None:
            raise RuntimeError("Session was created without a network")
        if isinstance(fetches, _InitOp):
            self.set_variables(_params.init_params(self._gnn["_config"]["d"], seed=fetches.seed))
            return None
        single = not isinstance(fetches, (list, tuple))
        flist = [fetches] if single else list(fetches)
        feed = {}
        for k, v in (feed_dict or {}).items():
            feed[k.name if isinstance(k, Placeholder) else str(k)] = v
        names = [f.name for f in flist]
        train = "train_step" in names
        if self._params is None:
            raise RuntimeError("Attempting to use uninitialized variables: run global_variables_initializer() "
                               "or load_weights() first")
        for need in ("EV", "edge_weight", "target_cost", "time_steps", "n_vertices", "edges"):
            if need not in feed:
                raise ValueError("You must feed a value for placeholder %r" % need)
        if train and self._generic is not None:
            raise NotImplementedError(
                "train_step at d=%d: the reverse pass is built for the reference's default embedding size d=64 "
                "(train.py:108); other sizes run the inference fetches on the generic CUDA path"
                % self._gnn["_config"]["d"])
        eng = self._generic if self._generic is not None else self._ensure_engine()
        EV = feed["EV"]
        if not isinstance(EV, Incidence):
            from .engine import dense_ev_to_coo
            EVd = np.asarray(EV)
            src, dst = dense_ev_to_coo(EVd)
            EV = Incidence(src, dst, EVd.shape[1])
        nv = np.asarray(feed["n_vertices"]).astype(np.int64)
        ne = np.asarray(feed["edges"]).astype(np.int64)
        W = np.asarray(feed["edge_weight"], dtype=np.float32).reshape(-1)
        C = np.asarray(feed["target_cost"], dtype=np.float32).reshape(-1)
        # shape checks of graphnn.check_run on the fed matrices
        if EV.shape[0] != int(ne.sum()) or EV.shape[0] != W.shape[0] or W.shape != C.shape:
            raise ValueError("Matrix EV doesn't have the same number of nodes as the initial embeddings of its variable E")
        if EV.shape[1] != int(nv.sum()):
            raise ValueError("Matrix EV doesn't have the same number of nodes as the initial embeddings of its variable V")
        key = (EV.src.tobytes(), EV.dst.tobytes(), nv.tobytes(), ne.tobytes())
        if key != self._plan_key:
            eng.plan(nv, ne, EV.src, EV.dst)
            self._plan_key = key
        if train:
            if "route_exists" not in feed:
                raise ValueError("You must feed a value for placeholder 'route_exists'")
            _, logits, preds = eng.train_step_host(W, C, feed["route_exists"], int(feed["time_steps"]))
            self._params_stale = True
        elif self._generic is not None:
            logits, preds = eng.forward(W, C, int(feed["time_steps"]))
        else:
            logits, preds = eng.forward_host(W, C, int(feed["time_steps"]))
        out = {"predictions": preds, "logits": logits, "train_step": None}
        if any(n in names for n in ("TP", "FP", "TN", "FN", "acc", "loss")):
            if "route_exists" not in feed:
                raise ValueError("You must feed a value for placeholder 'route_exists'")
            out.update(_metrics(logits, preds, feed["route_exists"]))
        if "last_states" in names and self._generic is not None:
            out["last_states"] = {v: LSTMStateTuple(c=t.c.cpu().numpy(), h=t.h.cpu().numpy())
                                  for v, t in eng.states.items()}
        elif "last_states" in names:
            st = eng.get_states()
            out["last_states"] = {v: LSTMStateTuple(c=st[v][0].cpu().numpy(), h=st[v][1].cpu().numpy())
                                  for v in ("V", "E")}
        res = [out[n] for n in names]
        return res[0] if single else res
