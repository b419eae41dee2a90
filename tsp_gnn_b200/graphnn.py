"""Host-side mirror of the reference's ``GraphNN`` (graphnn.py:4-272).

Keeps the declarative interface (var / mat / msg / loop dictionaries), the construction
time consistency checks (``check_model``, graphnn.py:72-103) and the run-time shape checks
(``check_run``, graphnn.py:185-271) with the reference's messages.  Execution is mapped onto
the fused CUDA timestep kernels, which implement the topology ``build_network`` declares
(model.py:57-94): two variables, one incidence matrix used plain and transposed, one
message MLP per direction, one LayerNorm-LSTM update per variable.  Any other topology
(transfer functions, matrix-only inputs, several update terms per variable, other sizes,
graphnn.py:142-173) runs update by update on the generic CUDA building blocks of
tsp_gnn_b200/generic.py -- same semantics, one kernel per op, never a CPU path.
"""
import collections
import numpy as np

from .mlp import Mlp

LSTMStateTuple = collections.namedtuple("LSTMStateTuple", ("c", "h"))


class GraphNN(object):
    def __init__(self, var, mat, msg, loop, MLP_depth=3, MLP_weight_initializer="xavier",
                 MLP_bias_initializer="zeros", RNN_cell="LayerNormBasicLSTMCell", Cell_activation="relu",
                 Msg_activation="relu", Msg_last_activation=None, float_dtype="float32", name="GraphNN"):
        self.var, self.mat, self.msg, self.loop, self.name = var, mat, msg, loop, name
        self.MLP_depth = MLP_depth
        self.MLP_weight_initializer = MLP_weight_initializer
        self.MLP_bias_initializer = MLP_bias_initializer
        self.RNN_cell = RNN_cell
        self.Cell_activation = Cell_activation
        self.Msg_activation = Msg_activation
        self.Msg_last_activation = Msg_last_activation
        self.float_dtype = float_dtype
        self.check_model()
        self._init_parameters()
        self._engine = None
        try:
            self._kernel_roles = self._match_fused_topology()
        except NotImplementedError:
            self._kernel_roles = None          # executed on the generic building blocks
        self._generic_params = None
        self._generic_device = {}

    # graphnn.py:72-103 -------------------------------------------------------------
    def check_model(self):
        for v in self.var:
            if v not in self.loop:
                raise Warning("Variable {v} is not updated anywhere! Consider removing it from the model".format(v=v))
        for v in self.loop:
            if v not in self.var:
                raise Exception("Updating variable {v}, which has not been declared!".format(v=v))
        for mat, (v1, v2) in self.mat.items():
            if v1 not in self.var:
                raise Exception("Matrix {mat} definition depends on undeclared variable {v}".format(mat=mat, v=v1))
            if v2 not in self.var and type(v2) is not int:
                raise Exception("Matrix {mat} definition depends on undeclared variable {v}".format(mat=mat, v=v2))
        for msg, (v1, v2) in self.msg.items():
            if v1 not in self.var:
                raise Exception("Message {msg} maps from undeclared variable {v}".format(msg=msg, v=v1))
            if v2 not in self.var:
                raise Exception("Message {msg} maps to undeclared variable {v}".format(msg=msg, v=v2))

    # graphnn.py:105-126 ------------------------------------------------------------
    def _init_parameters(self):
        self._RNN_cells = {v: {"num_units": d, "activation": self.Cell_activation, "cell": self.RNN_cell}
                           for (v, d) in self.var.items()}
        self._msg_MLPs = {
            msg: Mlp(layer_sizes=[self.var[vin] for _ in range(self.MLP_depth)],
                     output_size=self.var[vout],
                     activations=[self.Msg_activation for _ in range(self.MLP_depth)],
                     output_activation=self.Msg_last_activation,
                     kernel_initializer=self.MLP_weight_initializer,
                     bias_initializer=self.MLP_weight_initializer,      # graphnn.py:121 (sic)
                     name=msg, name_internal_layers=True)
            for msg, (vin, vout) in self.msg.items()
        }

    def variable_names(self):
        names = []
        for msg in self._msg_MLPs.values():
            names += msg.variable_names(scope=self.name + "/")
        for v in self.var:
            base = "{}/{}_cell/layer_norm_basic_lstm_cell".format(self.name, v)
            names.append(base + "/kernel")
            for g in ("input", "transform", "forget", "output", "state"):
                names += ["{}/{}/gamma".format(base, g), "{}/{}/beta".format(base, g)]
        return names

    # -------------------------------------------------------------------------------
    def _match_fused_topology(self):
        """Maps the declaration onto the roles of the fused kernels or raises."""
        def bad(why):
            raise NotImplementedError(
                "GraphNN topology not supported by the fused CUDA path (%s). Built: the TSP wiring of "
                "model.py:57-94 -- two variables joined by one incidence matrix, one message MLP per "
                "direction, LayerNormBasicLSTMCell/relu updates, MLP_depth=3, d=64." % why)
        if len(self.var) != 2 or len(self.mat) != 1 or len(self.msg) != 2:
            bad("need exactly 2 variables, 1 matrix, 2 messages")
        if self.MLP_depth != 3 or self.Msg_activation != "relu" or self.Msg_last_activation is not None:
            bad("message MLPs must be 3x relu + linear output")
        if self.RNN_cell != "LayerNormBasicLSTMCell" or self.Cell_activation != "relu":
            bad("cells must be LayerNormBasicLSTMCell with relu")
        if self.float_dtype != "float32":
            bad("float_dtype must be float32")
        (mname, (row_var, col_var)), = self.mat.items()
        if type(col_var) is int or row_var == col_var:
            bad("matrix must join two distinct variables")
        if any(d != 64 for d in self.var.values()):
            bad("embedding size must be 64")
        roles = {"mat": mname, "row_var": row_var, "col_var": col_var}
        for v, other, transposed in ((row_var, col_var, False), (col_var, row_var, True)):
            ups = self.loop[v]
            if len(ups) != 1:
                bad("each variable takes exactly one update term")
            u = ups[0]
            if u.get("mat") != mname or u.get("var") != other or "fun" in u or "msg" not in u:
                bad("update of %s must be mat x msg(var)" % v)
            if bool(u.get("transpose?", False)) != transposed:
                bad("transpose? flag of %s" % v)
            if tuple(self.msg[u["msg"]]) != (other, v):
                bad("message %s must map %s -> %s" % (u["msg"], other, v))
            roles["msg_to_" + v] = u["msg"]
        return roles

    # graphnn.py:185-271 ------------------------------------------------------------
    def check_run(self, adjacency_matrices, initial_embeddings, time_steps, LSTM_initial_states):
        num_vars = {}
        for v, d in self.var.items():
            init_shape = tuple(initial_embeddings[v].shape)
            num_vars[v] = init_shape[0]
            if init_shape[1] != d:
                raise ValueError("Initial embedding of variable {v} doesn't have the same dimensionality {d} as "
                                 "declared".format(v=v, d=d))
            if v in LSTM_initial_states:
                ls = tuple(LSTM_initial_states[v].shape)
                if ls[1] != d:
                    raise ValueError("Initial hidden state of variable {v}'s LSTM doesn't have the same "
                                     "dimensionality {d} as declared".format(v=v, d=d))
                if ls != init_shape:
                    raise ValueError("Initial embeddings of variable {v} don't have the same shape as the its "
                                     "LSTM's initial hidden state".format(v=v))
        for mat, (v1, v2) in self.mat.items():
            mshape = tuple(adjacency_matrices[mat].shape)
            if mshape[0] != num_vars[v1]:
                raise ValueError("Matrix {m} doesn't have the same number of nodes as the initial embeddings of its "
                                 "variable {v}".format(v=v1, m=mat))
            if type(v2) is int:
                if mshape[1] != v2:
                    raise ValueError("Matrix {m} doesn't have the same dimensionality {d} on the second variable as "
                                     "declared".format(m=mat, d=v2))
            elif mshape[1] != num_vars[v2]:
                raise ValueError("Matrix {m} doesn't have the same number of nodes as the initial embeddings of its "
                                 "variable {v}".format(v=v2, m=mat))

    # graphnn.py:128-183 ------------------------------------------------------------
    def bind(self, engine):
        """Attach the CUDA engine (parameters + plan already set) that executes this network."""
        self._engine = engine
        return self

    # -- generic execution (any topology) ----------------------------------------------------------------
    def input_width(self, v):
        """Width of the concatenated cell input of variable v (graphnn.py:146-167)."""
        w = 0
        for u in self.loop[v]:
            if "var" in u:
                w += self.var[self.msg[u["msg"]][1]] if "msg" in u else self.var[u["var"]]
            else:
                second = self.mat[u["mat"]][1]
                w += second if type(second) is int else self.var[second]
        return w

    def init_parameters(self, seed=0):
        """Variables of the generic path under the reference's initialisers: message MLP kernels AND biases
        xavier-uniform (graphnn.py:120-121), LSTM kernels glorot-uniform, LayerNorm gamma 1 / beta 0.
        Returns {TF variable name: array}; names as variable_names()."""
        rng = np.random.RandomState(seed)
        params = {}
        for msg, (vin, vout) in self.msg.items():
            for k, a in self._msg_MLPs[msg].init_parameters(self.var[vin], rng).items():
                params["%s/%s" % (self.name, k)] = a
        for v, d in self.var.items():
            base = "{}/{}_cell/layer_norm_basic_lstm_cell".format(self.name, v)
            fan_in, fan_out = self.input_width(v) + d, 4 * d
            lim = np.sqrt(6.0 / (fan_in + fan_out))
            params[base + "/kernel"] = rng.uniform(-lim, lim, size=(fan_in, fan_out)).astype(np.float32)
            for g in ("input", "transform", "forget", "output", "state"):
                params["{}/{}/gamma".format(base, g)] = np.ones(d, dtype=np.float32)
                params["{}/{}/beta".format(base, g)] = np.zeros(d, dtype=np.float32)
        self.set_parameters(params)
        return params

    def set_parameters(self, params):
        """``params``: {TF variable name: array} for every name of variable_names()."""
        missing = [n for n in self.variable_names() if n not in params]
        if missing:
            raise KeyError("missing variables: %r" % (missing[:4],))
        self._generic_params = {k: np.asarray(params[k], dtype=np.float32) for k in self.variable_names()}
        for msg in self._msg_MLPs.values():
            msg.set_parameters(self._generic_params, scope=self.name + "/")
        self._generic_device = {}

    def _cell_params(self, v, device):
        import torch
        key = (v, device)
        if key not in self._generic_device:
            base = "{}/{}_cell/layer_norm_basic_lstm_cell".format(self.name, v)
            gates = ("input", "transform", "forget", "output", "state")
            P = self._generic_params
            to = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device)
            self._generic_device[key] = (to(P[base + "/kernel"]),
                                         to(np.stack([P["{}/{}/gamma".format(base, g)] for g in gates])),
                                         to(np.stack([P["{}/{}/beta".format(base, g)] for g in gates])))
        return self._generic_device[key]

    def _call_generic(self, adjacency_matrices, initial_embeddings, time_steps, LSTM_initial_states):
        """graphnn.py:134-179 op by op: for every variable the update terms ( fun -> message MLP -> matrix
        product | the matrix itself ) are concatenated and fed to its LayerNorm-LSTM cell; every variable reads
        the time-t states.  CUDA tensors in, CUDA tensors out; ``fun`` entries are callables on CUDA tensors."""
        import torch
        from . import generic
        if self.RNN_cell != "LayerNormBasicLSTMCell":
            raise NotImplementedError("only LayerNormBasicLSTMCell cells are built (the reference's default, graphnn.py:14)")
        if self._generic_params is None:
            raise RuntimeError("Attempting to use uninitialized variables: call init_parameters() or set_parameters() first")
        def dev_tensor(a):
            t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
            return (t if t.is_cuda else t.cuda()).float().contiguous()
        h = {v: dev_tensor(a) for v, a in initial_embeddings.items()}
        device = next(iter(h.values())).device
        c = {v: (dev_tensor(LSTM_initial_states[v]) if v in LSTM_initial_states else torch.zeros_like(h[v])) for v in h}
        used = {u["mat"] for ups in self.loop.values() for u in ups if "mat" in u}
        as_coo = lambda M: M if isinstance(M, generic.CooMatrix) else generic.CooMatrix(M, device)
        coo = {m: as_coo(adjacency_matrices[m]) for m in used
               if any("var" in u and u.get("mat") == m for ups in self.loop.values() for u in ups)}
        raw = {m: dev_tensor(adjacency_matrices[m]) for m in used
               if any("var" not in u and u.get("mat") == m for ups in self.loop.values() for u in ups)}
        for _ in range(int(time_steps)):
            new_c, new_h = {}, {}
            for v in self.var:
                inputs = []
                for u in self.loop[v]:
                    if "var" in u:
                        y = h[u["var"]]
                        if "fun" in u:
                            y = u["fun"](y)
                        if "msg" in u:
                            y = self._msg_MLPs[u["msg"]](y)
                        if "mat" in u:
                            y = coo[u["mat"]].matmul(y, transpose=bool(u.get("transpose?", False)))
                        inputs.append(y)
                    else:
                        inputs.append(raw[u["mat"]])
                x = inputs[0] if len(inputs) == 1 else torch.cat(inputs, dim=1)
                kernel, gamma, beta = self._cell_params(v, device)
                new_c[v], new_h[v] = generic.lnlstm(x, c[v], h[v], kernel, gamma, beta, activation=self.Cell_activation)
            c, h = new_c, new_h
        return {v: LSTMStateTuple(c=c[v], h=h[v]) for v in self.var}

    def __call__(self, adjacency_matrices, initial_embeddings, time_steps, LSTM_initial_states={}):
        """Runs ``time_steps`` message-passing iterations.

        The TSP topology bound to an engine (build_network / Session) runs on the fused kernels:
        adjacency_matrices[mat] is only shape-checked (the engine's plan holds the incidence structure);
        initial_embeddings / LSTM_initial_states are row-major fp32 CUDA tensors [N_var, d].  Every other
        case runs on the generic building blocks with this object's own parameters (init_parameters /
        set_parameters).  Returns {var: LSTMStateTuple(c, h)} of CUDA tensors.
        """
        self.check_run(adjacency_matrices, initial_embeddings, time_steps, LSTM_initial_states)
        if self._engine is None or self._kernel_roles is None:
            return self._call_generic(adjacency_matrices, initial_embeddings, time_steps, LSTM_initial_states)
        import torch
        eng, R = self._engine, self._kernel_roles
        row, col = R["row_var"], R["col_var"]      # row variable = edges 'E', column variable = vertices 'V'
        zeros = lambda t: torch.zeros_like(t)
        Eh, Vh = initial_embeddings[row].contiguous().float(), initial_embeddings[col].contiguous().float()
        Ec = LSTM_initial_states[row].contiguous().float() if row in LSTM_initial_states else zeros(Eh)
        Vc = LSTM_initial_states[col].contiguous().float() if col in LSTM_initial_states else zeros(Vh)
        eng.set_states(Vh=Vh, Vc=Vc, Eh=Eh, Ec=Ec)
        eng.step(int(time_steps))
        st = eng.get_states()
        return {col: LSTMStateTuple(c=st["V"][0], h=st["V"][1]), row: LSTMStateTuple(c=st["E"][0], h=st["E"][1])}
