"""Host-side mirror of the reference's ``GraphNN`` (graphnn.py:4-272).

Keeps the declarative interface (var / mat / msg / loop dictionaries), the construction
time consistency checks (``check_model``, graphnn.py:72-103) and the run-time shape checks
(``check_run``, graphnn.py:185-271) with the reference's messages.  Execution is mapped onto
the fused CUDA timestep kernels, which implement the topology ``build_network`` declares
(model.py:57-94): two variables, one incidence matrix used plain and transposed, one
message MLP per direction, one LayerNorm-LSTM update per variable.  Any other topology is
rejected loudly -- there is no generic fallback path.
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
        self._kernel_roles = self._match_fused_topology()

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

    def __call__(self, adjacency_matrices, initial_embeddings, time_steps, LSTM_initial_states={}):
        """Runs ``time_steps`` message-passing iterations on the bound engine.

        adjacency_matrices[mat] is only shape-checked here (the engine's plan holds the
        incidence structure); initial_embeddings / LSTM_initial_states are row-major fp32
        CUDA tensors [N_var, d].  Returns {var: LSTMStateTuple(c, h)} of CUDA tensors.
        """
        if self._engine is None:
            raise RuntimeError("GraphNN is not bound to an engine; use build_network()/Session or GraphNN.bind")
        import torch
        self.check_run(adjacency_matrices, initial_embeddings, time_steps, LSTM_initial_states)
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
