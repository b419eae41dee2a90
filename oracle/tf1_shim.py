"""A tiny numpy stand-in for the TensorFlow 1.x API surface that the reference's model.py,
graphnn.py and mlp.py touch.  TEST INFRASTRUCTURE ONLY (see oracle/tspgnn_oracle.py).

Purpose: TensorFlow 1.x cannot be installed here, so the reference cannot run as shipped.  With
this module registered as ``tensorflow`` the reference's OWN graph-construction code
(``build_network`` -> ``GraphNN.__call__`` -> ``Mlp.__call__``, /root/reference/*.py, unmodified
and imported in place) executes on numpy.  That pins everything the reference's Python decides --
topology wiring, op order, which matrix is transposed, concat order, state-tuple handling, the
read-out loop, metric definitions and the variable names its scopes produce -- against the
oracle restatement.  What it cannot pin is the arithmetic inside TensorFlow itself
(LayerNormBasicLSTMCell, layer_norm, Dense, initialisers): those are restated below from the
TF 1.x sources ("TF:" comments) exactly as in the oracle, so parity with real TensorFlow stays
unpinned for them.

Design: deferred evaluation.  Every op returns a ``Tensor`` holding a closure; ``Session.run``
evaluates closures against the feed dict.  ``while_loop`` runs as a Python loop at evaluation
time, calling the reference's body function on constant tensors each iteration.
Use ``install()`` to register the module and ``reset()`` between graphs.

``reset(backend="torch")`` evaluates the same closures on float64 torch tensors, which gives the
reference's training graph (model.py:157-167) something to differentiate: ``tf.gradients`` becomes
``torch.autograd.grad`` over the graph the reference's code built, ``tf.clip_by_global_norm`` and
``tf.train.AdamOptimizer`` are restated from the TF 1.x sources ("TF:" comments), and the gradients
/ Adam slots of the last step can be read back with ``last_gradients()`` / ``optimizer_slots()``.
"""
import sys
import types
import contextlib

import numpy as np

float32 = np.float32
int32 = np.int32

_STATE = {"scope": [], "variables": {}, "trainable": [], "dtype": np.float64, "env": None, "init_rng": None,
          "backend": "numpy"}


def reset(dtype=np.float64, seed=0, backend="numpy"):
    _STATE.update(scope=[], variables={}, trainable=[], assertions=[], dtype=dtype, env=None,
                  init_rng=np.random.RandomState(seed), backend=backend, last_grads={}, adam={})


# ----------------------------------------------------------------------------------------
# array backend: numpy (default) or float64 torch tensors (differentiable)
# ----------------------------------------------------------------------------------------
def _torch_mode():
    return _STATE.get("backend") == "torch"


def _T():
    import torch
    return torch


def _is_t(a):
    return _torch_mode() and isinstance(a, _T().Tensor)


def _farr(a):
    """value -> float array of the working dtype"""
    if _torch_mode():
        t = _T()
        return a.to(t.float64) if isinstance(a, t.Tensor) else t.as_tensor(np.asarray(a, dtype=np.float64))
    return np.asarray(a).astype(_STATE["dtype"])


def _iarr(a):
    if _torch_mode():
        t = _T()
        return a.to(t.int64) if isinstance(a, t.Tensor) else t.as_tensor(np.asarray(a).astype(np.int64))
    return np.asarray(a).astype(np.int64)


def _cat(xs, axis):
    return _T().cat(list(xs), dim=axis) if _torch_mode() else np.concatenate(list(xs), axis=axis)


def _mean(a, axis=None, keepdims=False):
    if _is_t(a):
        return a.mean() if axis is None else a.mean(dim=axis, keepdim=keepdims)
    return np.mean(a) if axis is None else a.mean(axis=axis, keepdims=keepdims)


def _sum(a):
    return a.sum() if _is_t(a) else np.sum(a)


def _fn(name, a):
    """elementwise sqrt / exp / abs / log1p / round / square"""
    if _is_t(a):
        return getattr(_T(), name)(a)
    return getattr(np, name)(a)


def _relu_v(a):
    return _T().clamp(a, min=0) if _is_t(a) else np.maximum(a, 0)


def _sigmoid_v(a):
    e = _fn("exp", -_fn("abs", a))
    if _is_t(a):
        return _T().where(a >= 0, 1.0 / (1.0 + e), e / (1.0 + e))
    return np.where(a >= 0, 1.0 / (1.0 + e), e / (1.0 + e))


def _shape_of(a):
    return tuple(a.shape) if hasattr(a, "shape") else np.shape(a)


# ----------------------------------------------------------------------------------------
# tensors
# ----------------------------------------------------------------------------------------
class Tensor(object):
    def __init__(self, fn, name=None):
        self._fn, self.name = fn, name

    def eval(self):
        env = _STATE["env"]
        if env is None:
            raise RuntimeError("tensor evaluated outside Session.run")
        key = id(self)
        if key not in env["cache"]:
            env["cache"][key] = (self, self._fn())     # holding `self` keeps the id from being recycled
        return env["cache"][key][1]

    # the operators the reference uses on tensors
    def __add__(self, o):
        return Tensor(lambda: _v(self) + _v(o))

    __radd__ = __add__

    def __sub__(self, o):
        return Tensor(lambda: _v(self) - _v(o))

    def __rsub__(self, o):
        return Tensor(lambda: _v(o) - _v(self))

    def __mul__(self, o):
        return Tensor(lambda: _v(self) * _v(o))

    def __getitem__(self, idx):
        def ev():
            if isinstance(idx, slice):
                return _v(self)[slice(_i(idx.start), _i(idx.stop), _i(idx.step))]
            return _v(self)[_i(idx)]
        return Tensor(ev)

    def __hash__(self):
        return id(self)

    def __eq__(self, o):
        return self is o


class Placeholder(Tensor):
    def __init__(self, dtype, shape, name):
        Tensor.__init__(self, None, name)
        self.dtype, self.shape = dtype, shape

    def eval(self):
        env = _STATE["env"]
        if self not in env["feed"]:
            raise ValueError("You must feed a value for placeholder tensor %r" % self.name)
        a = env["feed"][self]
        return _farr(a) if self.dtype is float32 else _iarr(a)

    __hash__ = Tensor.__hash__
    __eq__ = Tensor.__eq__


class Variable(Tensor):
    def __init__(self, name, init_fn):
        Tensor.__init__(self, None, name)
        self._init_fn = init_fn

    def eval(self):
        store = _STATE["variables"]
        if self.name not in store:
            store[self.name] = np.asarray(self._init_fn())
        if _torch_mode():
            t = _T()
            if not isinstance(store[self.name], t.Tensor):      # leaf tensor: gradients flow to the variable
                store[self.name] = t.tensor(np.asarray(store[self.name], dtype=np.float64), requires_grad=True)
            return store[self.name]
        return np.asarray(store[self.name]).astype(_STATE["dtype"])

    __hash__ = Tensor.__hash__
    __eq__ = Tensor.__eq__


def _v(x):
    return x.eval() if isinstance(x, Tensor) else x


def _i(x):
    return None if x is None else int(_v(x))


def _const(value):
    return Tensor(lambda: value)


# ----------------------------------------------------------------------------------------
# scopes / variables
# ----------------------------------------------------------------------------------------
@contextlib.contextmanager
def variable_scope(name, *a, **k):
    _STATE["scope"].append(name)
    try:
        yield
    finally:
        _STATE["scope"].pop()


@contextlib.contextmanager
def control_dependencies(deps):
    yield


def _scoped(name):
    return "/".join(_STATE["scope"] + [name])


def _make_variable(name, init_fn, trainable=True):
    full = _scoped(name)
    for v in _STATE["trainable"]:
        if v.name == full:
            return v                      # TF: variable reuse inside the same scope (while_loop re-entry)
    v = Variable(full, init_fn)
    if trainable:
        _STATE["trainable"].append(v)
    return v


def get_variable(name=None, initializer=None, dtype=None, shape=None, **k):
    init = initializer
    return _make_variable(name, (lambda: _v(init)) if isinstance(init, Tensor) else (lambda: init(shape)))


def trainable_variables():
    return list(_STATE["trainable"])


def zeros_initializer():
    return lambda shape: np.zeros(shape)


def _xavier_initializer():
    # TF: contrib.layers.xavier_initializer(uniform=True) = U(+-sqrt(6/(fan_in+fan_out))); a 1-D shape
    # (n,) gives fan_in = fan_out = n
    def init(shape):
        shape = tuple(int(s) for s in shape)
        fan_in, fan_out = (shape[0], shape[0]) if len(shape) == 1 else (shape[0], shape[1])
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        return _STATE["init_rng"].uniform(-lim, lim, size=shape)
    return init


def random_normal(shape, **k):
    return Tensor(lambda: _STATE["init_rng"].normal(size=tuple(shape)))


# ----------------------------------------------------------------------------------------
# ops
# ----------------------------------------------------------------------------------------
def placeholder(dtype, shape=None, name=None):
    return Placeholder(dtype, shape, name)


def shape(x):
    return Tensor(lambda: np.array(_shape_of(_v(x)), dtype=np.int64))


def zeros_like(x, dtype=None):
    return Tensor(lambda: _v(x) * 0)


def ones_like(x):
    return Tensor(lambda: _v(x) * 0 + 1)


def matmul(a, b, adjoint_a=False, **k):
    return Tensor(lambda: (_v(a).T if adjoint_a else _v(a)) @ _v(b))


def concat(values, axis=0):
    return Tensor(lambda: _cat([_v(t) for t in values], axis))


def less(a, b):
    return Tensor(lambda: _v(a) < _v(b))


def tile(x, multiples):
    def ev():
        a, reps = _v(x), [int(_v(m)) for m in multiples]
        return a.repeat(*reps) if _is_t(a) else np.tile(a, reps)
    return Tensor(ev)


def div(a, b):
    return Tensor(lambda: _v(a) / _v(b))


def sqrt(x):
    return Tensor(lambda: _fn("sqrt", _v(x)))


def cast(x, dtype):
    if dtype is float32:
        return Tensor(lambda: _farr(_v(x)))
    return Tensor(lambda: _iarr(_v(x)))


def reshape(x, shp):
    def ev():
        a = _v(x)
        return a.reshape(tuple(shp)) if _is_t(a) else np.reshape(a, shp)
    return Tensor(ev)


def reduce_mean(x, **k):
    return Tensor(lambda: _mean(_v(x)))


def reduce_sum(x, **k):
    return Tensor(lambda: _sum(_v(x)))


def sigmoid(x):
    return Tensor(lambda: _sigmoid_v(_v(x)))


def multiply(a, b):
    return Tensor(lambda: _v(a) * _v(b))


def equal(a, b):
    return Tensor(lambda: _v(a) == _v(b))


def not_equal(a, b):
    return Tensor(lambda: _v(a) != _v(b))


def round(x):  # noqa: A001  (TF: round half to even, like numpy)
    return Tensor(lambda: _fn("round", _v(x)))


def add_n(xs):
    return Tensor(lambda: sum(_v(t) for t in xs))


def assert_equal(a, b, data=None, message=None, **k):
    def ev():
        va, vb = _v(a), _v(b)
        va = va.detach().numpy() if _is_t(va) else va
        vb = vb.detach().numpy() if _is_t(vb) else vb
        if not np.all(np.asarray(va) == np.asarray(vb)):
            raise errors.InvalidArgumentError(message)
        return True
    t = Tensor(ev)
    _STATE.setdefault("assertions", []).append(t)
    return t


def gradients(ys, xs, **k):
    """TF: tf.gradients(ys, xs) -- here torch.autograd.grad over the graph the reference built (torch backend)."""
    xs = list(xs)

    def all_grads():
        if not _torch_mode():
            raise NotImplementedError("tf.gradients needs tf1_shim.reset(backend='torch')")
        y = _v(ys)
        leaves = [_v(x) for x in xs]
        gs = _T().autograd.grad(y, leaves, allow_unused=True)
        out = [g if g is not None else leaves[i] * 0 for i, g in enumerate(gs)]
        _STATE["last_grads"] = {x.name: g.detach().numpy().copy() for x, g in zip(xs, out)}
        return out
    whole = Tensor(all_grads)
    return [Tensor(lambda i=i: _v(whole)[i]) for i in range(len(xs))]


def clip_by_global_norm(grads, clip):
    """TF: global_norm = sqrt(sum_i ||g_i||^2); every g_i is scaled by clip / max(global_norm, clip)."""
    grads = list(grads)

    def norm():
        return _fn("sqrt", sum(_sum(_fn("square", _v(g))) for g in grads))
    gn = Tensor(norm)

    def scaled(i):
        n = _v(gn)
        nv = float(n.detach()) if _is_t(n) else float(n)
        _STATE["last_global_norm"] = nv
        return _v(grads[i]) * (float(_v(clip)) / max(nv, float(_v(clip))))
    return [Tensor(lambda i=i: scaled(i)) for i in range(len(grads))], gn


def while_loop(cond, body, loop_vars, **k):
    scope_at_build = list(_STATE["scope"])      # TF traces the body here, under the current scopes
    try:
        # TF builds the loop body ONCE at graph-construction time, which is when its variables come into
        # existence (model.py:166 lists tf.trainable_variables() right after); the result is discarded
        body(*loop_vars)
    except Exception:
        pass                                     # bodies that need concrete values (TensorArray.write) create no variables

    def ev():
        saved, _STATE["scope"] = _STATE["scope"], list(scope_at_build)
        try:
            return run()
        finally:
            _STATE["scope"] = saved

    def run():
        def concretise(v):
            if isinstance(v, Tensor):
                return _const(_v(v))
            if isinstance(v, dict):
                return {kk: concretise(vv) for kk, vv in v.items()}
            if isinstance(v, LSTMStateTuple):
                return LSTMStateTuple(c=concretise(v.c), h=concretise(v.h))
            if isinstance(v, TensorArray):
                return v
            return _const(v)
        cur = [concretise(v) for v in loop_vars]
        while bool(_v(cond(*cur))):
            cur = [concretise(v) for v in body(*cur)]
        return cur
    whole = Tensor(ev)

    def pick(i, template):
        if isinstance(template, dict):
            return {kk: pick_sub(i, lambda r, kk=kk: r[kk], vv) for kk, vv in template.items()}
        if isinstance(template, TensorArray):
            return _LoopArray(lambda: _v(whole)[i])
        return Tensor(lambda: _v(_v(whole)[i]))

    def pick_sub(i, getter, template):
        if isinstance(template, LSTMStateTuple):
            return LSTMStateTuple(c=Tensor(lambda: _v(getter(_v(whole)[i]).c)),
                                  h=Tensor(lambda: _v(getter(_v(whole)[i]).h)))
        return Tensor(lambda: _v(getter(_v(whole)[i])))
    return [pick(i, t) for i, t in enumerate(loop_vars)]


class TensorArray(object):
    def __init__(self, size=None, dtype=None, items=None):
        self.items = dict(items or {})

    def write(self, i, value):
        new = TensorArray(items=self.items)
        new.items[int(_v(i))] = _v(value)
        return new

    def stack(self):
        def ev():
            vals = [self.items[k] for k in sorted(self.items)]
            return _T().stack(vals) if _torch_mode() else np.array(vals)
        return Tensor(ev)


class _LoopArray(object):
    def __init__(self, getter):
        self._getter = getter

    def stack(self):
        return Tensor(lambda: _v(self._getter().stack()))


class LSTMStateTuple(object):
    def __init__(self, c=None, h=None):
        self.c, self.h = c, h

    def __iter__(self):
        return iter((self.c, self.h))


# ----------------------------------------------------------------------------------------
# layers / cells (TF-internal arithmetic, restated)
# ----------------------------------------------------------------------------------------
def _relu(x):
    return Tensor(lambda: _relu_v(_v(x)))


class _Dense(object):
    """TF: tf.layers.Dense -- kernel [in, units], bias [units]; outputs = act(x @ kernel + bias).
    Variables are created on the first call under <current scope>/<layer name>/{kernel,bias}."""

    def __init__(self, units, activation=None, use_bias=True, kernel_initializer=None, bias_initializer=None,
                 name=None, **k):
        self.units, self.activation, self.use_bias, self.name = int(units), activation, use_bias, name
        self.kernel_initializer = kernel_initializer or _xavier_initializer()
        self.bias_initializer = bias_initializer or zeros_initializer()
        self.kernel = self.bias = None

    def __call__(self, x):
        if self.kernel is None:
            with variable_scope(self.name):
                ki, bi, units = self.kernel_initializer, self.bias_initializer, self.units
                self.kernel = _make_variable("kernel", lambda: ki((_shape_of(_v(x))[1], units)))
                self.bias = _make_variable("bias", lambda: bi((units,)))
        kern, bias, act = self.kernel, self.bias, self.activation
        y = Tensor(lambda: _v(x) @ _v(kern) + (_v(bias) if self.use_bias else 0))
        return act(y) if act is not None else y


def _layer_norm(u, gamma, beta):
    # TF: contrib.layers.layer_norm(begin_norm_axis=1): nn.moments + nn.batch_normalization, eps 1e-12
    mean = _mean(u, axis=1, keepdims=True)
    var = _mean(_fn("square", u - mean), axis=1, keepdims=True)
    inv = (1.0 / _fn("sqrt", var + 1e-12)) * gamma
    return u * inv + (beta - mean * inv)


class _LayerNormBasicLSTMCell(object):
    """TF: contrib.rnn.LayerNormBasicLSTMCell(num_units, forget_bias=1.0, activation, layer_norm=True,
    dropout_keep_prob=1.0).  call(): args = concat([inputs, h], 1); concat = args @ kernel (no bias when
    layer_norm); i, j, f, o = split(concat, 4, 1); each gate layer-normalised under scopes input /
    transform / forget / output; g = activation(j); new_c = c*sigmoid(f + forget_bias) + sigmoid(i)*g;
    new_c = LN(new_c, 'state'); new_h = activation(new_c)*sigmoid(o); returns (new_h, (new_c, new_h)).
    Variables live under <scope>/layer_norm_basic_lstm_cell/."""

    def __init__(self, num_units, activation=None, forget_bias=1.0, **k):
        self.num_units, self.activation, self.forget_bias = int(num_units), activation, forget_bias
        self.vars = None

    def __call__(self, inputs=None, state=None):
        d = self.num_units
        with variable_scope("layer_norm_basic_lstm_cell"):
            xavier = _xavier_initializer()
            vs = {"kernel": _make_variable("kernel", lambda: xavier((_shape_of(_v(inputs))[1] + d, 4 * d)))}
            for g in ("input", "transform", "forget", "output", "state"):
                with variable_scope(g):
                    vs[g + "/gamma"] = _make_variable("gamma", lambda: np.ones(d))
                    vs[g + "/beta"] = _make_variable("beta", lambda: np.zeros(d))
        act = self.activation

        def ev():
            c, h = _v(state.c), _v(state.h)
            z = _cat([_v(inputs), h], 1) @ _v(vs["kernel"])
            i, j, f, o = z[:, :d], z[:, d:2 * d], z[:, 2 * d:3 * d], z[:, 3 * d:]
            ln = lambda u, s: _layer_norm(u, _v(vs[s + "/gamma"]), _v(vs[s + "/beta"]))
            i, j, f, o = ln(i, "input"), ln(j, "transform"), ln(f, "forget"), ln(o, "output")
            g = _v(act(_const(j)))
            sg = lambda a: _v(sigmoid(_const(a)))
            new_c = c * sg(f + self.forget_bias) + sg(i) * g
            new_c = ln(new_c, "state")
            new_h = _v(act(_const(new_c))) * sg(o)
            return new_c, new_h
        both = Tensor(ev)
        new_c, new_h = Tensor(lambda: _v(both)[0]), Tensor(lambda: _v(both)[1])
        return new_h, LSTMStateTuple(c=new_c, h=new_h)


def _sigmoid_cross_entropy_with_logits(labels=None, logits=None):
    # TF: max(x, 0) - x*z + log(1 + exp(-|x|))
    return Tensor(lambda: _relu_v(_v(logits)) - _v(logits) * _v(labels)
                  + _fn("log1p", _fn("exp", -_fn("abs", _v(logits)))))


class _Adam(object):
    """TF: tf.train.AdamOptimizer(learning_rate, beta1=0.9, beta2=0.999, epsilon=1e-8)._apply_dense:
    lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t); m = beta1 m + (1-beta1) g; v = beta2 v + (1-beta2) g^2;
    var -= lr_t * m / (sqrt(v) + epsilon).  The update is committed after every fetch of the same
    Session.run has been evaluated, so loss / predictions fetched alongside are pre-update values."""

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, **k):
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta1, beta2, epsilon

    def apply_gradients(self, gv):
        gv = list(gv)

        def ev():
            st = _STATE["adam"]
            grads = [(_v(g), v) for g, v in gv]
            st["t"] = st.get("t", 0) + 1
            t = st["t"]
            lr_t = self.lr * np.sqrt(1.0 - self.b2 ** t) / (1.0 - self.b1 ** t)
            updates = {}
            for g, var in grads:
                g = g.detach().numpy() if _is_t(g) else np.asarray(g)
                m = st.setdefault("m", {}).get(var.name, np.zeros_like(g))
                vv = st.setdefault("v", {}).get(var.name, np.zeros_like(g))
                m = self.b1 * m + (1 - self.b1) * g
                vv = self.b2 * vv + (1 - self.b2) * g * g
                st["m"][var.name], st["v"][var.name] = m, vv
                cur = _v(var)
                cur = cur.detach().numpy() if _is_t(cur) else np.asarray(cur)
                updates[var.name] = cur - lr_t * m / (np.sqrt(vv) + self.eps)
            _STATE["env"].setdefault("post", []).append(lambda: set_variables(updates))
            return None
        return Tensor(ev)


class errors(object):
    class InvalidArgumentError(Exception):
        pass


class Session(object):
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def run(self, fetches, feed_dict=None):
        _STATE["env"] = {"feed": dict(feed_dict or {}), "cache": {}}
        try:
            for t in _STATE.get("assertions", []):
                t.eval()

            def ev(f):
                if isinstance(f, (list, tuple)):
                    return [ev(x) for x in f]
                if isinstance(f, dict):
                    return {k: ev(v) for k, v in f.items()}
                if isinstance(f, LSTMStateTuple):
                    return LSTMStateTuple(c=ev(f.c), h=ev(f.h))
                return _v(f)
            out = ev(fetches)
            for commit in _STATE["env"].get("post", []):
                commit()
            return _to_numpy(out)
        finally:
            _STATE["env"] = None


def _to_numpy(x):
    if isinstance(x, list):
        return [_to_numpy(v) for v in x]
    if isinstance(x, dict):
        return {k: _to_numpy(v) for k, v in x.items()}
    if isinstance(x, LSTMStateTuple):
        return LSTMStateTuple(c=_to_numpy(x.c), h=_to_numpy(x.h))
    if _torch_mode() and isinstance(x, _T().Tensor):
        return x.detach().numpy()
    return x


def global_variables_initializer():
    return _const(None)


def set_variables(values):
    """name -> array; names are the TF variable names the reference's scopes produced."""
    _STATE["variables"].update({k: np.asarray(v) for k, v in values.items()})


def get_variables():
    out = {}
    for k, v in _STATE["variables"].items():
        out[k] = v.detach().numpy().copy() if _torch_mode() and isinstance(v, _T().Tensor) else np.asarray(v).copy()
    return out


def last_gradients():
    """name -> d(loss + l2 * vars_cost)/d(variable) of the last evaluated tf.gradients (before clipping)."""
    return dict(_STATE.get("last_grads", {}))


def last_global_norm():
    return _STATE.get("last_global_norm")


def optimizer_slots():
    st = _STATE.get("adam", {})
    return {"step": st.get("t", 0), "m": dict(st.get("m", {})), "v": dict(st.get("v", {}))}


def variable_names():
    return [v.name for v in _STATE["trainable"]]


def install():
    """Registers this module as ``tensorflow`` (with the contrib / nn / layers / train sub-namespaces)."""
    me = sys.modules[__name__]
    contrib = types.SimpleNamespace(
        layers=types.SimpleNamespace(xavier_initializer=_xavier_initializer),
        rnn=types.SimpleNamespace(LayerNormBasicLSTMCell=_LayerNormBasicLSTMCell, LSTMStateTuple=LSTMStateTuple))
    me.contrib = contrib
    me.nn = types.SimpleNamespace(relu=_relu, sigmoid_cross_entropy_with_logits=_sigmoid_cross_entropy_with_logits,
                                  l2_loss=lambda v: Tensor(lambda: _sum(_fn("square", _v(v))) / 2))
    me.layers = types.SimpleNamespace(Dense=_Dense)
    me.train = types.SimpleNamespace(AdamOptimizer=_Adam)
    sys.modules["tensorflow"] = me
    return me
