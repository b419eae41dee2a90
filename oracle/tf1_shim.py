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
"""
import sys
import types
import contextlib

import numpy as np

float32 = np.float32
int32 = np.int32

_STATE = {"scope": [], "variables": {}, "trainable": [], "dtype": np.float64, "env": None, "init_rng": None}


def reset(dtype=np.float64, seed=0):
    _STATE.update(scope=[], variables={}, trainable=[], assertions=[], dtype=dtype, env=None,
                  init_rng=np.random.RandomState(seed))


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
        a = np.asarray(env["feed"][self])
        if self.dtype is float32:
            return a.astype(_STATE["dtype"])
        return a.astype(np.int64)

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
        return store[self.name].astype(_STATE["dtype"])

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
    return Tensor(lambda: np.array(np.shape(_v(x)), dtype=np.int64))


def zeros_like(x, dtype=None):
    return Tensor(lambda: np.zeros_like(_v(x)))


def ones_like(x):
    return Tensor(lambda: np.ones_like(_v(x)))


def matmul(a, b, adjoint_a=False, **k):
    return Tensor(lambda: (_v(a).T if adjoint_a else _v(a)) @ _v(b))


def concat(values, axis=0):
    return Tensor(lambda: np.concatenate([_v(t) for t in values], axis=axis))


def less(a, b):
    return Tensor(lambda: _v(a) < _v(b))


def tile(x, multiples):
    return Tensor(lambda: np.tile(_v(x), [int(_v(m)) for m in multiples]))


def div(a, b):
    return Tensor(lambda: _v(a) / _v(b))


def sqrt(x):
    return Tensor(lambda: np.sqrt(_v(x)))


def cast(x, dtype):
    if dtype is float32:
        return Tensor(lambda: np.asarray(_v(x)).astype(_STATE["dtype"]))
    return Tensor(lambda: np.asarray(_v(x)).astype(np.int64))


def reshape(x, shp):
    return Tensor(lambda: np.reshape(_v(x), shp))


def reduce_mean(x, **k):
    return Tensor(lambda: np.mean(_v(x)))


def reduce_sum(x, **k):
    return Tensor(lambda: np.sum(_v(x)))


def sigmoid(x):
    def ev():
        a = _v(x)
        e = np.exp(-np.abs(a))
        return np.where(a >= 0, 1.0 / (1.0 + e), e / (1.0 + e))
    return Tensor(ev)


def multiply(a, b):
    return Tensor(lambda: _v(a) * _v(b))


def equal(a, b):
    return Tensor(lambda: _v(a) == _v(b))


def not_equal(a, b):
    return Tensor(lambda: _v(a) != _v(b))


def round(x):  # noqa: A001  (TF: round half to even, like numpy)
    return Tensor(lambda: np.round(_v(x)))


def add_n(xs):
    return Tensor(lambda: sum(_v(t) for t in xs))


def assert_equal(a, b, data=None, message=None, **k):
    def ev():
        if not np.all(np.asarray(_v(a)) == np.asarray(_v(b))):
            raise errors.InvalidArgumentError(message)
        return True
    t = Tensor(ev)
    _STATE.setdefault("assertions", []).append(t)
    return t


def gradients(ys, xs, **k):
    return [Tensor(lambda: (_ for _ in ()).throw(NotImplementedError("tf.gradients is not emulated"))) for _ in xs]


def clip_by_global_norm(grads, clip):
    return grads, None


def while_loop(cond, body, loop_vars, **k):
    scope_at_build = list(_STATE["scope"])      # TF traces the body here, under the current scopes

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
        return Tensor(lambda: np.array([self.items[k] for k in sorted(self.items)]))


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
    return Tensor(lambda: np.maximum(_v(x), 0))


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
                self.kernel = _make_variable("kernel", lambda: ki((np.shape(_v(x))[1], units)))
                self.bias = _make_variable("bias", lambda: bi((units,)))
        kern, bias, act = self.kernel, self.bias, self.activation
        y = Tensor(lambda: _v(x) @ _v(kern) + (_v(bias) if self.use_bias else 0))
        return act(y) if act is not None else y


def _layer_norm(u, gamma, beta):
    # TF: contrib.layers.layer_norm(begin_norm_axis=1): nn.moments + nn.batch_normalization, eps 1e-12
    mean = u.mean(axis=1, keepdims=True)
    var = np.square(u - mean).mean(axis=1, keepdims=True)
    inv = (1.0 / np.sqrt(var + u.dtype.type(1e-12))) * gamma
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
            vs = {"kernel": _make_variable("kernel", lambda: xavier((np.shape(_v(inputs))[1] + d, 4 * d)))}
            for g in ("input", "transform", "forget", "output", "state"):
                with variable_scope(g):
                    vs[g + "/gamma"] = _make_variable("gamma", lambda: np.ones(d))
                    vs[g + "/beta"] = _make_variable("beta", lambda: np.zeros(d))
        act = self.activation

        def ev():
            c, h = _v(state.c), _v(state.h)
            z = np.concatenate([_v(inputs), h], axis=1) @ _v(vs["kernel"])
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
    return Tensor(lambda: np.maximum(_v(logits), 0) - _v(logits) * _v(labels) + np.log1p(np.exp(-np.abs(_v(logits)))))


class _Adam(object):
    def __init__(self, **k):
        pass

    def apply_gradients(self, gv):
        return Tensor(lambda: (_ for _ in ()).throw(NotImplementedError("train_step is not emulated")))


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
            return ev(fetches)
        finally:
            _STATE["env"] = None


def global_variables_initializer():
    return _const(None)


def set_variables(values):
    """name -> array; names are the TF variable names the reference's scopes produced."""
    _STATE["variables"].update({k: np.asarray(v) for k, v in values.items()})


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
                                  l2_loss=lambda v: Tensor(lambda: np.sum(np.square(_v(v))) / 2))
    me.layers = types.SimpleNamespace(Dense=_Dense)
    me.train = types.SimpleNamespace(AdamOptimizer=_Adam)
    sys.modules["tensorflow"] = me
    return me
