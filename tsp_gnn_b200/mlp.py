"""Host-side mirror of the reference's ``Mlp`` building block (mlp.py:3-64).

Inside build_network's model an Mlp is a *description* (layer sizes, activations, variable names) whose
arithmetic runs in the fused CUDA kernels; called on its own (``Mlp(...)(inputs)``) it is evaluated layer by
layer with the generic dense kernel.  Constructor arguments keep the reference's names and meaning.
Activations are given by name ('relu' / 'tanh' / 'sigmoid' / None) instead of TF callables.
"""


class Mlp(object):
    def __init__(self, layer_sizes, output_size=None, activations=None, output_activation=None, use_bias=True,
                 kernel_initializer="xavier", bias_initializer="zeros", kernel_regularizer=None,
                 bias_regularizer=None, activity_regularizer=None, kernel_constraint=None, bias_constraint=None,
                 trainable=True, name=None, name_internal_layers=True):
        layer_sizes = list(layer_sizes)
        # mlp.py:26-28: a single activation is repeated for every layer
        if not isinstance(activations, list):
            activations = [activations for _ in layer_sizes]
        # mlp.py:30-33: optional output layer
        if output_size is not None:
            layer_sizes = layer_sizes + [output_size]
            activations = activations + [output_activation]
        if len(activations) != len(layer_sizes):
            raise ValueError("activations and layer_sizes differ in length")
        self.name = name
        self.use_bias = use_bias
        self.kernel_initializer = kernel_initializer
        self.bias_initializer = bias_initializer
        self.trainable = trainable
        self._params = None
        self._device_params = {}
        self.layers = []
        for i, (size, activation) in enumerate(zip(layer_sizes, activations)):
            # tf.layers.Dense casts float sizes (model.py:34 passes d/8, d/4, d/2)
            internal_name = (name + "_MLP_layer_{}".format(i + 1)) if name_internal_layers else None   # mlp.py:36-38
            self.layers.append({"units": int(size), "activation": activation, "name": internal_name})

    def layer_sizes(self):
        return [l["units"] for l in self.layers]

    def variable_names(self, scope=""):
        out = []
        for l in self.layers:
            out.append(scope + l["name"] + "/kernel")
            if self.use_bias:
                out.append(scope + l["name"] + "/bias")
        return out

    # -- parameters (tf.layers.Dense creates kernel [in, units] / bias [units] on first call) ------------------
    def init_parameters(self, input_size, rng=None):
        """Creates the variables for inputs of width ``input_size`` with the initialisers named at construction
        ('xavier' = tf.contrib.layers.xavier_initializer: uniform +-sqrt(6 / (fan_in + fan_out)); a 1-D bias under
        xavier uses fan_in = fan_out = its length; 'zeros')."""
        import numpy as np
        rng = rng if rng is not None else np.random.RandomState(0)

        def make(kind, shape):
            if kind == "xavier":
                fan_in, fan_out = (shape[0], shape[1]) if len(shape) == 2 else (shape[0], shape[0])
                lim = np.sqrt(6.0 / (fan_in + fan_out))
                return rng.uniform(-lim, lim, size=shape).astype(np.float32)
            if kind == "zeros":
                return np.zeros(shape, dtype=np.float32)
            raise ValueError("unknown initializer %r" % (kind,))
        params, fan_in = {}, int(input_size)
        for l in self.layers:
            params[l["name"] + "/kernel"] = make(self.kernel_initializer, (fan_in, l["units"]))
            if self.use_bias:
                params[l["name"] + "/bias"] = make(self.bias_initializer, (l["units"],))
            fan_in = l["units"]
        self.set_parameters(params)
        return params

    def set_parameters(self, params, scope=""):
        """``params``: {'<scope><layer name>/kernel' | '/bias': array}, e.g. a slice of a checkpoint."""
        self._params = {}
        for l in self.layers:
            k = params[scope + l["name"] + "/kernel"]
            b = params[scope + l["name"] + "/bias"] if self.use_bias else None
            self._params[l["name"]] = (k, b)
        self._device_params = {}

    def _on_device(self, device):
        import numpy as np
        import torch
        if device not in self._device_params:
            self._device_params[device] = {
                name: (torch.from_numpy(np.ascontiguousarray(k, dtype=np.float32)).to(device),
                       torch.from_numpy(np.ascontiguousarray(b, dtype=np.float32)).to(device) if b is not None else None)
                for name, (k, b) in self._params.items()}
        return self._device_params[device]

    def __call__(self, inputs, *args, **kwargs):
        """mlp.py:57-63: feeds ``inputs`` through every layer.  ``inputs`` is a CUDA tensor [rows, in] (a numpy array
        is copied to cuda:0 and the result comes back as numpy); every layer is one tspgnn_dense_forward launch."""
        import numpy as np
        import torch
        from . import generic
        if getattr(self, "_params", None) is None:
            raise RuntimeError("Attempting to use uninitialized variables of Mlp %r: call init_parameters() or "
                               "set_parameters() first" % (self.name,))
        as_numpy = isinstance(inputs, np.ndarray)
        x = torch.from_numpy(np.ascontiguousarray(inputs, dtype=np.float32)).cuda() if as_numpy else inputs
        dev = self._on_device(x.device)
        for l in self.layers:
            k, b = dev[l["name"]]
            x = generic.dense(x, k, b, l["activation"])
        return x.cpu().numpy() if as_numpy else x
