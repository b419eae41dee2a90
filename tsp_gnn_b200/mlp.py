"""Host-side mirror of the reference's ``Mlp`` building block (mlp.py:3-64).

Here an Mlp is a *description* (layer sizes, activations, variable names); the arithmetic
runs inside the fused CUDA kernels.  Constructor arguments keep the reference's names and
meaning.  Activations are given by name ('relu' / None) instead of TF callables.
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

    def __call__(self, inputs, *args, **kwargs):
        raise NotImplementedError(
            "Mlp objects describe layers; they are evaluated inside the fused CUDA kernels "
            "(E_init, message MLPs and E_vote of build_network). Stand-alone evaluation is not built.")
