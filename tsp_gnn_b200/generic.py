"""Generic CUDA building blocks (include/tspgnn.h, "generic building blocks"): a dense layer, the product with
an arbitrary adjacency matrix and a LayerNormBasicLSTMCell of any size, on fp32 CUDA tensors.

They execute what the fused TSP kernels do not cover: ``Mlp.__call__`` (mlp.py:57-63) and ``GraphNN``
topologies other than build_network's (graphnn.py:142-173).  PyTorch only holds the device memory; every
arithmetic step is a libtspgnn kernel.  There is no CPU path: CPU tensors are rejected.
"""
import ctypes

import numpy as np

from . import _lib

ACTIVATIONS = {None: 0, "none": 0, "linear": 0, "relu": 1, "tanh": 2, "sigmoid": 3}


def _act(code):
    if code not in ACTIVATIONS:
        raise ValueError("unknown activation %r (known: relu, tanh, sigmoid, None)" % (code,))
    return ACTIVATIONS[code]


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _cuda_f32(t, what):
    import torch
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError("%s must be a CUDA tensor (the generic path has no CPU implementation)" % what)
    return t.contiguous().float()


def _stream(t):
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def dense(x, kernel, bias=None, activation=None):
    """tf.layers.Dense: act(x . kernel + bias), kernel [in, out]."""
    import torch
    x, kernel = _cuda_f32(x, "x"), _cuda_f32(kernel, "kernel")
    bias = _cuda_f32(bias, "bias") if bias is not None else None
    if x.dim() != 2 or kernel.dim() != 2 or x.shape[1] != kernel.shape[0]:
        raise ValueError("dense: x %r does not match kernel %r" % (tuple(x.shape), tuple(kernel.shape)))
    y = torch.empty(x.shape[0], kernel.shape[1], dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib.tspgnn_dense_forward(x.device.index, _p(x), x.shape[0], x.shape[1], _p(kernel), _p(bias),
                                             kernel.shape[1], _act(activation), _p(y), _stream(x)))
    return y


class CooMatrix(object):
    """Stored entries of an adjacency matrix on the device (built once per fed matrix)."""

    def __init__(self, dense_matrix, device):
        import torch
        M = np.asarray(dense_matrix.detach().cpu().numpy() if isinstance(dense_matrix, torch.Tensor) else dense_matrix)
        if M.ndim != 2:
            raise ValueError("adjacency matrices are 2-D")
        r, c = np.nonzero(M)
        self.shape = M.shape
        self.rows = torch.from_numpy(r.astype(np.int32)).to(device)
        self.cols = torch.from_numpy(c.astype(np.int32)).to(device)
        vals = M[r, c].astype(np.float32)
        self.vals = None if np.all(vals == 1.0) else torch.from_numpy(vals).to(device)
        self.nnz = int(r.shape[0])

    @classmethod
    def from_entries(cls, rows, cols, vals, shape, device):
        """The same object from the stored entries themselves (int rows / cols, vals None = all ones), for
        matrices that are never held densely (the [sumE, sumV] incidence matrix of a large batch)."""
        import torch
        self = cls.__new__(cls)
        self.shape = (int(shape[0]), int(shape[1]))
        self.rows = torch.from_numpy(np.ascontiguousarray(rows, dtype=np.int32)).to(device)
        self.cols = torch.from_numpy(np.ascontiguousarray(cols, dtype=np.int32)).to(device)
        self.vals = None if vals is None else torch.from_numpy(np.ascontiguousarray(vals, dtype=np.float32)).to(device)
        self.nnz = int(self.rows.shape[0])
        return self

    def matmul(self, y, transpose=False):
        """tf.matmul(M, y, adjoint_a=transpose)."""
        import torch
        y = _cuda_f32(y, "y")
        n_in, n_out = (self.shape[0], self.shape[1]) if transpose else (self.shape[1], self.shape[0])
        if y.shape[0] != n_in:
            raise ValueError("matmul: matrix %r (transpose=%r) does not match y %r" % (self.shape, transpose, tuple(y.shape)))
        out = torch.empty(n_out, y.shape[1], dtype=torch.float32, device=y.device)
        _lib.check(_lib.lib.tspgnn_matmul_coo(y.device.index, _p(self.rows), _p(self.cols), _p(self.vals), self.nnz,
                                              1 if transpose else 0, _p(y), y.shape[1], n_out, _p(out), _stream(y)))
        return out


def lnlstm(x, c, h, kernel, gamma, beta, activation="relu", forget_bias=1.0):
    """tf.contrib.rnn.LayerNormBasicLSTMCell(num_units, activation)(x, (c, h)) -> (c', h').
    kernel [(in + units), 4 units]; gamma / beta [5, units] in gate order input, transform, forget, output, state."""
    import torch
    x, c, h = _cuda_f32(x, "x"), _cuda_f32(c, "c"), _cuda_f32(h, "h")
    kernel, gamma, beta = _cuda_f32(kernel, "kernel"), _cuda_f32(gamma, "gamma"), _cuda_f32(beta, "beta")
    rows, units, in_dim = h.shape[0], h.shape[1], x.shape[1]
    if kernel.shape != (in_dim + units, 4 * units) or gamma.shape != (5, units) or beta.shape != (5, units):
        raise ValueError("lnlstm: parameter shapes do not match in=%d units=%d" % (in_dim, units))
    xh = torch.cat([x, h], dim=1).contiguous()            # tf.concat([inputs, h], 1) of the cell
    c_out, h_out = torch.empty_like(c), torch.empty_like(h)
    scratch = torch.empty(rows, 4 * units, dtype=torch.float32, device=h.device)
    _lib.check(_lib.lib.tspgnn_lnlstm_forward(h.device.index, _p(xh), in_dim, _p(c), rows, units, _p(kernel), _p(gamma),
                                              _p(beta), _act(activation), float(forget_bias), _p(c_out), _p(h_out),
                                              _p(scratch), _stream(h)))
    return c_out, h_out
