"""Thin object wrapper over the C ABI: one Engine = one tspgnn_handle on one GPU.

PyTorch is used only as the device-memory / stream container; every compute call goes
through libtspgnn.so.  There is no CPU fallback.
"""
import ctypes
import numpy as np

from . import _lib
from .params import flatten, param_offsets


def _np_ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class Engine(object):
    def __init__(self, d=64, mode="bf16x3", device=0):
        self.d = d
        self.mode = mode
        self.device = int(device)
        h = ctypes.c_void_p()
        _lib.check(_lib.lib.tspgnn_create(d, _lib.MODES[mode], self.device, ctypes.byref(h)))
        self._h = h
        self._stream = None
        self.B = 0
        self.n_edges_total = 0
        self.n_vertices_total = 0

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib.tspgnn_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing ---------------------------------------------------------------
    def stream(self):
        """A dedicated (non-legacy) torch stream so the timestep loop can be graph-captured."""
        if self._stream is None:
            import torch
            self._stream = torch.cuda.Stream(device=self.device)
        return self._stream

    def _sptr(self, stream=None):
        s = stream if stream is not None else self.stream()
        return ctypes.c_void_p(s.cuda_stream)

    def _order_in(self, stream, *tensors):
        """Caller tensors are produced on torch's current stream; the engine launches on its own
        (non-blocking) stream.  Make the engine stream wait for the current one and tell the caching
        allocator that the tensors are in use there, so neither a read-before-write nor a premature
        reuse of their memory can happen."""
        import torch
        s = stream if stream is not None else self.stream()
        cur = torch.cuda.current_stream(self.device)
        if cur != s:
            s.wait_stream(cur)
            for t in tensors:
                if t is not None:
                    t.record_stream(s)
        return s

    def _order_out(self, stream):
        """Results written on the engine stream become visible to work queued later on torch's current stream."""
        import torch
        s = stream if stream is not None else self.stream()
        cur = torch.cuda.current_stream(self.device)
        if cur != s:
            cur.wait_stream(s)

    def set_option(self, name, value):
        """Tuning / diagnosis knobs of include/tspgnn.h (e.g. ``fused`` 0/1)."""
        _lib.check(_lib.lib.tspgnn_set_option(self._h, name.encode(), float(value)))

    # -- parameters / plan --------------------------------------------------------
    def set_params(self, params):
        blob = params if isinstance(params, np.ndarray) else flatten(params, self.d)
        blob = np.ascontiguousarray(blob, dtype=np.float32)
        _lib.check(_lib.lib.tspgnn_set_params(self._h, _np_ptr(blob), blob.size))

    def plan(self, n_vertices, n_edges, edge_src, edge_dst):
        nv = np.ascontiguousarray(n_vertices, dtype=np.int32)
        ne = np.ascontiguousarray(n_edges, dtype=np.int32)
        src = np.ascontiguousarray(edge_src, dtype=np.int32)
        dst = np.ascontiguousarray(edge_dst, dtype=np.int32)
        if nv.shape != ne.shape or nv.ndim != 1:
            raise ValueError("n_vertices and n_edges must be 1-D arrays of equal length")
        if src.shape != dst.shape or src.ndim != 1 or src.shape[0] != int(ne.sum()):
            raise ValueError("edge_src/edge_dst must have sum(n_edges) entries")
        _lib.check(_lib.lib.tspgnn_plan(self._h, nv.shape[0], _np_ptr(nv), _np_ptr(ne), _np_ptr(src), _np_ptr(dst)))
        self.B = int(nv.shape[0])
        self.n_edges_total = int(ne.sum())
        self.n_vertices_total = int(nv.sum())

    # -- forward ------------------------------------------------------------------
    def forward_host(self, W, C, time_steps, stream=None):
        """Host numpy in, host numpy out (H2D + forward + D2H inside the call)."""
        W = np.ascontiguousarray(np.asarray(W, dtype=np.float32).reshape(-1))
        C = np.ascontiguousarray(np.asarray(C, dtype=np.float32).reshape(-1))
        if W.shape[0] != self.n_edges_total or C.shape[0] != self.n_edges_total:
            raise ValueError("W and C must have sum(n_edges)=%d rows" % self.n_edges_total)
        logits = np.empty(self.B, dtype=np.float32)
        preds = np.empty(self.B, dtype=np.float32)
        _lib.check(_lib.lib.tspgnn_forward_host(self._h, _np_ptr(W), _np_ptr(C), int(time_steps), _np_ptr(logits),
                                                _np_ptr(preds), self._sptr(stream)))
        return logits, preds

    def forward_device(self, dW, dC, time_steps, d_logits, d_preds, stream=None):
        self._order_in(stream, dW, dC, d_logits, d_preds)
        _lib.check(_lib.lib.tspgnn_forward_device(self._h, ctypes.c_void_p(dW.data_ptr()), ctypes.c_void_p(dC.data_ptr()),
                                                  int(time_steps), ctypes.c_void_p(d_logits.data_ptr()),
                                                  ctypes.c_void_p(d_preds.data_ptr()), self._sptr(stream)))
        self._order_out(stream)

    def init_embeddings(self, dW, dC, stream=None):
        self._order_in(stream, dW, dC)
        _lib.check(_lib.lib.tspgnn_init_embeddings(self._h, ctypes.c_void_p(dW.data_ptr()),
                                                   ctypes.c_void_p(dC.data_ptr()), self._sptr(stream)))

    def step(self, n_steps, stream=None):
        _lib.check(_lib.lib.tspgnn_step(self._h, int(n_steps), self._sptr(stream)))

    def readout(self, d_logits, d_preds, stream=None):
        self._order_in(stream, d_logits, d_preds)
        _lib.check(_lib.lib.tspgnn_readout(self._h, ctypes.c_void_p(d_logits.data_ptr()),
                                           ctypes.c_void_p(d_preds.data_ptr()), self._sptr(stream)))
        self._order_out(stream)

    def get_states(self, stream=None):
        """{'V': (c,h), 'E': (c,h)} as row-major fp32 torch tensors on the device."""
        import torch
        dev = torch.device("cuda", self.device)
        s = stream if stream is not None else self.stream()
        Vh = torch.empty(self.n_vertices_total, self.d, dtype=torch.float32, device=dev)
        Vc = torch.empty_like(Vh)
        Eh = torch.empty(self.n_edges_total, self.d, dtype=torch.float32, device=dev)
        Ec = torch.empty_like(Eh)
        _lib.check(_lib.lib.tspgnn_get_states(self._h, *[ctypes.c_void_p(t.data_ptr()) for t in (Vh, Vc, Eh, Ec)],
                                              self._sptr(s)))
        s.synchronize()
        return {"V": (Vc, Vh), "E": (Ec, Eh)}

    def set_states(self, Vh=None, Vc=None, Eh=None, Ec=None, stream=None):
        s = self._order_in(stream, Vh, Vc, Eh, Ec)
        ptrs = [ctypes.c_void_p(t.data_ptr()) if t is not None else None for t in (Vh, Vc, Eh, Ec)]
        _lib.check(_lib.lib.tspgnn_set_states(self._h, *ptrs, self._sptr(s)))
        s.synchronize()

    # -- training step (model.py:157-167) -------------------------------------------
    def train_step_host(self, W, C, route_exists, time_steps, stream=None):
        """One ``sess.run([train_step, loss, predictions])`` (train.py:35-42): host numpy in,
        (loss, logits, predictions) of the pre-update variables out; the variables, Adam
        slots and every derived operand image are updated on the device."""
        W = np.ascontiguousarray(np.asarray(W, dtype=np.float32).reshape(-1))
        C = np.ascontiguousarray(np.asarray(C, dtype=np.float32).reshape(-1))
        y = np.ascontiguousarray(np.asarray(route_exists, dtype=np.float32).reshape(-1))
        if W.shape[0] != self.n_edges_total or C.shape[0] != self.n_edges_total:
            raise ValueError("W and C must have sum(n_edges)=%d rows" % self.n_edges_total)
        if y.shape[0] != self.B:
            raise ValueError("route_exists must have one entry per instance (%d)" % self.B)
        loss = ctypes.c_float(0)
        logits = np.empty(self.B, dtype=np.float32)
        preds = np.empty(self.B, dtype=np.float32)
        _lib.check(_lib.lib.tspgnn_train_step_host(self._h, _np_ptr(W), _np_ptr(C), _np_ptr(y), int(time_steps),
                                                   ctypes.byref(loss), _np_ptr(logits), _np_ptr(preds),
                                                   self._sptr(stream)))
        return float(loss.value), logits, preds

    def train_forward(self, dW, dC, time_steps, d_logits=None, d_preds=None, stream=None):
        """Forward pass that keeps the per-timestep state the reverse pass needs (device tensors)."""
        self._order_in(stream, dW, dC, d_logits, d_preds)
        _lib.check(_lib.lib.tspgnn_train_forward(
            self._h, ctypes.c_void_p(dW.data_ptr()), ctypes.c_void_p(dC.data_ptr()), int(time_steps),
            ctypes.c_void_p(d_logits.data_ptr()) if d_logits is not None else None,
            ctypes.c_void_p(d_preds.data_ptr()) if d_preds is not None else None, self._sptr(stream)))
        self._order_out(stream)

    def backward(self, d_route_exists, global_batch=0, stream=None):
        """Loss and flat gradient blob (device tensors) of the last train_forward.  ``global_batch``
        is the divisor of the loss mean; with instances sharded over ranks pass the size of the
        whole batch and sum (all-reduce) the returned blobs."""
        import torch
        dev = torch.device("cuda", self.device)
        s = self._order_in(stream, d_route_exists)
        with torch.cuda.stream(s):
            grads = torch.empty(self.param_count, dtype=torch.float32, device=dev)
            loss = torch.empty(1, dtype=torch.float32, device=dev)
        _lib.check(_lib.lib.tspgnn_backward(self._h, ctypes.c_void_p(d_route_exists.data_ptr()), int(global_batch),
                                            ctypes.c_void_p(grads.data_ptr()), ctypes.c_void_p(loss.data_ptr()),
                                            self._sptr(s)))
        self._order_out(s)
        return loss, grads

    def apply_gradients(self, d_grads, stream=None):
        """L2 term + clip_by_global_norm + Adam (model.py:160-167); returns the global norm."""
        norm = ctypes.c_float(0)
        self._order_in(stream, d_grads)
        _lib.check(_lib.lib.tspgnn_apply_gradients(self._h, ctypes.c_void_p(d_grads.data_ptr()), ctypes.byref(norm),
                                                   self._sptr(stream)))
        return float(norm.value)

    @property
    def param_count(self):
        return int(_lib.lib.tspgnn_param_count(self.d))

    def get_params(self):
        """Current variables as a flat float32 blob (params.unflatten turns it into a dict)."""
        blob = np.empty(self.param_count, dtype=np.float32)
        _lib.check(_lib.lib.tspgnn_get_params(self._h, _np_ptr(blob), blob.size))
        return blob

    def set_hyper(self, learning_rate=2e-5, l2norm_scaling=1e-10, clip_norm=0.65, beta1=0.9, beta2=0.999,
                  epsilon=1e-8):
        _lib.check(_lib.lib.tspgnn_set_hyper(self._h, learning_rate, l2norm_scaling, clip_norm, beta1, beta2, epsilon))

    def get_optimizer_state(self):
        n = self.param_count
        m = np.empty(n, dtype=np.float32)
        v = np.empty(n, dtype=np.float32)
        step = ctypes.c_int64(0)
        _lib.check(_lib.lib.tspgnn_get_optimizer_state(self._h, _np_ptr(m), _np_ptr(v), ctypes.byref(step), n))
        return {"m": m, "v": v, "step": int(step.value)}

    def set_optimizer_state(self, state):
        m = np.ascontiguousarray(state["m"], dtype=np.float32)
        v = np.ascontiguousarray(state["v"], dtype=np.float32)
        _lib.check(_lib.lib.tspgnn_set_optimizer_state(self._h, _np_ptr(m), _np_ptr(v), int(state["step"]), m.size))

    def time_kernel(self, which, iters, stream=None):
        """Mean device time (ms) of one launch of K1 (which=0), K2 (which=1) or the fused timestep kernel (which=2)."""
        ms = ctypes.c_float(0)
        _lib.check(_lib.lib.tspgnn_time_kernel(self._h, int(which), int(iters), ctypes.byref(ms), self._sptr(stream)))
        return float(ms.value)

    @property
    def launch_count(self):
        return int(_lib.lib.tspgnn_launch_count(self._h))


def dense_ev_to_coo(EV):
    """Dense [sumE,sumV] float32/float64 EV -> (edge_src, edge_dst) via the C helper."""
    EV = np.ascontiguousarray(EV)
    if EV.dtype not in (np.float32, np.float64):
        EV = EV.astype(np.float32)
    src = np.empty(EV.shape[0], dtype=np.int32)
    dst = np.empty(EV.shape[0], dtype=np.int32)
    _lib.check(_lib.lib.tspgnn_dense_ev_to_coo(_np_ptr(EV), EV.dtype.itemsize, EV.shape[0], EV.shape[1],
                                               _np_ptr(src), _np_ptr(dst)))
    return src, dst
