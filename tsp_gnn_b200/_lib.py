"""ctypes binding of libtspgnn.so (include/tspgnn.h).

The product path has no CPU fallback: if the shared library cannot be loaded (or built
with nvcc when missing) importing this module raises, and every engine call raises when
the CUDA call fails.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtspgnn.so")

MODE_SIMT_FP32 = 0
MODE_TC_BF16X3 = 1
MODE_TC_BF16 = 2
MODES = {"simt": MODE_SIMT_FP32, "fp32": MODE_SIMT_FP32, "bf16x3": MODE_TC_BF16X3, "bf16": MODE_TC_BF16}

_c_f32p = ctypes.POINTER(ctypes.c_float)
_c_i32p = ctypes.POINTER(ctypes.c_int32)

# (name, restype, argtypes) -- one entry per symbol declared in include/tspgnn.h
SIGNATURES = [
    ("tspgnn_last_error", ctypes.c_char_p, []),
    ("tspgnn_version", ctypes.c_int, []),
    ("tspgnn_param_count", ctypes.c_int64, [ctypes.c_int]),
    ("tspgnn_create", ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    ("tspgnn_destroy", ctypes.c_int, [ctypes.c_void_p]),
    ("tspgnn_get_mode", ctypes.c_int, [ctypes.c_void_p]),
    ("tspgnn_set_option", ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_double]),
    ("tspgnn_set_params", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]),
    ("tspgnn_plan", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                   ctypes.c_void_p, ctypes.c_void_p]),
    ("tspgnn_forward_host", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    ("tspgnn_forward_device", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    ("tspgnn_init_embeddings", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    ("tspgnn_step", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    ("tspgnn_readout", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    ("tspgnn_get_states", ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_void_p] * 5),
    ("tspgnn_set_states", ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_void_p] * 5),
    ("tspgnn_sum_edges", ctypes.c_int64, [ctypes.c_void_p]),
    ("tspgnn_sum_vertices", ctypes.c_int64, [ctypes.c_void_p]),
    ("tspgnn_launch_count", ctypes.c_int64, [ctypes.c_void_p]),
    ("tspgnn_time_kernel", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                          ctypes.POINTER(ctypes.c_float), ctypes.c_void_p]),
    ("tspgnn_debug_timeline", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
                                             ctypes.c_void_p]),
    ("tspgnn_train_forward", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    ("tspgnn_backward", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p]),
    ("tspgnn_apply_gradients", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_float),
                                              ctypes.c_void_p]),
    ("tspgnn_train_step_host", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.c_void_p,
                                              ctypes.c_void_p, ctypes.c_void_p]),
    ("tspgnn_grad_buffer", ctypes.c_void_p, [ctypes.c_void_p]),
    ("tspgnn_get_params", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]),
    ("tspgnn_set_hyper", ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_float] * 6),
    ("tspgnn_get_optimizer_state", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                  ctypes.POINTER(ctypes.c_int64), ctypes.c_int64]),
    ("tspgnn_set_optimizer_state", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                                  ctypes.c_int64]),
    ("tspgnn_dense_forward", ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    ("tspgnn_matmul_coo", ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                         ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p,
                                         ctypes.c_void_p]),
    ("tspgnn_lnlstm_forward", ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
                                             ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                             ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_void_p]),
    ("tspgnn_dense_ev_to_coo", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int64,
                                              ctypes.c_void_p, ctypes.c_void_p]),
]


class TspGnnError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        from . import build as _build
        _build.build()
    lib = ctypes.CDLL(LIB_PATH)
    for name, restype, argtypes in SIGNATURES:
        fn = getattr(lib, name)      # AttributeError if the library does not export it
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


lib = _load()


def check(rc):
    if rc != 0:
        msg = lib.tspgnn_last_error()
        raise TspGnnError("libtspgnn error %d: %s" % (rc, msg.decode() if msg else "?"))
