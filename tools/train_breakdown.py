"""Host-side timing of the pieces of the training step (forward with snapshots, reverse pass, optimizer +
operand refresh) at the north-star size.  usage: python tools/train_breakdown.py [T]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from tsp_gnn_b200 import instances as inst, params as P      # noqa: E402
from tsp_gnn_b200.engine import Engine                       # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 32
EV, W, C, y, nv, ne = inst.synth_batch([40] * 128, seed=42)
eng = Engine(64, "bf16x3", 0)
eng.set_params(P.init_params(64, seed=0))
eng.plan(nv, ne, EV.src, EV.dst)
dev = torch.device("cuda", 0)
dW = torch.from_numpy(W.astype(np.float32).reshape(-1)).to(dev)
dC = torch.from_numpy(C.astype(np.float32).reshape(-1)).to(dev)
dy = torch.from_numpy(y.astype(np.float32)).to(dev)
s = eng.stream()


def t(f, n=8):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / n


for _ in range(3):
    eng.train_forward(dW, dC, T)
    loss, g = eng.backward(dy)
    eng.apply_gradients(g)
print("train_forward    %.3f ms" % t(lambda: eng.train_forward(dW, dC, T)))
print("backward         %.3f ms" % t(lambda: eng.backward(dy)))
print("apply_gradients  %.3f ms" % t(lambda: eng.apply_gradients(g)))
print("train_step_host  %.3f ms" % t(lambda: eng.train_step_host(W, C, y, T)))
eng.close()
