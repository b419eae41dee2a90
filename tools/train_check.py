"""Development aid: per-tensor error report of the CUDA reverse pass against the float64 oracle.
usage: python tools/train_check.py [mode] [T] [sizes...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import tspgnn_oracle as orc            # noqa: E402
from oracle import tspgnn_oracle_grad as og        # noqa: E402
from tsp_gnn_b200 import instances as inst         # noqa: E402
from tsp_gnn_b200 import params as P               # noqa: E402
from tsp_gnn_b200.engine import Engine             # noqa: E402
import torch                                       # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sizes = [int(a) for a in sys.argv[3:]] or [5, 12, 20, 7, 33, 9]
EV, W, C, y, nv, ne = inst.synth_batch(sizes, seed=11)
params = orc.init_params(64, seed=5, perturb_ln=True)
t0 = time.time()
big = int(np.sum(ne)) > 20000
ref = None if big else og.forward_backward(params, EV.src, EV.dst, W, C, nv, ne, y, T)
print("oracle %.1fs, sumE %d sumV %d" % (time.time() - t0, int(np.sum(ne)), int(np.sum(nv))))
eng = Engine(64, mode, 0)
eng.set_params(params)
if os.environ.get("TSPGNN_TRAIN_TC") is not None:
    eng.set_option("train_tc", int(os.environ["TSPGNN_TRAIN_TC"]))
eng.plan(nv, ne, EV.src, EV.dst)
dev = torch.device("cuda", 0)
s = eng.stream()
with torch.cuda.stream(s):
    dW = torch.from_numpy(np.asarray(W, dtype=np.float32).reshape(-1)).to(dev)
    dC = torch.from_numpy(np.asarray(C, dtype=np.float32).reshape(-1)).to(dev)
    dy = torch.from_numpy(np.asarray(y, dtype=np.float32)).to(dev)
    logits = torch.empty(eng.B, dtype=torch.float32, device=dev)
    preds = torch.empty(eng.B, dtype=torch.float32, device=dev)
s.synchronize()
eng.train_forward(dW, dC, T, logits, preds)
loss, grads = eng.backward(dy, 0)
s.synchronize()
if ref is not None:
    print("mode %s T %d  loss cuda %.8f oracle %.8f  max|dlogit| %.2e" %
          (mode, T, float(loss.cpu()[0]), ref["loss"], np.abs(logits.cpu().numpy() - ref["logits"]).max()))
got = P.unflatten(grads.cpu().numpy())
for k, r in (ref["grads"].items() if ref is not None else []):
    err = np.abs(got[k] - r).max()
    scale = np.abs(r).max()
    flag = "" if err <= 2e-3 * scale + 1e-9 else "   <-- MISMATCH"
    print("%-62s err %.3e scale %.3e rel %.2e%s" % (k, err, scale, err / (scale + 1e-30), flag))
# timing of the training step at this size
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn in (("train_forward", lambda: eng.train_forward(dW, dC, T, logits, preds)),
                 ("backward", lambda: eng.backward(dy, 0))):
    fn()
    s.synchronize()
    with torch.cuda.stream(s):
        ev0.record(s)
        fn()
        ev1.record(s)
    s.synchronize()
    print("%s: %.3f ms" % (name, ev0.elapsed_time(ev1)))
t0 = time.time()
for _ in range(3):
    out = eng.train_step_host(W, C, y, T)
print("train_step_host: %.3f ms per step, loss %.6f" % ((time.time() - t0) / 3 * 1e3, out[0]))
eng.close()
