"""Host-side timing of the pieces of the end-to-end call (plan, init, loop, readout) at the north-star size."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tsp_gnn_b200 import instances as inst, params as P
from tsp_gnn_b200.engine import Engine

EV, W, C, y, nv, ne = inst.synth_batch([40] * 128, seed=42)
W = W.astype(np.float32).reshape(-1); C = C.astype(np.float32).reshape(-1)
eng = Engine(64, "bf16x3", 0); eng.set_params(P.init_params(64, seed=0))
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
hW, hC, hs, hd = pin(W), pin(C), pin(EV.src), pin(EV.dst)
eng.plan(nv, ne, hs, hd); eng.forward_host(hW, hC, 32)
def t(f, n=50):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return 1e3 * (time.perf_counter() - t0) / n
dW, dC = torch.from_numpy(W).cuda(), torch.from_numpy(C).cuda()
dl = torch.empty(128, device="cuda"); dp = torch.empty(128, device="cuda")
s = eng.stream()
def sync(): s.synchronize()
print("plan                      %.3f ms" % t(lambda: eng.plan(nv, ne, hs, hd)))
print("forward_host (32 steps)   %.3f ms" % t(lambda: eng.forward_host(hW, hC, 32)))
print("forward_host (0 steps)    %.3f ms" % t(lambda: eng.forward_host(hW, hC, 0)))
print("init_embeddings           %.3f ms" % t(lambda: (eng.init_embeddings(dW, dC), sync())))
print("step(32)                  %.3f ms" % t(lambda: (eng.step(32), sync())))
print("readout                   %.3f ms" % t(lambda: (eng.readout(dl, dp), sync())))
