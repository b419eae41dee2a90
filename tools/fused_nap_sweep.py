import sys, os
sys.path.insert(0, '.')
import numpy as np, torch
from tsp_gnn_b200 import instances as inst, params as P
from tsp_gnn_b200.engine import Engine
EV, W, C, y, nv, ne = inst.synth_batch([40] * 128, seed=42)
eng = Engine(64, "bf16x3", 0)
eng.set_params(P.init_params(64, seed=0))
eng.set_option("fused", 1)
eng.plan(nv, ne, EV.src, EV.dst)
dW = torch.from_numpy(W.astype(np.float32).reshape(-1)).cuda(); dC = torch.from_numpy(C.astype(np.float32).reshape(-1)).cuda()
for nap in (0, 20, 40, 80, 160, 320):
    eng.set_option("dbg", nap << 8)
    eng.init_embeddings(dW, dC); eng.step(2)
    print("nap %d ns: %.2f us per timestep" % (nap, 1e3 * eng.time_kernel(2, 32)))
eng.close()
