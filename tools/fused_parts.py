"""Where the fused timestep kernel's time goes: launch time with parts of the work switched off
(results are wrong in those runs; timing only)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tsp_gnn_b200 import instances as inst, params as P
from tsp_gnn_b200.engine import Engine
mode = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
EV, W, C, y, nv, ne = inst.synth_batch([40] * 128, seed=42)
eng = Engine(64, mode, 0)
eng.set_params(P.init_params(64, seed=0))
eng.plan(nv, ne, EV.src, EV.dst)
dW = torch.from_numpy(W.astype(np.float32).reshape(-1)).cuda(); dC = torch.from_numpy(C.astype(np.float32).reshape(-1)).cuda()
for name, dbg in (("full", 0), ("no reductions", 1), ("no staging / scatter / store", 2), ("cells only (no MLP chain)", 4)):
    eng.set_option("dbg", dbg)
    eng.init_embeddings(dW, dC); eng.step(2)
    ms = eng.time_kernel(2, 30)
    print("%-32s %.1f us per launch" % (name, 1e3 * ms))
eng.set_option("dbg", 0)
for w in (1.0, 1.3, 1.6, 2.0):
    eng.set_option("v_pair_weight", w)
    eng.init_embeddings(dW, dC); eng.step(2)
    print("v_pair_weight %.1f: %.1f us" % (w, 1e3 * eng.time_kernel(2, 30)))
eng.close()
