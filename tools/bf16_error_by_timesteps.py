import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from oracle import tspgnn_oracle as orc
from tsp_gnn_b200 import instances as inst
from tsp_gnn_b200.engine import Engine
EV, W, C, y, nv, ne = inst.synth_batch([20] * 16, seed=42)
params = orc.init_params(64, seed=0)
for T in (1,2,4,8,16,32):
    eng = Engine(64, "bf16", 0); eng.set_params(params); eng.plan(nv, ne, EV.src, EV.dst)
    logits, preds = eng.forward_host(W, C, T); st = eng.get_states(); eng.close()
    ref = orc.forward(params, EV.src, EV.dst, W, C, nv, ne, T, dtype=np.float64)
    e = lambda a, b: float((np.abs(a.cpu().numpy()-b)/np.maximum(1,np.abs(b))).max())
    print(T, "pred %.2e" % np.abs(preds-ref["predictions"]).max(), "Eh %.2e Ec %.2e Vh %.2e Vc %.2e" % (e(st["E"][1],ref["E_h"]), e(st["E"][0],ref["E_c"]), e(st["V"][1],ref["V_h"]), e(st["V"][0],ref["V_c"])))
