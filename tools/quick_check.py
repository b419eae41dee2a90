"""Fast GPU sanity run used while developing kernels: small forwards in each tensor-core mode
against the float64 oracle, flushing progress as it goes (a hang shows where it stopped)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import tspgnn_oracle as orc
from tsp_gnn_b200 import instances as inst
from tsp_gnn_b200.engine import Engine

def p(*a):
    print(*a); sys.stdout.flush()

modes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["bf16x3", "bf16"]
cases = [([5, 6, 7, 8], 1), ([5, 6, 7, 8], 4), ([20] * 16, 8), ([40] * 128, 4)]
params = orc.init_params(64, seed=7, perturb_ln=True)
for sizes, T in cases:
    EV, W, C, y, nv, ne = inst.synth_batch(sizes, seed=11)
    ref = orc.forward(params, EV.src, EV.dst, W, C, nv, ne, T, dtype=np.float64)
    for mode in modes:
        p("case", len(sizes), "x", sizes[0], "T", T, mode, "...")
        t0 = time.time()
        eng = Engine(64, mode, 0)
        eng.set_params(params)
        if os.environ.get("TSPGNN_FUSED"):
            eng.set_option("fused", int(os.environ["TSPGNN_FUSED"]))
        eng.plan(nv, ne, EV.src, EV.dst)
        logits, preds = eng.forward_host(W, C, T)
        st = eng.get_states()
        eng.close()
        errs = {k: float(np.abs(v.cpu().numpy() - ref[n]).max()) for k, v, n in
                (("Vc", st["V"][0], "V_c"), ("Vh", st["V"][1], "V_h"), ("Ec", st["E"][0], "E_c"), ("Eh", st["E"][1], "E_h"))}
        p("   pred err %.2e" % np.abs(preds - ref["predictions"]).max(), {k: "%.1e" % v for k, v in errs.items()},
          "%.2fs" % (time.time() - t0))
