"""Prints the clock64() trace of the warp roles of a few CTAs for K1 / K2 at the north-star config."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tsp_gnn_b200 import instances as inst, params as P, _lib
from tsp_gnn_b200.engine import Engine
import torch

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
EV, W, C, y, nv, ne = inst.synth_batch([40] * 128, seed=42)
eng = Engine(64, mode, 0)
eng.set_params(P.init_params(64, seed=0))
eng.plan(nv, ne, EV.src, EV.dst)
dW = torch.from_numpy(W.astype(np.float32).reshape(-1)).cuda(); dC = torch.from_numpy(C.astype(np.float32).reshape(-1)).cuda()
eng.init_embeddings(dW, dC); eng.step(4); eng.stream().synchronize()
which_sel = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [2]
ctas_sel = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1, 70, 147]
for which, name, roles in ((0, "K1 lnlstm", ["epiWG0", "epiWG1", "mma", "producer"]), (1, "K2 mlp", ["chainWG0", "chainWG1", "mma", "-"]),
                           (2, "fused step", ["chainWG0", "chainWG1", "mma/relay", "producer"])):
    if which not in which_sel:
        continue
    buf = np.zeros(148 * 4 * 64 * 8, dtype=np.int64)
    _lib.check(_lib.lib.tspgnn_debug_timeline(eng._h, which, buf.ctypes.data_as(ctypes.c_void_p), buf.size, eng._sptr()))
    tl = buf.reshape(148, 4, 64, 8)
    gbase = tl[tl > 0].min()
    spans = []
    for cta in range(148):
        t = tl[cta]
        if (t > 0).any():
            ntiles = int(((t[0] > 0).any(axis=1)).sum() + ((t[1] > 0).any(axis=1)).sum())
            spans.append((int(t.max() - t[t > 0].min()), cta, ntiles))     # SM clocks have per-SM offsets
    spans.sort(reverse=True)
    print("==", name, "per-CTA (span cycles, cta, tiles): slowest 10 / fastest 4; mean span %.0f" % np.mean([x[0] for x in spans]))
    print("   ", spans[:10], "...", spans[-4:])
    for cta in ctas_sel:
        t = tl[cta]
        base = t[t > 0].min() if (t > 0).any() else 0
        print("==", name, "cta", cta, "span", int(t.max() - base))
        for ri, rn in enumerate(roles):
            for tile in range(8):
                ev = t[ri, tile]
                if (ev > 0).any():
                    print("  %-9s tile %d: %s" % (rn, tile, " ".join("%7d" % (v - base) if v > 0 else "      -" for v in ev)))
eng.close()
