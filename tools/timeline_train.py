"""clock64() trace of the warp roles of the reverse pass's tcgen05 row GEMM on edge-sized matrices (north-star size).
usage: python tools/timeline_train.py [4|5|6|7] [cta,cta,...]    4 = 64-wide layer, 5 = dz.K^T, 6 = z = [x,h].K,
       7 = tc_layer_reverse_kernel (epilogue: 0 tile start, 1 mask requested, 2 accumulator ready, 3 stores issued;
           mma: 0 start, 1 stage full, 2 accumulator free, 3 issued; producer 0: 0 start, 1 loads issued, 2 stage free, 4 published)
events  epilogue WG: 0 tile start, 1 accumulator ready, 2 stores issued
        mma        : 0 waits for the operand block, 1 block ready, 2 MMAs issued (per k-block)
        producer 0 : 0 block start, 1 loads issued, 2 slot free, 3 first group converted (data arrived), 4 block published"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tsp_gnn_b200 import instances as inst, params as P, _lib      # noqa: E402
from tsp_gnn_b200.engine import Engine                             # noqa: E402

which = int(sys.argv[1]) if len(sys.argv) > 1 else 4
ctas = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 70]
EV, W, C, y, nv, ne = inst.synth_batch([40] * 128, seed=42)
eng = Engine(64, "bf16x3", 0)
eng.set_params(P.init_params(64, seed=0))
eng.plan(nv, ne, EV.src, EV.dst)
buf = np.zeros(148 * 4 * 64 * 8, dtype=np.int64)
_lib.check(_lib.lib.tspgnn_debug_timeline(eng._h, which, buf.ctypes.data_as(ctypes.c_void_p), buf.size, eng._sptr()))
tl = buf.reshape(148, 4, 64, 8)
roles = ["epiWG0", "epiWG1", "mma", "producer0"]
spans = sorted(((int(tl[c].max() - tl[c][tl[c] > 0].min()), c) for c in range(148) if (tl[c] > 0).any()), reverse=True)
print("== row GEMM %d: per-CTA span (cycles, cta): slowest 5 %s fastest 3 %s mean %.0f"
      % (which, spans[:5], spans[-3:], np.mean([x[0] for x in spans])))
for cta in ctas:
    t = tl[cta]
    base = t[t > 0].min()
    print("== cta", cta, "span", int(t.max() - base))
    for ri, rn in enumerate(roles):
        for tile in range(24):
            ev = t[ri, tile]
            if (ev > 0).any():
                print("  %-9s %2d: %s" % (rn, tile, " ".join("%7d" % (v - base) if v > 0 else "      -" for v in ev[:5])))
eng.close()
