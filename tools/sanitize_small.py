"""Small end-to-end run for compute-sanitizer: forward in every mode + one training step, ragged sizes with tail tiles.
usage: compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from tsp_gnn_b200 import instances as inst, params as P      # noqa: E402
from tsp_gnn_b200.engine import Engine                       # noqa: E402

EV, W, C, y, nv, ne = inst.synth_batch([5, 23, 17, 9, 30], seed=3)
params = P.init_params(64, seed=1)
for mode in ("simt", "bf16x3", "bf16"):
    eng = Engine(64, mode, 0)
    eng.set_params(params)
    eng.plan(nv, ne, EV.src, EV.dst)
    logits, preds = eng.forward_host(W, C, 3)
    loss, _, _ = eng.train_step_host(W, C, y, 3)
    st = eng.get_states()
    print(mode, "pred", preds[:2], "loss", loss)
    eng.close()
