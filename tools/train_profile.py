"""Development aid: one training step at config-2 size (T timesteps) for an ncu launch list.
usage: ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/train_profile.py [T]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from tsp_gnn_b200 import instances as inst, params as P      # noqa: E402
from tsp_gnn_b200.engine import Engine                       # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 2
EV, W, C, y, nv, ne = inst.synth_batch([40] * 128, seed=42)
eng = Engine(64, "bf16x3", 0)
eng.set_params(P.init_params(64, seed=0))
eng.plan(nv, ne, EV.src, EV.dst)
for _ in range(2):
    out = eng.train_step_host(W, C, y, T)
print("loss", out[0])
eng.close()
