"""Device-resident timestep throughput of every BASELINE.json config on one GPU (the bench line itself is config 2).
usage: python tools/config_sweep.py  -> one line per config: timesteps/s, instances/s (full forward through host buffers)"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from tsp_gnn_b200 import instances as inst, params as P      # noqa: E402
from tsp_gnn_b200.engine import Engine                       # noqa: E402

T = 32
CONFIGS = [
    ("config1 16 x n=20", "bf16x3", [20] * 16),
    ("config2 128 x n=40", "bf16x3", [40] * 128),
    ("config3 128 x n=40 bf16 embeddings", "bf16", [40] * 128),
    ("config4 512 x n in 20..60 (one GPU holds the whole batch)", "bf16x3", inst.mixed_sizes(512, 20, 60, seed=11)),
    ("config4 per-GPU shard: 64 x n in 20..60", "bf16x3", inst.mixed_sizes(64, 20, 60, seed=12)),
    ("config5 32 x n=80", "bf16x3", [80] * 32),
    ("config5 32 x n=160", "bf16x3", [160] * 32),
    ("config5 32 x n=320", "bf16x3", [320] * 32),
]
params = P.init_params(64, seed=0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, mode, sizes in CONFIGS:
    insts = inst.synth_instances(sizes, seed=42, two_opt_sweeps=0)
    EV, W, C, y, nv, ne = inst.create_batch(insts)
    W = W.astype(np.float32).reshape(-1)
    C = C.astype(np.float32).reshape(-1)
    eng = Engine(64, mode, 0)
    eng.set_params(params)
    eng.plan(nv, ne, EV.src, EV.dst)
    s = eng.stream()
    dW, dC = torch.from_numpy(W).cuda(), torch.from_numpy(C).cuda()
    with torch.cuda.stream(s):
        eng.init_embeddings(dW, dC)
        for _ in range(3):
            eng.step(T)
    s.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    with torch.cuda.stream(s):
        for a, b in ev:
            flush.fill_(1)
            a.record(s)
            eng.step(T)
            b.record(s)
    s.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / len(ev)
    eng.forward_host(W, C, T)
    t0 = time.perf_counter()
    for _ in range(5):
        logits, preds = eng.forward_host(W, C, T)
    fwd = (time.perf_counter() - t0) / 5
    nE, nV = int(ne.sum()), int(nv.sum())
    bytes_step = 4 * 64 * (4 * nE + 4 * nV) * (1 if mode != "bf16" else 1) + 4 * 4 * nE
    print("%-58s sumE %8d  %9.0f timesteps/s  %7.1f us/timestep  %5.2f TB/s algorithmic  forward %8.2f ms = %8.0f instances/s"
          % (name, nE, T / (ms * 1e-3), ms * 1e3 / T, bytes_step / (ms * 1e-3 / T) / 1e12, fwd * 1e3, len(sizes) / fwd), flush=True)
    eng.close()
