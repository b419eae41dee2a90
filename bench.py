#!/usr/bin/env python
"""Benchmark of the TSP-GNN message-passing hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode bf16x3|bf16|simt]

Workload (config 2 of BASELINE.json, per GPU): 128 synthetic 2-D Euclidean instances of n=40
(complete graphs, 780 edges each: sumE=99,840 edge rows, sumV=5,120 vertex rows), d=64,
32 message-passing timesteps, seeded reference-initialiser parameters.

One "step" = one pass of the hot loop over the batch = 32 timesteps (graphnn.py:175-179).
  value : timesteps/s with inputs and recurrent state resident in HBM (tspgnn_step), CUDA
          events per step, L2 flushed between steps, max over ranks, summed over GPUs.
  e2e   : same metric through the host-buffer call a user of the reference makes
          (sess.run -> tspgnn_plan + tspgnn_forward_host: H2D of the incidence columns, W and
          C, E_init, 32 timesteps, vote read-out, D2H of logits/predictions) per step.
  roofline    : the LayerNorm-LSTM kernel (K1), which carries the recurrent-state traffic.
  cpu_baseline: the oracle's dense-EV fp32 loop (what the reference executes on CPU) on a
                bounded sample, timed on this box's host cores (rank 0, N=1 only).
``--impl reference`` times that CPU path alone and prints the same line with impl=reference.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "message-passing timesteps/sec (batch=128, n=40, d=64)"
UNIT = "timesteps/s"
BATCH, N_CITIES, D, T_STEPS = 128, 40, 64, 32
WORKLOAD = "config2: batch=128 n=40 complete Euclidean, d=64, 32 timesteps"


def algorithmic_bytes(nE, nV, state_bytes=4):
    """bytes one timestep must move = read+write of h and c of every edge and
    vertex row + one int32 column index per non-zero per direction (nnz = 2*sumE each way)."""
    nnz = 2 * nE
    total = state_bytes * D * (4 * nE + 4 * nV) + 4 * (2 * nnz)
    k1 = state_bytes * D * (4 * nE + 4 * nV) + 4 * nnz          # LSTM kernel: h,c in; h,c out; gather ids
    return total, k1


class ClockSampler(object):
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line): started ahead of the
    region, samples are kept if their timestamp falls inside [t0, t1] (host clock)."""

    def __init__(self, index):
        self.index, self.rows, self.proc, self.thr = index, [], None, None

    def start(self):
        q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.thr = threading.Thread(target=self._read, daemon=True)
        self.thr.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def wait_first_sample(self, timeout=5.0):
        t_end = time.time() + timeout
        while self.proc is not None and not self.rows and time.time() < t_end:
            time.sleep(0.02)

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.05)
        self.proc.terminate()
        self.thr.join(timeout=2)
        sm, smax, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, r in self.rows:
            if ts < t0 or ts > t1 + 0.03:
                continue
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
                power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


def make_workload(rank):
    from tsp_gnn_b200 import instances as inst
    from tsp_gnn_b200 import params as P
    EV, W, C, y, nv, ne = inst.synth_batch([N_CITIES] * BATCH, seed=42 + 1000 * rank)
    params = P.init_params(D, seed=0)
    return EV, W.astype(np.float32).reshape(-1), C.astype(np.float32).reshape(-1), y, nv, ne, params


def cpu_loop_timesteps_per_s(params, EV, W, C, nv, ne, n_instances, timesteps, repeats):
    """Times the oracle's dense-EV fp32 loop (graphnn.py:156-160 as executed by the reference
    on CPU) over the first ``n_instances`` instances; returns (timesteps/s, seconds, threads)."""
    from oracle import tspgnn_oracle as orc
    nE = int(np.sum(ne[:n_instances]))
    nV = int(np.sum(nv[:n_instances]))
    P32 = {k: v.astype(np.float32) for k, v in params.items()}
    src, dst = EV.src[:nE].astype(np.int64), EV.dst[:nE].astype(np.int64)
    dense = orc.dense_EV(src, dst, nV, np.float32)
    E_h = orc.mlp(np.stack([W[:nE], C[:nE]], axis=1).astype(np.float32), P32, "E_init_MLP")
    V_h = np.tile(P32["V_init"] / np.sqrt(np.float32(D)), (nV, 1)).astype(np.float32)
    E_c, V_c = np.zeros_like(E_h), np.zeros_like(V_h)
    best = float("inf")
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    from threadpoolctl import threadpool_info, threadpool_limits
    # every host core, also under torchrun (which exports OMP_NUM_THREADS=1)
    with threadpool_limits(limits=ncpu):
        for _ in range(repeats):
            t0 = time.perf_counter()
            orc.message_passing(P32, src, dst, dense, E_c, E_h, V_c, V_h, timesteps)
            best = min(best, time.perf_counter() - t0)
        threads = max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    return timesteps / best, best, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    EV, W, C, y, nv, ne, params = make_workload(0)
    ts = 2                                        # timesteps per bench step (bounded sample)
    for _ in range(max(args.warmup, 0) and 1):
        cpu_loop_timesteps_per_s(params, EV, W, C, nv, ne, BATCH, 1, 1)
    t_total, threads = 0.0, 1
    steps = max(1, min(args.steps, 4))
    for _ in range(steps):
        v, sec, threads = cpu_loop_timesteps_per_s(params, EV, W, C, nv, ne, BATCH, ts, 1)
        t_total += sec
    value = ts * steps / t_total
    sample = "full batch 128 x n=40, dense fp32 EV [99840,5120], %d timesteps per step, %d steps" % (ts, steps)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": 1, "ms_per_step": 1e3 * t_total / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "timesteps_per_step": ts, "note": "oracle port of the reference's "
                       "dense-EV TensorFlow CPU loop (TensorFlow 1.x is not installable here); numpy/BLAS threads"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the hot path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from tsp_gnn_b200.engine import Engine
    from tsp_gnn_b200 import sharding

    EV, W, C, y, nv, ne, params = make_workload(rank)
    nE, nV = int(ne.sum()), int(nv.sum())
    eng = Engine(D, args.mode, local)
    eng.set_params(params)
    eng.plan(nv, ne, EV.src, EV.dst)
    stream = eng.stream()
    dW, dC = torch.from_numpy(W).to(dev), torch.from_numpy(C).to(dev)
    d_logits = torch.empty(BATCH, dtype=torch.float32, device=dev)
    d_preds = torch.empty(BATCH, dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident leg: `value` --------------------------------
    with torch.cuda.stream(stream):
        eng.init_embeddings(dW, dC)
        for _ in range(args.warmup):
            eng.step(T_STEPS)
    stream.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    sampler.wait_first_sample()
    barrier()
    t_clock0 = time.time()
    launches0 = eng.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall0 = time.perf_counter()
    with torch.cuda.stream(stream):
        for a, b in ev:
            flush.fill_(1)                 # evict the recurrent state from L2 (not timed)
            a.record(stream)
            eng.step(T_STEPS)
            b.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = eng.launch_count - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    clocks = sampler.stop(t_clock0, time.time())
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    value = world * T_STEPS * args.steps / (dev_ms * 1e-3)

    # ---------------- per-kernel roofline ---------------------------------------------
    roof = None
    if args.mode != "simt":
        with torch.cuda.stream(stream):
            k1_ms = eng.time_kernel(0, 20)
            k2_ms = eng.time_kernel(1, 20)
        total_b, k1_b = algorithmic_bytes(nE, nV, 4)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        ach = k1_b / (k1_ms * 1e-3) / 1e9
        traffic = None        # DRAM bytes per launch of this kernel from the committed ncu --set full capture
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(
                "tc_lnlstm_kernel<%d>" % (2 if args.mode == "bf16x3" else 1))
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": "tc_lnlstm_kernel", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": traffic, "peak_source": "measured" if peaks else "fallback",
                "algorithmic_bytes_per_launch": k1_b, "kernel_ms": k1_ms, "mlp_kernel_ms": k2_ms,
                "timing": "CUDA events around single launches on the engine's stream, 20 launches each, right after "
                          "the timed region (tspgnn_time_kernel); inside the graph the two kernels cannot be bracketed",
                "step_frac_of_hbm_floor": (total_b / (peak * 1e9)) / (dev_ms * 1e-3 / (T_STEPS * args.steps))}

    # ---------------- end-to-end leg through the host-buffer C-ABI call -----------------
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    hW, hC, hsrc, hdst = pin(W), pin(C), pin(EV.src), pin(EV.dst)
    idx = np.arange(BATCH)
    def e2e_step():
        eng.plan(nv, ne, hsrc, hdst)
        logits, preds = eng.forward_host(hW, hC, T_STEPS)
        if world > 1:      # instance-sharded batch: one all-reduce of the zero-padded logits
            full = torch.zeros(world * BATCH, dtype=torch.float32, device=dev)
            full[rank * BATCH:(rank + 1) * BATCH] = torch.from_numpy(logits).to(dev)
            dist.all_reduce(full)
            return full.cpu().numpy()
        return logits
    for _ in range(max(1, args.warmup)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e = {"value": world * T_STEPS * args.steps / e2e_s, "unit": UNIT,
           "h2d_bytes_per_step": int(hW.nbytes + hC.nbytes + hsrc.nbytes + hdst.nbytes + (BATCH + 1) * 8
                                     + ((2 * nE * 4 + (nV + 1) * 4) if args.mode == "simt" else 0)),
           "d2h_bytes_per_step": int(2 * BATCH * 4), "ms_per_step": 1e3 * e2e_s / args.steps,
           "includes": "tspgnn_plan (incidence upload) + E_init + 32 timesteps + vote read-out + D2H"}
    assert np.all(np.isfinite(out))

    # ---------------- training step (secondary) ---------------------------------
    train = None
    if args.train_steps > 0:
        yf = np.asarray(y, dtype=np.float32)
        eng.set_hyper()                                  # model.py:13-15 defaults
        eng.plan(nv, ne, hsrc, hdst)
        eng.train_step_host(hW, hC, yf, T_STEPS)         # warm-up (allocates snapshots / scratch)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.train_steps):
            tr_loss, _, _ = eng.train_step_host(hW, hC, yf, T_STEPS)
        barrier()
        tr_s = (time.perf_counter() - t0) / args.train_steps
        train = {"ms_per_step": 1e3 * tr_s, "instances_per_s": world * BATCH / tr_s, "loss": tr_loss,
                 "what": "tspgnn_train_step_host: H2D, training forward (32 timesteps, snapshots), reverse pass, "
                         "L2 + clip + Adam, operand refresh, D2H (model.py:157-167); per-GPU batch, no gradient "
                         "all-reduce in this leg"}
        eng.set_params(params)                           # the legs below use the seeded variables again

    # ---------------- CPU baseline beside it (rank 0, N=1) ---------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, sec, threads = cpu_loop_timesteps_per_s(params, EV, W, C, nv, ne, BATCH, 2, 2)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "full batch 128 x n=40 with the dense fp32 EV [99840,5120] like graphnn.py:156-160, "
                         "2 timesteps, best of 2 (%.1f s each)" % sec}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": {"bf16x3": "f32 (tcgen05 bf16 hi/lo split operands x3, fp32 accumulate, fp32 state)",
                          "bf16": "bf16 operands / fp32 accumulate, fp32 c state", "simt": "f32"}[args.mode],
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "per_gpu_batch": BATCH, "mode": args.mode,
                           "timesteps_per_step": T_STEPS, "l2": "flushed between steps (256 MiB fill); within a "
                           "step the 53.7 MB recurrent state is re-used across the 32 timesteps",
                           "wall_ms_per_step_incl_flush": 1e3 * t_wall / args.steps},
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
                "train_step": train}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="bf16x3", choices=["bf16x3", "bf16", "simt"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--train-steps", type=int, default=3, help="training steps timed for the secondary train_step block (0 = skip)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
