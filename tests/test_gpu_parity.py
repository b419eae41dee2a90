"""Parity of the CUDA path (through the C ABI) against the oracle.  Needs a B200.

Tolerances (north star: outputs within 1e-4 of the fp32 reference path).  Predictions are
gated on absolute error; recurrent states on |got-ref| <= tol * max(1, |ref|) (LayerNorm'd
cell states reach |c| ~ 4, so a pure absolute bound would be tighter than fp32-relative 1e-4):
  simt   fp32 FMA                       predictions 1e-5, states 1e-4
  bf16x3 tcgen05, hi/lo split operands  predictions 1e-4, states 1e-4
  bf16   tcgen05, bf16 embeddings       predictions 5e-3, states 1e-1 (BASELINE config 3:
         reported with its measured error, not gated at 1e-4)
"""
import os
import sys

import numpy as np
import pytest

from oracle import tspgnn_oracle as orc
from tsp_gnn_b200 import instances as inst
from tsp_gnn_b200 import params as P

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import make_golden  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_forward.npz"))
MODES = ["simt", "bf16x3", "bf16"]
TOL_PRED = {"simt": 1e-5, "bf16x3": 1e-4, "bf16": 5e-3}
TOL_STATE = {"simt": 1e-4, "bf16x3": 1e-4, "bf16": 1e-1}


def state_err(got, ref):
    """max over elements of |got-ref| / max(1, |ref|)."""
    return float((np.abs(got - ref) / np.maximum(1.0, np.abs(ref))).max())


def make_engine(mode, params):
    from tsp_gnn_b200.engine import Engine
    eng = Engine(64, mode, 0)
    eng.set_params(params)
    return eng


def run_engine(mode, params, EV, W, C, nv, ne, T):
    eng = make_engine(mode, params)
    eng.plan(nv, ne, EV.src, EV.dst)
    logits, preds = eng.forward_host(W, C, T)
    st = eng.get_states()
    out = dict(logits=logits, predictions=preds, V_c=st["V"][0].cpu().numpy(), V_h=st["V"][1].cpu().numpy(),
               E_c=st["E"][0].cpu().numpy(), E_h=st["E"][1].cpu().numpy())
    eng.close()
    return out


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_forward_matches_oracle_and_golden(mode, name):
    sizes, iseed, pseed, perturb, T, conn = make_golden.CASES[name]
    EV, W, C, y, nv, ne = inst.synth_batch(sizes, seed=iseed, connectivity=conn)
    params = orc.init_params(64, seed=pseed, perturb_ln=perturb)
    got = run_engine(mode, params, EV, W, C, nv, ne, T)
    ref = orc.forward(params, EV.src, EV.dst, W, C, nv, ne, T, dtype=np.float64)
    err = {k: state_err(got[k], ref[k]) for k in ("E_h", "E_c", "V_h", "V_c")}
    err["predictions"] = float(np.abs(got["predictions"] - ref["predictions"]).max())
    print(mode, name, err)
    assert err["predictions"] <= TOL_PRED[mode], err
    for k in ("E_h", "E_c", "V_h", "V_c"):
        assert err[k] <= TOL_STATE[mode], err
    assert np.abs(got["predictions"] - GOLD[name + "/predictions"]).max() <= TOL_PRED[mode]
    assert state_err(got["E_c"][:32], GOLD[name + "/E_c_head"]) <= TOL_STATE[mode]


# Discriminating parity case: predictions spread over (0.15, 0.85) (with the reference initialisers all 16
# predictions agree to four digits, so a 1e-4 gate on them alone says little).  The scaled kernels also make
# the loop less contractive: the fp32 dense oracle itself moves 1.4e-6 (x2.5) / 1.1e-5 (x3) away from float64.
# bf16x3 carries h with 16 mantissa bits (bf16 hi + lo) and drops the lo.lo products, i.e. about 100x the
# rounding noise of fp32: measured on B200 1.2e-4 (x2.5) where fp32 SIMT gives 2e-6 -- the 1e-4 north-star
# tolerance holds with 100x margin on reference-initialiser parameters (8e-7) and is just exceeded here.
# The single-bf16 mode (config 3) is reported, not gated, on these sets: 7e-2 measured at x2.5.
TOL_SPREAD_PRED = {"spread25_16x20": {"simt": 2e-5, "bf16x3": 3e-4, "bf16": 0.5},
                   "spread30_16x20": {"simt": 1e-4, "bf16x3": 1e-3, "bf16": 0.5}}


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("name", sorted(make_golden.SPREAD_CASES))
def test_spread_predictions_match_oracle_and_golden(mode, name):
    EV, W, C, y, nv, ne, params, T = make_golden.spread_case_inputs(name)
    got = run_engine(mode, params, EV, W, C, nv, ne, T)
    ref = orc.forward(params, EV.src, EV.dst, W, C, nv, ne, T, dtype=np.float64)
    assert ref["predictions"].max() - ref["predictions"].min() > 0.3      # the case really spreads
    errp = float(np.abs(got["predictions"] - ref["predictions"]).max())
    err = {k: state_err(got[k], ref[k]) for k in ("E_h", "E_c", "V_h", "V_c")}
    print(mode, name, "pred err", errp, err)
    assert errp <= TOL_SPREAD_PRED[name][mode], errp
    assert np.abs(got["predictions"] - GOLD[name + "/predictions"]).max() <= TOL_SPREAD_PRED[name][mode]
    if mode != "bf16":
        for k in ("E_h", "E_c", "V_h", "V_c"):
            assert err[k] <= 10 * TOL_STATE[mode], (k, err)


@pytest.mark.parametrize("mode", ["bf16x3", "bf16"])
def test_fused_and_two_kernel_timesteps_agree(mode):
    """tspgnn_step's fused CTA-pair kernel against the two-kernel sequence (K2, K1) it replaces: same
    arithmetic, different routing (h' stays on chip, one launch per timestep)."""
    sizes = inst.mixed_sizes(24, 10, 40, seed=3)
    EV, W, C, y, nv, ne = inst.synth_batch(sizes, seed=17)
    params = orc.init_params(64, seed=5, perturb_ln=True)
    outs = []
    for fused in (1, 0):
        eng = make_engine(mode, params)
        eng.set_option("fused", fused)
        eng.plan(nv, ne, EV.src, EV.dst)
        logits, preds = eng.forward_host(W, C, 7)
        st = eng.get_states()
        outs.append((preds, st["E"][1].cpu().numpy(), st["V"][1].cpu().numpy(), st["E"][0].cpu().numpy()))
        eng.close()
    # same arithmetic, but the atomic scatter order differs (and with it the rounding of every later
    # step): compared like the parity tests, |a - b| <= tol * max(1, |b|)
    tol = 1e-4 if mode == "bf16x3" else 5e-2
    for a, b in zip(outs[0], outs[1]):
        assert state_err(a, b) <= tol, state_err(a, b)


# BASELINE config 3 (bf16 embeddings / fp32 accumulate), reported per timestep count.  Measured on B200
# (tools/bf16_error_by_timesteps.py, reference initialisers, 16 x n=20; max over elements of |err| / max(1, |ref|)):
#    T      1        2        4        8        16       32
#    E_h    1.1e-2   8.4e-3   9.8e-3   1.2e-2   1.4e-2   9.6e-3
#    E_c    1.4e-2   1.2e-2   1.9e-2   1.8e-2   1.9e-2   1.7e-2
#    V_c    2.1e-2   1.4e-2   1.4e-2   1.1e-2   2.0e-2   2.2e-2
#    pred   5.5e-4   5.2e-4   5.5e-4   3.1e-4   2.5e-4   4.0e-4
# i.e. flat in T: the error is the bf16 rounding of the stored h (2^-9 relative) pushed once through a cell,
# and the LayerNorms keep it from accumulating.  Gates: 1.5x the largest measured value.
BF16_STATE_GATE = {1: 3.2e-2, 2: 3.2e-2, 4: 3.2e-2, 8: 3.2e-2, 16: 3.2e-2, 32: 3.2e-2}
BF16_PRED_GATE = 1.0e-3


def test_bf16_mode_state_error_per_timestep_count():
    EV, W, C, y, nv, ne = inst.synth_batch([20] * 16, seed=42)
    params = orc.init_params(64, seed=0)
    report = {}
    for T, gate in sorted(BF16_STATE_GATE.items()):
        got = run_engine("bf16", params, EV, W, C, nv, ne, T)
        ref = orc.forward(params, EV.src, EV.dst, W, C, nv, ne, T, dtype=np.float64)
        err = {k: state_err(got[k], ref[k]) for k in ("E_h", "E_c", "V_h", "V_c")}
        errp = float(np.abs(got["predictions"] - ref["predictions"]).max())
        report[T] = (errp, err)
        print("bf16 T=%d pred err %.2e" % (T, errp), {k: "%.1e" % v for k, v in err.items()})
        assert max(err.values()) <= gate, (T, err)
        assert errp <= BF16_PRED_GATE, (T, errp)


@pytest.mark.parametrize("mode", ["bf16x3", "bf16"])
def test_message_kernel_slot_reuse_is_race_free_under_stress(mode):
    """compute-sanitizer racecheck flags the in-place slot reuse of the message kernel (hidden activations, then the
    fp32 staging, written by different threads of a chain; ordered through mbarriers the tool does not model).
    Stress: the same timestep from the same state, many times; the edge states depend only on the vertex messages
    (plain stores, no atomics) and must be bit-identical every time, the vertex states (atomic scatter order) equal
    to rounding.  A write-after-write race on the slot would show up as a sporadic bit difference."""
    import torch
    rng = np.random.RandomState(5)
    EV, W, C, y, nv, ne = inst.synth_batch(inst.mixed_sizes(48, 12, 40, seed=9), seed=21)
    params = orc.init_params(64, seed=6, perturb_ln=True)
    nV, nE = int(nv.sum()), int(ne.sum())
    Vh = np.abs(rng.normal(size=(nV, 64))).astype(np.float32); Vc = rng.normal(size=(nV, 64)).astype(np.float32)
    Eh = np.abs(rng.normal(size=(nE, 64))).astype(np.float32); Ec = rng.normal(size=(nE, 64)).astype(np.float32)
    eng = make_engine(mode, params)
    eng.plan(nv, ne, EV.src, EV.dst)
    t = lambda a: torch.from_numpy(a).cuda()
    dVh, dVc, dEh, dEc = t(Vh), t(Vc), t(Eh), t(Ec)
    first = None
    for rep in range(300):
        eng.set_states(Vh=dVh, Vc=dVc, Eh=dEh, Ec=dEc)
        eng.step(1)
        st = eng.get_states()
        cur = (st["E"][1].clone(), st["E"][0].clone(), st["V"][1].clone())
        if first is None:
            first = cur
            continue
        assert torch.equal(cur[0], first[0]) and torch.equal(cur[1], first[1]), "edge states differ in repeat %d" % rep
        assert state_err(cur[2].cpu().numpy(), first[2].cpu().numpy()) <= 2e-5, rep
    eng.close()


@pytest.mark.parametrize("mode", MODES)
def test_single_timestep_from_random_state(mode):
    """One while_body iteration (graphnn.py:142-173) from arbitrary (c,h): isolates the step kernels."""
    import torch
    rng = np.random.RandomState(0)
    EV, W, C, y, nv, ne = inst.synth_batch([9, 14, 11], seed=4)
    params = orc.init_params(64, seed=8, perturb_ln=True)
    nV, nE = int(nv.sum()), int(ne.sum())
    Vh = np.abs(rng.normal(size=(nV, 64))).astype(np.float32); Vc = rng.normal(size=(nV, 64)).astype(np.float32)
    Eh = np.abs(rng.normal(size=(nE, 64))).astype(np.float32); Ec = rng.normal(size=(nE, 64)).astype(np.float32)
    eng = make_engine(mode, params)
    eng.plan(nv, ne, EV.src, EV.dst)
    t = lambda a: torch.from_numpy(a).cuda()
    eng.set_states(Vh=t(Vh), Vc=t(Vc), Eh=t(Eh), Ec=t(Ec))
    eng.step(1)
    st = eng.get_states()
    eng.close()
    P64 = {k: v.astype(np.float64) for k, v in params.items()}
    mE = orc.mlp(Eh.astype(np.float64), P64, "TSP/E_msg_V"); mV = orc.mlp(Vh.astype(np.float64), P64, "TSP/V_msg_E")
    xV = np.zeros((nV, 64)); np.add.at(xV, EV.src, mE); np.add.at(xV, EV.dst, mE)
    xE = mV[EV.src] + mV[EV.dst]
    rVc, rVh = orc.lnlstm(xV, Vc.astype(np.float64), Vh.astype(np.float64), P64, "TSP/V_cell/layer_norm_basic_lstm_cell")
    rEc, rEh = orc.lnlstm(xE, Ec.astype(np.float64), Eh.astype(np.float64), P64, "TSP/E_cell/layer_norm_basic_lstm_cell")
    tol = {"simt": 2e-5, "bf16x3": 5e-5, "bf16": 1e-1}[mode]
    for name, got, ref in (("V_c", st["V"][0], rVc), ("V_h", st["V"][1], rVh), ("E_c", st["E"][0], rEc), ("E_h", st["E"][1], rEh)):
        err = state_err(got.cpu().numpy(), ref)
        print(mode, name, err)
        assert err <= tol, (name, err)


@pytest.mark.parametrize("mode", MODES)
def test_zero_time_steps_returns_initial_embeddings(mode):
    EV, W, C, y, nv, ne = inst.synth_batch([6, 7], seed=1)
    params = orc.init_params(64, seed=2, perturb_ln=True)
    got = run_engine(mode, params, EV, W, C, nv, ne, 0)
    ref = orc.forward(params, EV.src, EV.dst, W, C, nv, ne, 0)
    tol = 1e-5 if mode != "bf16" else 2e-2
    assert np.abs(got["E_h"] - ref["E_h"]).max() <= tol
    assert np.abs(got["V_h"] - ref["V_h"]).max() <= tol
    assert np.all(got["E_c"] == 0) and np.all(got["V_c"] == 0)
    assert np.abs(got["predictions"] - ref["predictions"]).max() <= TOL_PRED[mode]


@pytest.mark.parametrize("mode", MODES)
def test_full_size_north_star_config_matches_oracle(mode):
    """BASELINE config 2: 128 x n=40, d=64, 32 timesteps -- predictions vs the float64 oracle."""
    sizes = [40] * 128
    EV, W, C, y, nv, ne = inst.synth_batch(sizes, seed=42)
    params = orc.init_params(64, seed=0)
    got = run_engine(mode, params, EV, W, C, nv, ne, 32)
    ref = orc.forward(params, EV.src, EV.dst, W, C, nv, ne, 32, dtype=np.float64)
    errp = float(np.abs(got["predictions"] - ref["predictions"]).max())
    errs = state_err(got["E_h"], ref["E_h"])
    print(mode, "north-star pred err", errp, "E_h err", errs)
    assert errp <= TOL_PRED[mode] and errs <= TOL_STATE[mode]


@pytest.mark.parametrize("mode", ["simt", "bf16x3"])
def test_batch_members_are_independent_at_full_size(mode):
    """Size-independent property: the first instances of a big mixed batch give the same logits
    when run alone (block-diagonal EV, instance_loader.py:56-66)."""
    sizes = inst.mixed_sizes(96, 20, 60, seed=7)
    insts = inst.synth_instances(sizes, seed=100)
    params = orc.init_params(64, seed=1, perturb_ln=True)
    EV, W, C, y, nv, ne = inst.create_batch(insts)
    full = run_engine(mode, params, EV, W, C, nv, ne, 8)
    k = 5
    e1 = int(ne[:k].sum())
    EVs, Ws, _, _, nvs, nes = inst.create_batch(insts[:k])
    sub = run_engine(mode, params, EVs, Ws, C[:e1], nvs, nes, 8)
    assert np.abs(full["logits"][:k] - sub["logits"]).max() <= 2e-5
    ref = orc.forward(params, EVs.src, EVs.dst, Ws, C[:e1], nvs, nes, 8)
    assert np.abs(sub["predictions"] - ref["predictions"]).max() <= TOL_PRED[mode]


def test_config4_mixed_512_and_config5_large_instances_at_full_size():
    """BASELINE config 4 (512 instances, n in 20..60, 32 timesteps; one GPU's view of the whole batch) and
    config 5 (n = 160 and 320, 32 timesteps) at their full sizes, where the float64 oracle for the whole batch
    would take many minutes: instances are independent blocks (instance_loader.py:56-66), so members picked
    from the big batch must reproduce the oracle's predictions for those members alone."""
    params = orc.init_params(64, seed=3, perturb_ln=True)
    # config 4
    sizes = inst.mixed_sizes(512, 20, 60, seed=11)
    insts = inst.synth_instances(sizes, seed=300, two_opt_sweeps=0)
    EV, W, C, y, nv, ne = inst.create_batch(insts)
    got = run_engine("bf16x3", params, EV, W, C, nv, ne, 32)
    assert np.isfinite(got["predictions"]).all() and got["predictions"].shape == (512,)
    eoff = np.concatenate([[0], np.cumsum(ne)])
    for k in (0, 255, 511):
        EVs, Ws, _, _, nvs, nes = inst.create_batch([insts[k]])
        ref = orc.forward(params, EVs.src, EVs.dst, Ws, C[eoff[k]:eoff[k + 1]], nvs, nes, 32)
        assert abs(got["predictions"][k] - ref["predictions"][0]) <= TOL_PRED["bf16x3"], k
    # config 5: one n=160 and one n=320 instance inside a batch of large instances
    big = inst.synth_instances([320, 160, 80, 160], seed=500, two_opt_sweeps=0)
    EV, W, C, y, nv, ne = inst.create_batch(big)
    got = run_engine("bf16x3", params, EV, W, C, nv, ne, 32)
    eoff = np.concatenate([[0], np.cumsum(ne)])
    for k in (0, 1):
        EVs, Ws, _, _, nvs, nes = inst.create_batch([big[k]])
        ref = orc.forward(params, EVs.src, EVs.dst, Ws, C[eoff[k]:eoff[k + 1]], nvs, nes, 32)
        err = abs(got["predictions"][k] - ref["predictions"][0])
        print("config 5 member n=%d pred err %.2e" % (nv[k], err))
        assert err <= TOL_PRED["bf16x3"], k


def test_session_drop_in_surface_with_dense_ev():
    import tsp_gnn_b200 as tg
    EV, W, C, y, nv, ne = inst.synth_batch([8, 9, 7, 10], seed=6)
    GNN = tg.build_network(64)
    with tg.Session(GNN) as sess:
        sess.run(tg.global_variables_initializer(seed=3))
        params = sess.get_variables()
        feed = {GNN["EV"]: EV.toarray(), GNN["W"]: W, GNN["C"]: C, GNN["time_steps"]: 5, GNN["route_exists"]: y,
                GNN["n_vertices"]: nv, GNN["n_edges"]: ne}
        loss, acc, preds, TP, FP, TN, FN = sess.run(
            [GNN["loss"], GNN["acc"], GNN["predictions"], GNN["TP"], GNN["FP"], GNN["TN"], GNN["FN"]], feed_dict=feed)
        states = sess.run(GNN["last_states"], feed_dict=feed)
        no_label = {k: v for k, v in feed.items() if k is not GNN["route_exists"]}
        with pytest.raises(ValueError, match="route_exists"):
            sess.run(GNN["train_step"], feed_dict=no_label)      # the training step needs labels
    ref = orc.forward(params, EV.src, EV.dst, W, C, nv, ne, 5)
    m = orc.metrics(ref["logits"], y)
    assert np.abs(preds - ref["predictions"]).max() <= 1e-4
    assert abs(loss - m["loss"]) <= 1e-4 and acc == pytest.approx(m["acc"])
    assert (TP, FP, TN, FN) == (m["TP"], m["FP"], m["TN"], m["FN"])
    assert np.abs(states["E"].h - ref["E_h"]).max() <= 1e-4 and states["V"].c.shape == (34, 64)


def test_error_behaviour():
    from tsp_gnn_b200.engine import Engine
    from tsp_gnn_b200._lib import TspGnnError
    EV, W, C, y, nv, ne = inst.synth_batch([5, 6], seed=1)
    eng = Engine(64, "simt", 0)
    with pytest.raises(TspGnnError, match="set_params"):
        eng.step(1)
    eng.set_params(P.init_params(64, seed=0))
    with pytest.raises(TspGnnError, match="plan"):
        eng.step(1)
    bad = EV.dst.copy(); bad[0] = 9            # vertex of the other instance
    with pytest.raises(TspGnnError, match="outside instance 0"):
        eng.plan(nv, ne, EV.src, bad)
    with pytest.raises(TspGnnError, match="expected 115529"):
        eng.set_params(np.zeros(10, np.float32))
    eng.plan(nv, ne, EV.src, EV.dst)
    with pytest.raises(ValueError):
        eng.forward_host(W[:-1], C[:-1], 2)
    eng.close()
    with pytest.raises(TspGnnError, match="only d=64"):
        Engine(32, "simt", 0)


@pytest.mark.parametrize("mode", ["simt", "bf16x3"])
def test_replanning_and_repeatability(mode):
    params = orc.init_params(64, seed=5)
    eng = make_engine(mode, params)
    outs = []
    for sizes, seed in (([12, 13], 1), ([30] * 6, 2), ([12, 13], 1)):
        EV, W, C, y, nv, ne = inst.synth_batch(sizes, seed=seed)
        eng.plan(nv, ne, EV.src, EV.dst)
        outs.append(eng.forward_host(W, C, 6)[1])
        again = eng.forward_host(W, C, 6)[1]
        assert np.abs(again - outs[-1]).max() <= 2e-6      # atomics may reorder fp32 sums
    eng.close()
    assert np.abs(outs[0] - outs[2]).max() <= 2e-6


@pytest.mark.parametrize("mode", ["bf16x3", "bf16"])
def test_shuffled_edge_rows_match_oracle(mode):
    """The message kernel's scatter plan (per tile the (row, endpoint) pairs sorted by vertex) must not rely on
    the reference's edge order (rows sorted by src, instance_loader.py:60): the rows of every instance are
    shuffled, with a tail tile (sum E = 1157) and a sparse instance among them."""
    EV, W, C, y, nv, ne = inst.synth_batch([30, 17, 25, 9, 21], seed=23, connectivity=0.85)
    rs = np.random.RandomState(4)
    src, dst, W, C = EV.src.copy(), EV.dst.copy(), W.copy(), C.copy()
    off = 0
    for m in ne:
        p = off + rs.permutation(int(m))
        src[off:off + m], dst[off:off + m] = src[p], dst[p]
        W[off:off + m], C[off:off + m] = W[p], C[p]
        off += int(m)
    params = orc.init_params(64, seed=9, perturb_ln=True)
    eng = make_engine(mode, params)
    eng.plan(nv, ne, src, dst)
    logits, preds = eng.forward_host(W, C, 5)
    st = eng.get_states()
    eng.close()
    ref = orc.forward(params, src, dst, W, C, nv, ne, 5, dtype=np.float64)
    assert np.abs(preds - ref["predictions"]).max() <= TOL_PRED[mode]
    for k, v in (("E_h", st["E"][1]), ("E_c", st["E"][0]), ("V_h", st["V"][1]), ("V_c", st["V"][0])):
        assert state_err(v.cpu().numpy(), ref[k]) <= TOL_STATE[mode], k


@pytest.mark.parametrize("mode", ["simt", "bf16x3"])
def test_cuda_path_matches_reference_graph_code_fixtures(mode):
    """Predictions of the CUDA path against the outputs of the reference's own model.py / graphnn.py /
    mlp.py executed on the numpy TF1 stand-in (tests/golden/make_reference_golden.py)."""
    import make_reference_golden as mrg
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_shim_forward.npz"))
    for name in sorted(mrg.CASES):
        sizes, iseed, pseed, T, conn = mrg.CASES[name]
        EV, W, C, y, nv, ne = inst.synth_batch(sizes, seed=iseed, connectivity=conn)
        params = orc.init_params(64, seed=pseed, perturb_ln=True)
        got = run_engine(mode, params, EV, W, C, nv, ne, T)
        assert np.abs(got["predictions"] - gold[name + "/predictions"]).max() <= TOL_PRED[mode]
        assert state_err(got["V_h"], gold[name + "/V_h"]) <= TOL_STATE[mode]
        assert state_err(got["E_c"][:48], gold[name + "/E_c_head"]) <= TOL_STATE[mode]


@pytest.mark.parametrize("mode", ["bf16x3", "bf16"])
def test_large_layernorm_gains_take_the_clamped_path(mode):
    """gamma of the input / forget gates scaled x12: tspgnn_set_params must select the kernel variant that
    clamps the logistic exponents (the unclamped (1+2^a)(1+2^b) product would overflow); gates saturate."""
    EV, W, C, y, nv, ne = inst.synth_batch([11, 12, 13], seed=3)
    params = orc.init_params(64, seed=4, perturb_ln=True)
    for v in ("V", "E"):
        for g in ("input", "forget", "output"):
            params["TSP/%s_cell/layer_norm_basic_lstm_cell/%s/gamma" % (v, g)] *= 12.0
    got = run_engine(mode, params, EV, W, C, nv, ne, 5)
    ref = orc.forward(params, EV.src, EV.dst, W, C, nv, ne, 5, dtype=np.float64)
    assert np.all(np.isfinite(got["E_h"])) and np.all(np.isfinite(got["V_c"]))
    # gains of 12 amplify operand rounding 12x (the fp32 oracle itself is 5e-5 off the float64 one here):
    # predictions keep the north-star bound in the parity mode, states get a proportionally wider one
    tol_p, tol_s = (1e-4, 1e-2) if mode == "bf16x3" else (5e-2, 4.0)
    assert np.abs(got["predictions"] - ref["predictions"]).max() <= tol_p
    assert state_err(got["E_h"], ref["E_h"]) <= tol_s


def test_generalisation_size_n80_matches_oracle():
    """BASELINE config 5 family (larger graphs than trained on): 6 instances of n=80, 3160 edges each."""
    EV, W, C, y, nv, ne = inst.synth_batch([80] * 6, seed=9)
    params = orc.init_params(64, seed=2)
    got = run_engine("bf16x3", params, EV, W, C, nv, ne, 32)
    ref = orc.forward(params, EV.src, EV.dst, W, C, nv, ne, 32, dtype=np.float64)
    assert np.abs(got["predictions"] - ref["predictions"]).max() <= 1e-4
    assert state_err(got["V_h"], ref["V_h"]) <= 1e-4
