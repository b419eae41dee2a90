"""The training-step oracle (oracle/tspgnn_oracle_grad.py): hand-derived reverse pass against
torch.autograd, against finite differences, and known answers for clip + Adam."""
import numpy as np
import pytest

from oracle import tspgnn_oracle as orc
from oracle import tspgnn_oracle_grad as og
from tsp_gnn_b200 import instances as inst


def small_batch(seed=3, sizes=(5, 7, 6, 4)):
    EV, W, C, y, nv, ne = inst.synth_batch(list(sizes), seed=seed)
    return EV, W, C, y, nv, ne


def test_forward_of_grad_oracle_is_the_forward_oracle():
    EV, W, C, y, nv, ne = small_batch()
    params = orc.init_params(64, seed=1, perturb_ln=True)
    a = og.forward_backward(params, EV.src, EV.dst, W, C, nv, ne, y, 5)
    f = orc.forward(params, EV.src, EV.dst, W, C, nv, ne, 5)
    assert np.abs(a["logits"] - f["logits"]).max() == 0.0
    assert abs(a["loss"] - orc.metrics(f["logits"], y)["loss"]) < 1e-15


@pytest.mark.parametrize("T", [0, 1, 4])
def test_manual_reverse_pass_matches_autograd(T):
    EV, W, C, y, nv, ne = small_batch()
    params = orc.init_params(64, seed=1, perturb_ln=True)
    a = og.forward_backward(params, EV.src, EV.dst, W, C, nv, ne, y, T)
    b = og.torch_forward_backward(params, EV.src, EV.dst, W, C, nv, ne, y, T)
    assert abs(a["loss"] - b["loss"]) < 1e-14
    for k in a["grads"]:
        scale = np.abs(b["grads"][k]).max() + 1e-30
        assert np.abs(a["grads"][k] - b["grads"][k]).max() <= 1e-9 * scale + 1e-18, k


def test_gradient_matches_finite_differences():
    EV, W, C, y, nv, ne = small_batch(sizes=(4, 5))
    params = orc.init_params(64, seed=2, perturb_ln=True)
    g = og.forward_backward(params, EV.src, EV.dst, W, C, nv, ne, y, 3)["grads"]
    rng = np.random.RandomState(0)
    P64 = {k: v.astype(np.float64) for k, v in params.items()}
    for name in ("TSP/E_cell/layer_norm_basic_lstm_cell/kernel", "TSP/V_msg_E_MLP_layer_2/bias",
                 "TSP/V_cell/layer_norm_basic_lstm_cell/state/gamma", "V_init", "E_init_MLP_MLP_layer_1/kernel",
                 "E_vote_MLP_layer_4/kernel"):
        direction = rng.normal(size=params[name].shape)
        eps = 1e-6
        lo = dict(P64); hi = dict(P64)
        lo[name] = P64[name] - eps * direction
        hi[name] = P64[name] + eps * direction
        fl = orc.metrics(orc.forward(lo, EV.src, EV.dst, W, C, nv, ne, 3)["logits"], y)["loss"]
        fh = orc.metrics(orc.forward(hi, EV.src, EV.dst, W, C, nv, ne, 3)["logits"], y)["loss"]
        fd = (fh - fl) / (2 * eps)
        an = float((g[name] * direction).sum())
        assert abs(fd - an) <= 1e-6 * max(1.0, abs(an)) + 1e-9, (name, fd, an)


def test_sharded_gradients_sum_to_the_batch_gradient():
    """loss is a mean over instances, so per-shard gradients computed with the
    global batch as divisor add up to the whole-batch gradient (what the all-reduce sums)."""
    from tsp_gnn_b200 import sharding
    EV, W, C, y, nv, ne = small_batch(sizes=(5, 7, 6, 4, 5))
    params = orc.init_params(64, seed=1)
    full = og.forward_backward(params, EV.src, EV.dst, W, C, nv, ne, y, 3)
    parts = sharding.partition_instances(ne, 2)
    acc = {k: np.zeros_like(v, dtype=np.float64) for k, v in params.items()}
    loss = 0.0
    for idx in parts:
        s, d, w, c, pv, pe = sharding.take_instances(idx, EV.src, EV.dst, W, C, nv, ne)
        r = og.forward_backward(params, s, d, w, c, pv, pe, np.asarray(y)[idx], 3, global_batch=len(ne))
        loss += r["loss"]
        for k in acc:
            acc[k] += r["grads"][k]
    # take_instances hands out float32 W / C, hence 1e-7-level input differences
    assert abs(loss - full["loss"]) < 1e-8
    for k in acc:
        assert np.abs(acc[k] - full["grads"][k]).max() <= 1e-5 * (np.abs(full["grads"][k]).max() + 1e-30), k


def test_clip_and_adam_known_answers():
    params = {"a": np.array([1.0, -2.0]), "b": np.array([[0.5]])}
    grads = {"a": np.array([3.0, 4.0]), "b": np.array([[12.0]])}          # global norm 13
    st = og.new_optimizer_state(params)
    new, gnorm = og.apply_gradients(params, grads, st, lr=0.1, l2=0.0, clip=0.65)
    assert abs(gnorm - 13.0) < 1e-12
    # first Adam step: lr_t = lr*sqrt(1-b2)/(1-b1); m = (1-b1) g; v = (1-b2) g^2  =>  step = lr * g/(|g| + eps*sqrt(1-b2)) ~ lr*sign(g)
    for k in params:
        gk = grads[k] * (0.65 / 13.0)
        expect = params[k] - 0.1 * np.sqrt(1 - 0.999) / (1 - 0.9) * (0.1 * gk) / (np.sqrt(0.001 * gk * gk) + 1e-8)
        assert np.abs(new[k] - expect).max() < 1e-12
        assert np.abs(np.abs(new[k] - params[k]) - 0.1).max() < 1e-5
    assert st["step"] == 1
    # below the clip threshold gradients pass unscaled; the L2 term adds l2 * var
    small = {"a": np.array([0.3, 0.0]), "b": np.array([[0.4]])}
    st2 = og.new_optimizer_state(params)
    _, gnorm2 = og.apply_gradients(params, small, st2, lr=0.1, l2=0.5, clip=10.0)
    expect_norm = np.sqrt((0.3 + 0.5) ** 2 + (0.0 - 1.0) ** 2 + (0.4 + 0.25) ** 2)
    assert abs(gnorm2 - expect_norm) < 1e-12
    assert np.allclose(st2["m"]["a"], 0.1 * np.array([0.8, -1.0]))


def test_gradient_kink_sensitivity_documented():
    """Why the tensor-core modes are gated looser than fp32 on gradients (tests/test_gpu_train.py):
    the loss is piecewise smooth, and perturbing the recurrent state by 3e-5 relative -- the size of
    the bf16x3 mode's state error -- flips ReLU masks and moves gradient entries by ~1e-2 of a
    tensor's scale on a tiny batch, while the loss moves by < 1e-6."""
    EV, W, C, y, nv, ne = inst.synth_batch([5, 12, 20, 7, 33, 9], seed=11)
    params = orc.init_params(64, seed=5, perturb_ln=True)
    base = og.forward_backward(params, EV.src, EV.dst, W, C, nv, ne, y, 3)
    rng = np.random.RandomState(1)
    orig = orc.lnlstm

    def noisy(x, c, h, P, b):
        nc, nh = orig(x, c, h, P, b)
        return nc * (1 + 3e-5 * rng.standard_normal(nc.shape)), nh * (1 + 3e-5 * rng.standard_normal(nh.shape))

    orc.lnlstm = noisy
    try:
        pert = og.forward_backward(params, EV.src, EV.dst, W, C, nv, ne, y, 3)
    finally:
        orc.lnlstm = orig
    assert abs(pert["loss"] - base["loss"]) < 1e-6
    worst = max(np.abs(pert["grads"][k] - base["grads"][k]).max() / (np.abs(base["grads"][k]).max() + 1e-30)
                for k in base["grads"])
    assert 1e-4 < worst < 5e-2, worst
