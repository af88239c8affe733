"""The float64 yardstick and its independence: numpy direct convolution vs the torch restatement, finite differences vs
autograd, closed-form float64 Adam vs the oracle's TFAdam -- all on CPU."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import xt_oracle as orc  # noqa: E402
from oracle import np_f64  # noqa: E402


@pytest.mark.parametrize("case", ["ppo_cnn", "impala_cnn", "dqn_cnn", "ppo_mlp"])
def test_numpy_direct_conv_pins_the_torch_restatement(case):
    arch = {"ppo_cnn": orc.ppo_cnn_arch, "impala_cnn": orc.impala_cnn_arch, "dqn_cnn": orc.dqn_cnn_arch,
            "ppo_mlp": orc.ppo_mlp_arch}[case]()
    w = orc.init_weights(arch, seed=3)
    rng = np.random.default_rng(0)
    for k in w:
        if k.endswith("/bias"):
            w[k] = (rng.standard_normal(w[k].shape) * 0.05).astype(np.float32)
    if arch["input_dtype"] == "uint8":
        obs = rng.integers(0, 256, (3,) + arch["state_dim"], dtype=np.uint8)
    else:
        obs = rng.standard_normal((3,) + arch["state_dim"]).astype(np.float32)
    ref = np_f64.forward(arch, w, obs)
    with orc.precision("f64"), torch.no_grad():
        t64 = orc.forward(arch, w, obs, keep=True)
    with torch.no_grad():
        t32 = orc.forward(arch, w, obs, keep=True)
    for name, _, _, _ in arch["layers"]:
        a = ref[name].reshape(3, -1)
        scale = max(np.abs(a).max(), 1e-12)
        assert t64[name].dtype == torch.float64
        assert np.abs(t64[name].reshape(3, -1).numpy() - a).max() / scale < 1e-13, name       # two float64 restatements
        assert np.abs(t32[name].reshape(3, -1).numpy() - a).max() / scale < 2e-6, name        # the fp32 oracle


def test_same_padding_matches_tensorflow_definition():
    # 84 -> SAME, k=8, s=4: out 21, total pad 4 -> (2, 2); 21 -> k=4, s=2: out 11, total 3 -> (1, 2) smaller half first
    assert np_f64._same_pad(84, 8, 4) == (2, 2) and np_f64._same_pad(21, 4, 2) == (1, 2)
    assert np_f64._same_pad(11, 3, 1) == (1, 1) and np_f64._same_pad(10, 3, 2) == (0, 1)
    x = np.arange(2 * 5 * 5 * 1, dtype=np.float64).reshape(2, 5, 5, 1)
    k = np.ones((3, 3, 1, 1))
    y = np_f64.conv2d_nhwc(x, k, np.zeros(1), 2, "same")
    assert y.shape == (2, 3, 3, 1) and y[0, 0, 0, 0] == x[0, 0:2, 0:2, 0].sum() and y[0, 1, 1, 0] == x[0, 1:4, 1:4, 0].sum()


def test_float64_ppo_gradient_matches_finite_differences():
    arch = orc.ppo_cnn_arch()
    w = orc.init_weights(arch, seed=1)
    ro = orc.synth_ppo_rollout(2, 1, 6)
    adv, ov, tv = orc.gae(ro["value"][0], ro["reward"], ro["done"])
    with orc.precision("f64"):
        ln = orc.PpoLearner(arch, w, batch_size=6, ent_coef=0.003, clip_ratio=0.1)
        args = (ro["obs"], ro["action"], ro["logp"], adv, ov, tv)
        loss, grads = ln.loss_and_grads(*args)
        assert loss.dtype == torch.float64 and grads[0].dtype == torch.float64
        rng = np.random.default_rng(5)
        d = [torch.from_numpy(rng.standard_normal(tuple(p.shape))) for p in ln.params]
        dirderiv = float(sum((g * x).sum() for g, x in zip(grads, d)))
        eps = 1e-8        # small enough that no ReLU / clip kink is crossed (1e-6 already crosses a few: 0.3 % off)
        vals = []
        for sgn in (+1, -1):
            with torch.no_grad():
                for p, x in zip(ln.params, d):
                    p.add_(sgn * eps * x)
            vals.append(float(ln.loss_and_grads(*args)[0].detach()))
            with torch.no_grad():
                for p, x in zip(ln.params, d):
                    p.sub_(sgn * eps * x)
        fd = (vals[0] - vals[1]) / (2 * eps)
    assert abs(fd - dirderiv) < 1e-6 * max(1.0, abs(dirderiv)), (fd, dirderiv)   # fd noise ~ 1e-16 / 1e-8


def test_float64_adam_matches_closed_form():
    rng = np.random.default_rng(0)
    p0 = rng.standard_normal(50)
    gs = [rng.standard_normal(50) * 0.01 for _ in range(7)]
    with orc.precision("f64"):
        params = [torch.from_numpy(p0.copy())]
        opt = orc.TFAdam(params, 2.5e-4)
        for g in gs:
            opt.step([torch.from_numpy(g)])
    np.testing.assert_allclose(params[0].numpy(), np_f64.adam_steps(p0, gs, 2.5e-4), rtol=0, atol=1e-15)
    # and the fp32 optimiser of the golden fixtures stays within fp32 rounding of it
    params32 = [torch.from_numpy(p0.astype(np.float32))]
    opt32 = orc.TFAdam(params32, 2.5e-4)
    for g in gs:
        opt32.step([torch.from_numpy(g.astype(np.float32))])
    upd64 = np_f64.adam_steps(p0, gs, 2.5e-4) - p0
    assert np.abs((params32[0].numpy() - p0.astype(np.float32)) - upd64).max() < 2e-6 * np.abs(upd64).max() + 2e-7


def test_rmsprop_and_lr_schedule_restatements():
    """TFRMSProp (centred, rms slot initialised to ones) against a scalar hand computation; linear_cosine_decay at its
    documented end points (tf.train.linear_cosine_decay: (alpha + linear) * cosine + beta)."""
    p = [torch.tensor([1.0, -2.0], dtype=torch.float64)]
    opt = orc.TFRMSProp(p, 0.01, decay=0.9, eps=0.1)
    g = torch.tensor([0.5, -0.25], dtype=torch.float64)
    opt.step([g])
    mg = 0.1 * g.numpy(); ms = 0.9 * 1.0 + 0.1 * g.numpy() ** 2
    want = np.array([1.0, -2.0]) - 0.01 * g.numpy() / np.sqrt(ms - mg ** 2 + 0.1)
    np.testing.assert_allclose(p[0].numpy(), want, rtol=0, atol=1e-15)
    assert abs(orc.linear_cosine_decay(0.01, 0, 20000.0, beta=1e-4) - 0.01 * (1.0 + 1e-4)) < 1e-15
    assert abs(orc.linear_cosine_decay(0.01, 20000, 20000.0, beta=1e-4) - 0.01 * 1e-4) < 1e-15
    assert abs(orc.linear_cosine_decay(0.01, 10 ** 9, 20000.0, beta=1e-4) - 0.01 * 1e-4) < 1e-15       # clamped at decay_steps
    mid = orc.linear_cosine_decay(0.01, 10000, 20000.0, beta=0.0)
    assert abs(mid - 0.01 * 0.5 * 0.5 * (1 + np.cos(np.pi * 0.5))) < 1e-15
