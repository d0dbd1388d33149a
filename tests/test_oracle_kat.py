"""Pins for the CPU oracle (the reference has no tests or golden vectors — SURVEY.md 8c):
  * hand-derived known-answer tests of every formula in SURVEY.md Appendix A,
  * structural facts the reference code implies,
  * an INDEPENDENT float64 torch-autograd model of the same network for every gradient.
"""
import math

import numpy as np
import pytest
import torch

from oracle.dnn_oracle import OracleConfig, OracleDNN, compute_prior, learning_rate, reference_init
from oracle.philox import dropout_keep_mask, dropout_threshold, philox4x32_10


def test_philox_known_answers():
    # Random123 kat_vectors: philox4x32 10 rounds
    assert [int(v[0]) for v in philox4x32_10([0], [0], [0], [0], 0, 0)] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    f = 0xFFFFFFFF
    assert [int(v[0]) for v in philox4x32_10([f], [f], [f], [f], f, f)] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert [int(v[0]) for v in philox4x32_10([0x243F6A88], [0x85A308D3], [0x13198A2E], [0x03707344], 0xA4093822, 0x299F31D0)] == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_dropout_mask_semantics():
    assert dropout_threshold(0.5) == 1 << 15
    assert dropout_threshold(1.0) == 0
    m = dropout_keep_mask(42, 512, 333, 0.8)
    assert m.shape == (512, 333) and abs(m.mean() - 0.8) < 0.01
    assert np.array_equal(m, dropout_keep_mask(42, 512, 333, 0.8))
    assert not np.array_equal(m, dropout_keep_mask(43, 512, 333, 0.8))


def test_init_follows_reference():
    cfg = OracleConfig(num_layers=2, input_dim=440, hidden_dim=256, output_dim=183, batch_norm=True)
    p = reference_init(cfg, np.random.default_rng(0))
    assert p["W0"].shape == (440, 256) and abs(p["W0"].std() - 1 / math.sqrt(440)) < 2e-3  # layer.py:39-44
    assert not p["W2"].any() and not p["b0"].any()  # dnn.py:67-68, layer.py:46-48
    assert not p["moving_mean0"].any() and (p["moving_var1"] == 1).all() and not p["beta0"].any()


def test_first_loss_is_ln_O_and_first_step_only_moves_output_layer():
    cfg = OracleConfig(num_layers=2, input_dim=20, hidden_dim=16, output_dim=7)
    rng = np.random.default_rng(1)
    o = OracleDNN(cfg, reference_init(cfg, rng))
    x, y = rng.standard_normal((50, 20)).astype(np.float32), rng.integers(0, 7, 50)
    w0, w1 = o.p["W0"].copy(), o.p["W1"].copy()
    o.accumulate(x, y)
    assert abs(o.loss_sum / 50 - math.log(7)) < 1e-6
    loss = o.apply(1e-3)
    assert abs(loss - math.log(7)) < 1e-6
    assert np.array_equal(w0, o.p["W0"]) and np.array_equal(w1, o.p["W1"]) and o.p["W2"].any()
    # Adam's first step is ~ +-lr for every non-zero gradient component (a hair less: eps 1e-8 vs sqrt(v))
    moved = np.abs(o.p["W2"][o.p["W2"] != 0])
    assert (moved <= 1e-3 * (1 + 1e-6)).all() and np.median(moved) > 0.99e-3
    assert o.num_frames == 0 and o.loss_sum == 0 and not o.grads["W2"].any() and o.global_step == 1


def test_hand_computed_tiny_network():
    """1 hidden unit chain, by hand: x=[1,2], W0=[[1],[-1]] b0=[0.5] -> z=-0.5 -> relu 0 ;
    second frame x=[2,0] -> z=2.5 -> relu 2.5; W1=[[1,-1]], b1=[0,0] -> logits [2.5,-2.5]."""
    cfg = OracleConfig(num_layers=1, input_dim=2, hidden_dim=1, output_dim=2)
    p = {"W0": [[1.0], [-1.0]], "b0": [0.5], "W1": [[1.0, -1.0]], "b1": [0.0, 0.0]}
    o = OracleDNN(cfg, {k: np.array(v, np.float32) for k, v in p.items()})
    x = np.array([[1, 2], [2, 0]], np.float32)
    logits, caches = o.forward(x, training=True)
    assert np.allclose(logits, [[0, 0], [2.5, -2.5]])
    loss, d = o.softmax_ce(logits, np.array([0, 1]))
    s = 1 / (1 + math.exp(-5.0))  # softmax prob of class 0 for frame 2
    assert abs(loss - (math.log(2) + (math.log(1 + math.exp(5.0))))) < 1e-5
    assert np.allclose(d, [[-0.5, 0.5], [s, -s]], atol=1e-6)  # softmax - onehot; frame 2: [s, (1-s) - 1]
    g = o.backward(caches, d)
    assert np.allclose(g["W1"], [[2.5 * s, -2.5 * s]], atol=1e-5)  # h = [0, 2.5]
    assert np.allclose(g["b1"], [s - 0.5, 0.5 - s], atol=1e-6)
    dh2 = s * 1 + (-s) * (-1)  # frame 2 only (frame 1 is cut by the relu)
    assert np.allclose(g["W0"], [[2 * dh2], [0]], atol=1e-5) and np.allclose(g["b0"], [dh2], atol=1e-6)


def test_adam_tf_form_by_hand():
    """g_acc=6 over 3 frames -> mean 2 -> clipped to 1; t=1: m=0.1, v=0.001,
    lr_t = lr*sqrt(1-0.999)/(1-0.9); w -= lr_t*m/(sqrt(v)+1e-8).  Second step with g=-0.3 (no clip)."""
    cfg = OracleConfig(num_layers=1, input_dim=1, hidden_dim=1, output_dim=1)
    o = OracleDNN(cfg, {"W0": np.ones((1, 1)), "b0": np.zeros(1), "W1": np.ones((1, 1)), "b1": np.zeros(1)})
    o.grads["W0"][...] = 6.0
    o.loss_sum, o.num_frames = 3.0, 3
    assert o.apply(0.01) == 1.0
    lr_t = 0.01 * math.sqrt(1 - 0.999) / (1 - 0.9)
    w = 1.0 - lr_t * 0.1 / (math.sqrt(0.001) + 1e-8)
    assert abs(o.p["W0"][0, 0] - w) < 1e-7 and abs(w - 0.99) < 1e-6
    o.grads["W0"][...] = -0.3
    o.loss_sum, o.num_frames = 1.0, 1
    o.halve_learning_rate()  # lr_fact 0.5
    o.apply(0.01)
    m = 0.1 + (-0.3 - 0.1) * 0.1
    v = 0.001 + (0.09 - 0.001) * 0.001
    lr_t2 = 0.005 * math.sqrt(1 - 0.999 ** 2) / (1 - 0.9 ** 2)
    assert abs(o.p["W0"][0, 0] - (w - lr_t2 * m / (math.sqrt(v) + 1e-8))) < 1e-7
    assert abs(learning_rate(1e-3, 0.5, 50, 100) - 1e-3 * 0.5 ** 0.5) < 1e-12  # trainer.py:110-112


def test_batchnorm_by_hand():
    """z column [1,3] -> mu=2, biased var=1, y=(z-2)/sqrt(1.001)+beta; moving stats move by 0.001."""
    cfg = OracleConfig(num_layers=1, input_dim=1, hidden_dim=1, output_dim=1, batch_norm=True, nonlin="linear")
    p = {"W0": np.ones((1, 1)), "b0": np.zeros(1), "beta0": np.array([0.25]), "moving_mean0": np.zeros(1),
         "moving_var0": np.ones(1), "W1": np.ones((1, 1)), "b1": np.zeros(1)}
    o = OracleDNN(cfg, p)
    logits, caches = o.forward(np.array([[1.0], [3.0]], np.float32), training=True)
    r = 1 / math.sqrt(1.001)
    assert np.allclose(logits[:, 0], [-r + 0.25, r + 0.25], atol=1e-6)
    assert abs(o.p["moving_mean0"][0] - 0.002) < 1e-7 and abs(o.p["moving_var0"][0] - 1.0) < 1e-7
    ev, _ = o.forward(np.array([[1.0]], np.float32), training=False)  # uses the moving statistics
    assert abs(ev[0, 0] - ((1 - 0.002) / math.sqrt(1.0 + 1e-3) + 0.25)) < 1e-6


def test_prior_and_loglik():
    prior = compute_prior([np.array([0, 1, 1], np.uint32), np.array([3], np.uint32)], 5)
    assert prior.dtype == np.float32 and np.allclose(prior, [0.25, 0.5, 0, 0.25, 0])
    cfg = OracleConfig(num_layers=1, input_dim=3, hidden_dim=4, output_dim=5)
    rng = np.random.default_rng(0)
    o = OracleDNN(cfg, reference_init(cfg, rng))
    ll = o.loglik(rng.standard_normal((2, 3)).astype(np.float32), prior)
    assert np.isposinf(ll[:, 2]).all() and np.allclose(ll[:, 1], math.log(0.2 / 0.5), atol=1e-6)  # no flooring (nnet.py:283)


# ---------------------------------------------------------------- independent autograd cross-check
def torch_model_grads(cfg, params, x, y, keep_masks):
    P = {k: torch.tensor(v, dtype=torch.float64, requires_grad=not k.startswith("moving")) for k, v in params.items()}
    a = torch.tensor(x, dtype=torch.float64)
    for l in range(cfg.num_layers):
        z = a @ P[f"W{l}"] + P[f"b{l}"]
        if cfg.batch_norm:
            mu = z.mean(0)
            var = ((z - mu) ** 2).mean(0)
            z = (z - mu) / torch.sqrt(var + cfg.bn_eps) + P[f"beta{l}"]
        if cfg.nonlin == "relu":
            z = torch.relu(z)
        elif cfg.nonlin == "sigmoid":
            z = torch.sigmoid(z)
        elif cfg.nonlin == "tanh":
            z = torch.tanh(z)
        if cfg.l2_norm:
            sig = (z ** 2).mean(1, keepdim=True)
            z = torch.where(sig > 1, z / sig, z)
        if cfg.keep_prob < 1:
            z = z / cfg.keep_prob * torch.tensor(keep_masks[l], dtype=torch.float64)
        a = z
    logits = a @ P[f"W{cfg.num_layers}"] + P[f"b{cfg.num_layers}"]
    loss = torch.nn.functional.cross_entropy(logits, torch.tensor(y, dtype=torch.long), reduction="sum")
    loss.backward()
    return float(loss.detach()), {k: v.grad.numpy() for k, v in P.items() if v.requires_grad}, logits.detach().numpy()


def test_l2norm_by_hand():
    """row [3, 0, -4, 0] -> mean square 6.25 > 1 -> row / 6.25 ; row [0.5, 0.5, 0, 0] -> 0.125 <= 1 -> unchanged"""
    cfg = OracleConfig(num_layers=1, input_dim=4, hidden_dim=4, output_dim=2, nonlin="linear", l2_norm=True)
    p = {"W0": np.eye(4), "b0": np.zeros(4), "W1": np.zeros((4, 2)), "b1": np.zeros(2)}
    o = OracleDNN(cfg, p)
    _, caches = o.forward(np.array([[3, 0, -4, 0], [0.5, 0.5, 0, 0]], np.float32), training=False)
    assert np.allclose(caches[0].y, [[0.48, 0, -0.64, 0], [0.5, 0.5, 0, 0]], atol=1e-7)


@pytest.mark.parametrize("bn,keep,nonlin", [(False, 1.0, "relu"), (True, 1.0, "relu"), (False, 0.5, "relu"),
                                            (True, 0.5, "relu"), (False, 0.7, "linear"), (True, 1.0, "linear"),
                                            (False, 1.0, "sigmoid"), (True, 0.6, "sigmoid"), (False, 0.8, "tanh"), (True, 1.0, "tanh")])
@pytest.mark.parametrize("l2", [False, True])
def test_explicit_backward_matches_float64_autograd(bn, keep, nonlin, l2):
    cfg = OracleConfig(num_layers=3, input_dim=24, hidden_dim=40, output_dim=11, batch_norm=bn, keep_prob=keep, nonlin=nonlin, l2_norm=l2)
    rng = np.random.default_rng(7)
    params = reference_init(cfg, rng)
    params["W3"] = (rng.standard_normal((40, 11)) / math.sqrt(40)).astype(np.float32)
    for l in range(4):
        params[f"b{l}"] = (0.1 * rng.standard_normal(params[f"b{l}"].shape)).astype(np.float32)
    x, y = rng.standard_normal((64, 24)).astype(np.float32), rng.integers(0, 11, 64)
    if l2:
        x[::2] *= 4  # make some frames exceed mean square 1 so both L2Norm branches are taken
    o = OracleDNN(cfg, params)
    seed = 99
    loss = o.accumulate(x, y, dropout_seed=seed)
    masks = [dropout_keep_mask(seed + l, 64, 40, keep) for l in range(3)]
    tl, tg, _ = torch_model_grads(cfg, params, x, y, masks)
    assert abs(loss - tl) < 1e-4 * abs(tl)
    for k, g in tg.items():
        if bn and k.startswith("b") and not k.startswith("beta") and int(k[1:]) < 3:
            assert np.abs(o.grads[k]).max() < 1e-4  # exactly zero in exact arithmetic
            continue
        assert np.abs(o.grads[k] - g).max() <= 2e-5 * max(1.0, np.abs(g).max()), k


def test_microbatch_accumulation_equals_one_big_batch_without_bn():
    """trainer.py:165-175: grads are SUMMED over micro-batches and divided by the total frame count."""
    cfg = OracleConfig(num_layers=2, input_dim=12, hidden_dim=16, output_dim=5)
    rng = np.random.default_rng(3)
    params = reference_init(cfg, rng)
    params["W2"] = rng.standard_normal((16, 5)).astype(np.float32)
    x, y = rng.standard_normal((40, 12)).astype(np.float32), rng.integers(0, 5, 40)
    a, b = OracleDNN(cfg, params), OracleDNN(cfg, params)
    a.accumulate(x, y)
    b.accumulate(x[:13], y[:13]); b.accumulate(x[13:], y[13:])
    for k in a.grads:
        assert np.allclose(a.grads[k], b.grads[k], atol=1e-5)
    assert abs(a.apply(1e-3) - b.apply(1e-3)) < 1e-6


def test_out_of_range_label_follows_the_tensorflow_op():
    """tf.one_hot gives an all-zero row for a label outside [0, O); softmax_cross_entropy_with_logits then contributes
    no loss, but its registered gradient is grad * (prob - labels) = softmax for that frame (trainer.py:526-531).
    Cannot occur through the reference's data path; restated literally."""
    from oracle.dnn_oracle import OracleDNN

    z = np.array([[0.0, 1.0, 2.0], [1.0, 1.0, 1.0], [3.0, 0.0, -1.0]], np.float32)
    loss, d = OracleDNN.softmax_ce(z, np.array([2, 3, -1]))
    p = np.exp(z - z.max(1, keepdims=True))
    p /= p.sum(1, keepdims=True)
    assert abs(loss - (-np.log(p[0, 2]))) < 1e-6  # only the first frame has a loss term
    assert np.allclose(d[0], p[0] - np.array([0, 0, 1]), atol=1e-7)
    assert np.allclose(d[1], p[1], atol=1e-7) and np.allclose(d[2], p[2], atol=1e-7)
