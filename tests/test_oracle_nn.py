"""CPU checks of the oracle's network side: operand roundings against numpy/torch, the canonical exp, and the forward formula
(DenseNet.jl:294-304) against a float64 numpy evaluation."""
import numpy as np
import torch

import oracle
from conftest import GAME_SPECS

f32 = np.float32


def test_rounding_modes_match_numpy_and_torch():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.normal(0, 1, 50000), rng.normal(0, 1e-4, 20000), rng.normal(0, 1e-6, 20000), rng.normal(0, 3e4, 10000),
                        [65504, 65519.9, 65520, 70000, -1e9, 2.0**-24, 2.0**-25, 2.0**-25 * 1.0001, 6.1e-5, 0.0]]).astype(f32)
    with np.errstate(over="ignore"):
        h = x.astype(np.float16)
    want = np.where(np.isinf(h), np.sign(x) * 65504, h.astype(f32)).astype(f32)       # cvt.rn.satfinite.f16.f32
    assert np.array_equal(oracle.round_to(x, "f16"), want)
    assert np.array_equal(oracle.round_to(x, "bf16"), torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy())


def test_canonical_exp_accuracy():
    x = np.concatenate([-np.random.default_rng(1).uniform(0, 86, 200000), np.linspace(-1, 1, 1001)]).astype(f32)
    y = oracle.c_expf(x)
    ref = np.exp(x.astype(np.float64))
    ulp = np.abs(y.astype(np.float64) - ref) / np.spacing(ref.astype(f32)).astype(np.float64)
    assert ulp.max() < 2.0
    assert np.all(oracle.c_expf(np.array([-88, -100, -1e9], f32)) == 0) and oracle.c_expf(np.array([0], f32))[0] == 1
    s = oracle.sigmoid(np.array([-30, -1, 0, 1, 30], f32))
    assert np.allclose(s, 1 / (1 + np.exp(-np.array([-30, -1, 0, 1, 30], np.float64))), rtol=1e-6)


def test_forward_formula_vs_float64():
    spec = oracle.Spec(*GAME_SPECS["connect4"])
    rng = np.random.default_rng(3)
    n, k = 64, 3
    g = lambda o, i: (rng.uniform(-1, 1, (o, i)) * np.sqrt(6 / (o + i))).astype(f32)
    base, res, pol, pb, val, vb = g(n, 84), [g(n, n) for _ in range(k)], g(7, n), rng.normal(0, .1, 7).astype(f32), g(1, n), rng.normal(0, .1, 1).astype(f32)
    net = oracle.Net(base, res, pol, pb, val, vb)
    x = (rng.uniform(0, 1, (50, 84)) < 0.3).astype(f32)
    logits, v = net.forward(x)
    b = np.maximum(base.astype(np.float64) @ x.T.astype(np.float64), 0)
    for w in res:
        b = np.maximum(b + np.maximum(w.astype(np.float64) @ b, 0), 0)
    want_l = (pol.astype(np.float64) @ b + pb[:, None]).T
    want_v = 1 / (1 + np.exp(-(val.astype(np.float64) @ b + vb)))[0]
    assert np.abs(logits - want_l).max() < 1e-5 and np.abs(v - want_v).max() < 1e-6
    for mode, tol in ((oracle.Net.BF16, 5e-2), (oracle.Net.F16, 8e-3), (oracle.Net.BF16_RESID, 8e-2)):
        lm, vm = net.forward(x, mode=mode)
        assert np.abs(lm - logits).max() < tol, (mode, np.abs(lm - logits).max())
    sm = oracle.softmax(logits)
    assert np.allclose(sm.sum(1), 1, atol=1e-6) and np.all(sm > 0)
