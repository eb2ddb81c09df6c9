"""The training-step oracle (oracle/train_oracle.cpp) against an independent numpy float64 statement of the same mathematics
(networkf forward DenseNet.jl:161-186, lossTot train.jl:12-15, Flux 0.12.6 ADAM + WeightDecay) and central finite differences."""
import numpy as np
import pytest

import oracle
from helpers_train import make_batch, make_net, np_forward_loss, np_adam_step, flat_of


@pytest.mark.parametrize("k", [0, 2])
def test_loss_matches_numpy_f64(k):
    inp, n, A, FS, B = 18, 24, 9, 9, 37
    d = make_net(inp, n, k, A, FS, seed=1)
    batch = make_batch(inp, A, FS, B, seed=2)
    tr = oracle.Trainer(inp, n, k, A, FS)
    tr.set_params(d)
    got = tr.loss_grad(*batch, want_grad=False)
    want = np_forward_loss(d, *batch)
    assert np.allclose(got, want, rtol=2e-5, atol=1e-6), (got, want)


@pytest.mark.parametrize("k,B", [(0, 5), (2, 300)])
def test_gradient_matches_finite_differences(k, B):
    """B=300 crosses one 256-sample slice boundary of the weight-gradient sum."""
    inp, n, A, FS = 10, 12, 5, 4
    d = make_net(inp, n, k, A, FS, seed=3, bias_scale=0.1)
    batch = make_batch(inp, A, FS, B, seed=4)
    tr = oracle.Trainer(inp, n, k, A, FS)
    tr.set_params(d)
    tr.loss_grad(*batch)
    g = tr.get(tr.GRADS).astype(np.float64)
    flat = tr.pack(d).astype(np.float64)
    rng = np.random.default_rng(0)
    idx = rng.choice(flat.size, size=60, replace=False)
    eps = 1e-4
    for p in idx:
        f = flat.copy(); f[p] += eps
        lp = np_forward_loss(tr.unpack(f), *batch)[0]
        f[p] -= 2 * eps
        lm = np_forward_loss(tr.unpack(f), *batch)[0]
        num = (lp - lm) / (2 * eps)
        assert abs(num - g[p]) <= 2e-4 * max(1.0, abs(num)) + 2e-6, (p, num, g[p])


def test_adam_weight_decay_matches_flux_semantics():
    inp, n, k, A, FS, B = 8, 8, 1, 4, 3, 16
    d = make_net(inp, n, k, A, FS, seed=5, bias_scale=0.1)
    tr = oracle.Trainer(inp, n, k, A, FS, lr=0.001, wd=1e-4)
    tr.set_params(d)
    x = tr.pack(d).copy()
    m = np.zeros_like(x); v = np.zeros_like(x); bp = np.array([0.9, 0.999])
    for step in range(3):
        batch = make_batch(inp, A, FS, B, seed=10 + step)
        tr.loss_grad(*batch)
        g = tr.get(tr.GRADS)
        x, m, v, bp = np_adam_step(x, g, m, v, bp, lr=0.001, wd=1e-4)
        tr.apply(1.0)
        assert np.array_equal(tr.get(tr.PARAMS), x), step
        assert np.array_equal(tr.get(tr.M), m) and np.array_equal(tr.get(tr.V), v)


def test_training_reduces_the_loss():
    inp, n, k, A, FS, B = 18, 32, 2, 9, 9, 64
    d = make_net(inp, n, k, A, FS, seed=6)
    batch = make_batch(inp, A, FS, B, seed=7)
    tr = oracle.Trainer(inp, n, k, A, FS, lr=0.01)
    tr.set_params(d)
    first = tr.step(*batch)[0]
    for _ in range(30):
        last = tr.step(*batch)[0]
    assert last < 0.9 * first


def test_pack_unpack_roundtrip():
    inp, n, k, A, FS = 6, 5, 2, 3, 4
    d = make_net(inp, n, k, A, FS, seed=8, bias_scale=0.3)
    tr = oracle.Trainer(inp, n, k, A, FS)
    u = tr.unpack(tr.pack(d))
    for key in ("base", "pol_w", "val_w", "feat_w", "pol_b", "val_b", "feat_b"):
        assert np.array_equal(np.asarray(u[key]).reshape(np.asarray(d[key]).shape), d[key]), key
    assert all(np.array_equal(a, b) for a, b in zip(u["res"], d["res"]))
    assert flat_of(d).size == tr.P
