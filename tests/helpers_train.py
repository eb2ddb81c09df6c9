"""Shared pieces of the training-step tests: random nets/batches and an independent numpy float64 statement of the mathematics."""
import numpy as np


def make_net(inp, n, k, A, FS, seed=0, bias_scale=0.0):
    rng = np.random.default_rng(seed)
    glorot = lambda o, i: (rng.uniform(-1, 1, size=(o, i)) * np.sqrt(6.0 / (o + i))).astype(np.float32)
    b = lambda m: (rng.standard_normal(m) * bias_scale).astype(np.float32)
    return dict(base=glorot(n, inp), res=[glorot(n, n) for _ in range(k)], pol_w=glorot(A, n), pol_b=b(A), val_w=glorot(1, n), val_b=b(1),
                feat_w=glorot(FS, n), feat_b=b(FS))


def make_batch(inp, A, FS, B, seed=0):
    """Samples shaped like self-play output: 0/1 board planes, a policy that sums to 1, value in {0, .5, 1}, fstate in {-1,0,1}."""
    rng = np.random.default_rng(seed)
    state = (rng.random((B, inp)) < 0.3).astype(np.int8)
    pol = rng.random((B, A)).astype(np.float32) * (rng.random((B, A)) < 0.7)
    pol[:, 0] += 1e-3
    pol = (pol / pol.sum(1, keepdims=True)).astype(np.float32)
    value = rng.choice(np.array([0.0, 0.5, 1.0], np.float32), size=B)
    fstate = rng.integers(-1, 2, size=(B, FS)).astype(np.int8)
    return state, pol, value, fstate


def flat_of(d):
    heads = np.concatenate([d["pol_w"], d["val_w"].reshape(1, -1), d["feat_w"]], axis=0)
    return np.concatenate([d["base"].ravel(order="F")] + [w.ravel(order="F") for w in d["res"]] +
                          [heads.ravel(order="F"), d["pol_b"].ravel(), d["val_b"].ravel(), d["feat_b"].ravel()])


def np_forward_loss(d, state, pol, value, fstate, fweight=0.001):
    """[total, policy, value, feature] in float64."""
    f8 = lambda a: np.asarray(a, np.float64)
    x = f8(state)
    h = np.maximum(x @ f8(d["base"]).T, 0)
    for w in d["res"]:
        h = np.maximum(h + np.maximum(h @ f8(w).T, 0), 0)
    p = h @ f8(d["pol_w"]).T + f8(d["pol_b"]).ravel()
    v = 1 / (1 + np.exp(-(h @ f8(d["val_w"]).reshape(-1) + f8(d["val_b"]).ravel()[0])))
    f = np.tanh(h @ f8(d["feat_w"]).T + f8(d["feat_b"]).ravel())
    m = p.max(1, keepdims=True)
    logsm = p - m - np.log(np.exp(p - m).sum(1, keepdims=True))
    lp = np.mean(-(f8(pol) * logsm).sum(1))
    lv = np.mean((v - f8(value)) ** 2)
    lf = np.mean((f - f8(fstate)) ** 2)
    return np.array([lp + lv + fweight * lf, lp, lv, lf])


def np_adam_step(x, g, m, v, bp, lr=0.001, b1=0.9, b2=0.999, eps=1e-8, wd=1e-4):
    """Flux 0.12.6: apply!(ADAM) then apply!(WeightDecay) then x .-= delta; Float64 hyper-parameters over Float32 arrays (each
    broadcast computes in Float64 and rounds on assignment)."""
    f4, f8 = np.float32, np.float64
    m = (b1 * m.astype(f8) + (1 - b1) * g.astype(f8)).astype(f4)
    v = (b2 * v.astype(f8) + (1 - b2) * (g * g).astype(f8)).astype(f4)
    delta = (m.astype(f8) / (1 - bp[0]) / (np.sqrt(v.astype(f8) / (1 - bp[1])) + eps) * lr).astype(f4)
    delta = (delta.astype(f8) + wd * x.astype(f8)).astype(f4)
    return (x - delta).astype(f4), m, v, bp * np.array([b1, b2])
