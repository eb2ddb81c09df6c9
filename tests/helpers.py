"""Shared test helpers: random nets in both front ends, random reachable positions, comparison utilities."""
import numpy as np

import oracle
from conftest import GAME_SPECS

f32 = np.float32


def make_nets(spec_t, n, k, seed):
    """The same random-init weights as an alphagpu_b200.SNetwork2 (product) and an oracle.Net (checker)."""
    import alphagpu_b200 as ag
    ospec = oracle.Spec(*spec_t)
    pnet = ag.ressimplesf(2 * ospec.VS, ospec.A, n, k, seed=seed)
    # non-zero head biases so that the bias path is exercised
    rng = np.random.default_rng(seed + 1000)
    pnet.policy_bias = rng.uniform(-0.2, 0.2, ospec.A).astype(f32)
    pnet.value_bias = rng.uniform(-0.2, 0.2, 1).astype(f32)
    onet = oracle.Net(pnet.base, pnet.res, pnet.policy, pnet.policy_bias, pnet.value, pnet.value_bias)
    return pnet, onet


def random_positions(ospec, n, seed, max_plies=None):
    """n positions reached by uniformly random legal play from Position() (not over), via the oracle."""
    rng = np.random.default_rng(seed)
    out = np.zeros(n, ospec.pos_dtype)
    filled = 0
    while filled < n:
        pos = ospec.position(1)
        depth = int(rng.integers(0, (max_plies or ospec.maxLen)))
        ok = True
        for _ in range(depth):
            legal = np.nonzero(ospec.legal(pos)[0])[0]
            if len(legal) == 0:
                ok = False
                break
            nxt = ospec.play(pos, int(rng.choice(legal)) + 1)
            if ospec.is_over(nxt)[0][0]:
                break
            pos = nxt
        if ok and not ospec.is_over(pos)[0][0]:
            out[filled] = pos[0]
            filled += 1
    return out


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_bits_equal(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if a.dtype.kind == "f":
        bad = np.nonzero(bits(a) != bits(b))
        if len(bad[0]):
            i = tuple(x[0] for x in bad)
            raise AssertionError(f"{what}: {len(bad[0])} of {a.size} float32 values differ in bits; first at {i}: {a[i]!r} vs {b[i]!r}")
    else:
        assert np.array_equal(a, b), what


def make_exact_nets(spec_t, n, k, seed, head_shift=4, npos=3):
    """An EXACT-ARITHMETIC network: trunk weights in {-1, 0, +1} and so sparse that every activation is a non-negative integer
    <= npos * 2^k (< 2048 for 128x6 and 512x8, so exactly representable as an fp16 MMA operand; <= 256 with npos small enough for bf16),
    head weights +-2^-head_shift, biases multiples of 2^-head_shift.  Every product and every partial sum of every dot product is then
    exactly representable in fp32, so ANY accumulation order and ANY operand format (fp32 CUDA cores, fp16/bf16 tensor cores with the
    fp32 or the 16-bit residual stream) yields bit-identical logits and values: the tensor-core kernels can be held to the fp32 oracle
    bit for bit.  Rows: base = npos entries +1 and two -1 (b0 <= npos); residual blocks = one +1 and two -1 per row, so
    relu(W b) <= max b and the stream at most doubles per block."""
    import alphagpu_b200 as ag
    ospec = oracle.Spec(*spec_t)
    rng = np.random.default_rng(seed)
    inp, A = 2 * ospec.VS, ospec.A
    base = np.zeros((n, inp), f32)
    for o in range(n):
        idx = rng.choice(inp, npos + 2, replace=False)
        base[o, idx[:npos]] = 1.0
        base[o, idx[npos:]] = -1.0
    res = []
    for _ in range(k):
        w = np.zeros((n, n), f32)
        for o in range(n):
            idx = rng.choice(n, 3, replace=False)
            w[o, idx[0]] = 1.0
            w[o, idx[1:1 + int(rng.integers(1, 3))]] = -1.0
        res.append(w)
    s = f32(2.0 ** -head_shift)
    pol = np.zeros((A, n), f32)
    for a in range(A):
        idx = rng.choice(n, 12, replace=False)
        pol[a, idx] = rng.choice([-1.0, 1.0], 12).astype(f32) * s
    val = np.zeros((1, n), f32)
    idx = rng.choice(n, 12, replace=False)
    val[0, idx] = rng.choice([-1.0, 1.0], 12).astype(f32) * s
    pb = (rng.integers(-8, 9, A).astype(f32) * s).astype(f32)
    vb = (rng.integers(-8, 9, 1).astype(f32) * s).astype(f32)
    pnet = ag.SNetwork2(base, res, pol, pb, val, vb)
    onet = oracle.Net(pnet.base, pnet.res, pnet.policy, pnet.policy_bias, pnet.value, pnet.value_bias)
    return pnet, onet
