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
