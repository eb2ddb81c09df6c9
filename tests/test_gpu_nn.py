"""GPU tests of the bf16 tcgen05/TMEM network chain (AGPU_NN_BF16_TC) — the one floating-point kernel with a
tolerance.  Checked against (a) the oracle's bf16-faithful mode (same operand roundings, fp32 accumulate:
only the accumulation order differs) and (b) the fp32 reference formula (DenseNet.jl:294-304); north-star
tolerance for policies and values: 1e-3."""
import numpy as np
import pytest

import oracle
from conftest import GAME_SPECS
from helpers import make_nets, random_positions

pytestmark = pytest.mark.gpu
f32 = np.float32


def ctx_for(name, R, L, n, k, nn_mode=0):
    import alphagpu_b200 as ag
    g, N, nv = GAME_SPECS[name]
    return ag.Context(ag.GameSpec(g, N, nv), R, L, n, k, 0, nn_mode)


@pytest.mark.parametrize("name,n,k,L", [("connect4", 128, 6, 1000), ("ttt", 128, 6, 300), ("hex7", 128, 4, 515), ("reversi8", 128, 2, 256),
                                        ("connect4", 128, 0, 77), ("hex5", 128, 1, 129)])
def test_tc_forward_matches_oracle(name, n, k, L):
    ospec = oracle.Spec(*GAME_SPECS[name])
    pnet, onet = make_nets(GAME_SPECS[name], n, k, seed=3)
    ctx = ctx_for(name, 4, 8, n, k)
    ctx.set_weights(pnet)
    x = ospec.encode(random_positions(ospec, L, seed=9))
    logits, v = ctx.forward(x)
    bl, bv = onet.forward(x, mode=oracle.Net.BF16)
    fl, fv = onet.forward(x, mode=oracle.Net.FP32)
    d_b = float(np.abs(logits - bl).max())
    d_f = float(np.abs(logits - fl).max())
    p, pb, pf = oracle.softmax(logits), oracle.softmax(bl), oracle.softmax(fl)
    print(f"\n{name} {n}x{k}: |logits - bf16 oracle| {d_b:.2e}  |logits - fp32 oracle| {d_f:.2e}  "
          f"|softmax - fp32| {np.abs(p - pf).max():.2e}  |value - fp32| {np.abs(v - fv).max():.2e}")
    # same roundings, different fp32 accumulation order inside the tensor core
    assert d_b < 2e-4, d_b
    assert np.abs(v - bv).max() < 1e-4
    assert np.abs(p - pb).max() < 1e-4
    # against the fp32 formula: the north-star tolerance
    assert np.abs(p - pf).max() < 1e-3
    assert np.abs(v - fv).max() < 1e-3
    ctx.close()


def test_tc_leaf_path_equals_direct_path():
    """The kernel's own bitboard->bf16 encoder (leaf states read from the tree) gives the same outputs as feeding the encoded batch."""
    name = "connect4"
    ospec = oracle.Spec(*GAME_SPECS[name])
    pnet, _ = make_nets(GAME_SPECS[name], 128, 6, seed=5)
    L = 700
    ctx = ctx_for(name, 8, L, 128, 6)
    ctx.set_weights(pnet)
    pos = random_positions(ospec, L, seed=2)
    ctx.re_init(pos)
    ctx.search_begin()
    ctx.select(0, 1.5)
    logits, v = ctx.eval()
    l2, v2 = ctx.forward(ospec.encode(pos))
    assert np.array_equal(logits, l2) and np.array_equal(v, v2)
    ctx.close()


def test_tc_search_close_to_fp32_search():
    """One full search with the tensor-core net: root policies stay within the tolerance of the fp32-net search for
    the large majority of games (a handful may branch differently once a sampled action flips)."""
    name = "connect4"
    ospec = oracle.Spec(*GAME_SPECS[name])
    pnet, _ = make_nets(GAME_SPECS[name], 128, 6, seed=7)
    L, R = 512, 32
    pos = random_positions(ospec, L, seed=4, max_plies=12)
    pols = []
    for mode in (0, 1):
        ctx = ctx_for(name, R, L, 128, 6, nn_mode=mode)
        ctx.set_weights(pnet)
        ctx.re_init(pos)
        ctx.mcts_single(R, training=True, cpuct=1.5, seed=3)
        pols.append(ctx.roots()[0])
        ctx.close()
    d = np.abs(pols[0] - pols[1]).max(1)
    print("\nroot policy |bf16 - fp32|: median %.2e  90%% %.2e  max %.2e  frac<1e-3 %.3f" % (np.median(d), np.quantile(d, 0.9), d.max(), (d < 1e-3).mean()))
    assert np.median(d) < 1e-3
    assert np.all(np.abs(pols[0].sum(1) - 1) < 2e-3)
    assert np.array_equal(pols[0] > 0, ospec.legal(pos))


def test_tc_selfplay_properties_full_size():
    """BASELINE config 2 at full size through the product path: every game ends, no illegal move, samples well-formed,
    deterministic under a fixed seed."""
    name = "connect4"
    pnet, _ = make_nets(GAME_SPECS[name], 128, 6, seed=0)
    games, R = 32768, 64
    ctx = ctx_for(name, R, games, 128, 6)
    ctx.set_weights(pnet)
    res, st, smp = ctx.selfplay(R, games, cpuct=1.5, seed=1)
    assert res.sum() == games and st["faults"] == 0
    assert st["sims"] == st["positions"] * R and len(smp["player"]) == st["positions"]
    assert set(np.unique(smp["value"])) <= {0.0, 0.5, 1.0}
    assert np.all(np.abs(smp["policy"].sum(1) - 1) < 2e-3)
    assert set(np.unique(smp["state"])) <= {0, 1} and set(np.unique(smp["fstate"])) <= {-1, 1}
    # stones on the board == ply, from the encoding
    assert np.array_equal(smp["state"].sum(1), smp["ply"])
    res2, st2, smp2 = ctx.selfplay(R, games, cpuct=1.5, seed=1, want_samples=True)
    assert np.array_equal(res, res2) and all(np.array_equal(smp[k], smp2[k]) for k in smp)
    ctx.close()
