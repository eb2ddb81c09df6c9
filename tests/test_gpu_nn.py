"""GPU tests of the tcgen05/TMEM network chain — the one floating-point kernel with a tolerance.

Two operand formats run the same kernel: fp16 (AGPU_NN_FP16_TC, the default product mode) and bf16
(AGPU_NN_BF16_TC).  Each is checked against (a) the oracle's operand-faithful mode (same roundings of weights and
activations, fp32 accumulate: only the accumulation order inside the tensor core differs) and (b) the fp32
reference formula (DenseNet.jl:294-304).  North-star tolerance for policies and values: 1e-3.  Measured on B200
(profiles/r01_nn_accuracy.txt): fp16 operands — mean |Δp| 1e-4, 99.9th percentile < 1e-3, worst single entry
1.3e-3 (Connect4 128x6); bf16 operands (8 significand bits) — worst 1e-2.  The bounds below are those measurements
with head-room; the exact-fp32 evaluator (AGPU_NN_FP32) is the bit-exact parity mode (tests/test_gpu_parity.py)."""
import numpy as np
import pytest

import oracle
from conftest import GAME_SPECS
from helpers import make_nets, random_positions

pytestmark = pytest.mark.gpu
f32 = np.float32
TOL_P999 = {2: 1e-3, 0: 8e-3}       # nn_mode -> 99.9th percentile of |softmax - fp32 softmax| and |value - fp32 value|
TOL_MAX = {2: 2e-3, 0: 2e-2}        # nn_mode -> worst single entry
TOL_MEAN = {2: 2e-4, 0: 1.5e-3}
ORACLE_MODE = {2: oracle.Net.F16, 0: oracle.Net.BF16}


def ctx_for(name, R, L, n, k, nn_mode=2):
    import alphagpu_b200 as ag
    g, N, nv = GAME_SPECS[name]
    return ag.Context(ag.GameSpec(g, N, nv), R, L, n, k, 0, nn_mode)


@pytest.mark.parametrize("nn_mode", [2, 0])
@pytest.mark.parametrize("name,n,k,L", [("connect4", 128, 6, 1000), ("ttt", 128, 6, 300), ("hex7", 128, 4, 515), ("reversi8", 128, 2, 256),
                                        ("connect4", 128, 0, 77), ("hex5", 128, 1, 129)])
def test_tc_forward_matches_oracle(name, n, k, L, nn_mode):
    ospec = oracle.Spec(*GAME_SPECS[name])
    pnet, onet = make_nets(GAME_SPECS[name], n, k, seed=3)
    ctx = ctx_for(name, 4, 8, n, k, nn_mode)
    ctx.set_weights(pnet)
    x = ospec.encode(random_positions(ospec, L, seed=9))
    logits, v = ctx.forward(x)
    bl, bv = onet.forward(x, mode=ORACLE_MODE[nn_mode])
    fl, fv = onet.forward(x, mode=oracle.Net.FP32)
    d_b = float(np.abs(logits - bl).max())
    d_f = float(np.abs(logits - fl).max())
    p, pb, pf = oracle.softmax(logits), oracle.softmax(bl), oracle.softmax(fl)
    print(f"\n{name} {n}x{k} mode {nn_mode}: |logits - faithful oracle| {d_b:.2e}  |logits - fp32 oracle| {d_f:.2e}  "
          f"|softmax - fp32| {np.abs(p - pf).max():.2e}  |value - fp32| {np.abs(v - fv).max():.2e}")
    med_b = float(np.median(np.abs(logits - bl)))
    print(f"   median |logits - bf16 oracle| {med_b:.2e}   |softmax - bf16 oracle| {np.abs(p - pb).max():.2e}")
    # Same operand roundings, different fp32 accumulation order inside the tensor core.  A last-bit difference in an
    # accumulator can flip the bf16 rounding of one activation (a 2^-9 relative step) and that propagates, so the
    # bound is "typically ~1e-6, never more than a few bf16 steps":
    assert med_b < 5e-5, med_b
    assert d_b < 2e-2, d_b
    assert np.abs(v - bv).max() < 5e-3
    assert np.abs(p - pb).max() < 5e-3
    assert np.quantile(np.abs(p - pb), 0.99) < 2e-4
    # against the fp32 formula (DenseNet.jl:294-304): north-star tolerance 1e-3 on policies and values
    print(f"   vs fp32: softmax max {np.abs(p - pf).max():.2e} mean {np.abs(p - pf).mean():.2e}; value max {np.abs(v - fv).max():.2e}")
    for d in (np.abs(p - pf), np.abs(v - fv)):
        assert np.quantile(d, 0.999) < TOL_P999[nn_mode], np.quantile(d, 0.999)
        assert d.max() < TOL_MAX[nn_mode], d.max()
        assert d.mean() < TOL_MEAN[nn_mode], d.mean()
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
    for mode in (2, 1):
        ctx = ctx_for(name, R, L, 128, 6, nn_mode=mode)
        ctx.set_weights(pnet)
        ctx.re_init(pos)
        ctx.mcts_single(R, training=True, cpuct=1.5, seed=3)
        pols.append(ctx.roots()[0])
        ctx.close()
    d = np.abs(pols[0] - pols[1]).max(1)
    print("\nroot policy |f16 TC - fp32|: median %.2e  90%% %.2e  max %.2e  frac<1e-3 %.3f" % (np.median(d), np.quantile(d, 0.9), d.max(), (d < 1e-3).mean()))
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


@pytest.mark.parametrize("name,games,R", [("connect4", 3000, 24), ("ttt", 1500, 16), ("connect4", 40000, 8),
                                          ("connect4", 1, 1), ("connect4", 257, 2), ("connect4", 300, 255), ("ttt", 33, 200),
                                          ("connect4", 148 * 64 + 1, 3), ("connect4", 148 * 128 + 9, 3)])
def test_fused_ply_kernel_equals_separate_kernels(monkeypatch, name, games, R):
    """The persistent per-ply kernel (fused.cuh) and the per-rollout kernels (search.cuh + nn_tc.cu) run the same device functions
    and the same MMA sequence: a whole self-play generation must come out identical, bit for bit.  Game counts that are not
    multiples of 256 exercise partially filled tiles and CTAs; 40000 games exceed one CTA per SM (full 256-game CTAs); one game with
    one rollout, and the maximum rollout count (node ids are bytes) with paths deeper than the 16 levels kept in shared memory; the last
    two straddle the kernel-variant boundaries at 64 and 128 games per CTA (swapped / one tile / two tiles)."""
    pnet, _ = make_nets(GAME_SPECS[name], 128, 6, seed=11)
    outs = []
    for fused in ("1", "0"):
        monkeypatch.setenv("AGPU_FUSED", fused)
        ctx = ctx_for(name, R, games, 128, 6, nn_mode=2)
        ctx.set_weights(pnet)
        res, st, smp = ctx.selfplay(R, games, cpuct=1.5, seed=21)
        outs.append((res, st, smp))
        ctx.close()
    (r1, s1, a), (r0, s0, b) = outs
    assert np.array_equal(r1, r0) and s1["positions"] == s0["positions"] and s1["faults"] == 0
    if R >= 8:
        assert s1["kernel_launches"] < s0["kernel_launches"] / 10
    for k in a:
        if a[k].dtype.kind == "f":
            assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), k
        else:
            assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("nn_mode", [2, 0])
@pytest.mark.parametrize("name,k,L", [("hex7", 8, 300), ("gobang9", 8, 200), ("reversi8", 8, 260), ("hex7", 0, 130), ("gobang9", 1, 64)])
def test_tc512_forward_matches_oracle(name, k, L, nn_mode):
    """Width-512 chain (nn_tc512.cu): activations/residual in 16-bit shared memory -> oracle's *_RESID modes; vs the fp32 formula."""
    n = 512
    ospec = oracle.Spec(*GAME_SPECS[name])
    pnet, onet = make_nets(GAME_SPECS[name], n, k, seed=3)
    ctx = ctx_for(name, 4, 8, n, k, nn_mode)
    ctx.set_weights(pnet)
    x = ospec.encode(random_positions(ospec, L, seed=9))
    logits, v = ctx.forward(x)
    bl, bv = onet.forward(x, mode={2: oracle.Net.F16_RESID, 0: oracle.Net.BF16_RESID}[nn_mode])
    fl, fv = onet.forward(x, mode=oracle.Net.FP32)
    p, pb, pf = oracle.softmax(logits), oracle.softmax(bl), oracle.softmax(fl)
    dp = np.abs(p - pf)
    print(f"\n{name} 512x{k} mode {nn_mode}: |logits - faithful| max {np.abs(logits - bl).max():.2e} median {np.median(np.abs(logits - bl)):.2e}; "
          f"vs fp32: softmax max {dp.max():.2e} p99.9 {np.quantile(dp, 0.999):.2e} mean {dp.mean():.2e}; value max {np.abs(v - fv).max():.2e}")
    # The residual stream itself is re-rounded to 16 bits every layer here (TMEM is full of accumulators), so a last-bit difference
    # in an accumulator moves a residual entry by one 16-bit ulp for good: the match with the faithful oracle is "within a few
    # ulps of the 16-bit format", and the distance to the fp32 formula is larger than for width 128 (fp32 residual).
    tol_faithful = {2: 3e-3, 0: 3e-2}[nn_mode]
    assert np.median(np.abs(logits - bl)) < tol_faithful
    assert np.quantile(np.abs(p - pb), 0.99) < tol_faithful
    assert np.quantile(dp, 0.999) < 4 * TOL_P999[nn_mode] and dp.mean() < TOL_MEAN[nn_mode]
    assert np.quantile(np.abs(v - fv), 0.999) < 4 * TOL_P999[nn_mode]
    ctx.close()


def test_tc512_search_runs_config3_shape():
    """Config 3 shape through the tensor-core path (segmented graph replay): legal, deterministic, policies normalised."""
    name = "hex7"
    pnet, _ = make_nets(GAME_SPECS[name], 512, 8, seed=5)
    games, R = 700, 16
    ctx = ctx_for(name, R, games, 512, 8, nn_mode=2)
    ctx.set_weights(pnet)
    res, st, smp = ctx.selfplay(R, games, cpuct=1.5, seed=2)
    res2, st2, smp2 = ctx.selfplay(R, games, cpuct=1.5, seed=2)
    assert res.sum() == games and st["faults"] == 0
    assert np.array_equal(res, res2) and all(np.array_equal(smp[k], smp2[k]) for k in smp)
    assert np.all(np.abs(smp["policy"].sum(1) - 1) < 2e-3)
    ctx.close()


# ---- the north-star bar itself: policies and values within 1e-3 of the fp32 formula (BASELINE.json north_star; DenseNet.jl:294-304) ----
# Asserted where the tensor-core chain meets it; where it does not, the case is an xfail(strict): the miss is a measured, documented
# property of the operand format and a silent change in either direction fails the suite.
_MISS = lambda why: pytest.mark.xfail(strict=True, reason=why)
NORTH_STAR_CASES = [
    # width 128: fp32 residual stream, 16-bit MMA operands
    pytest.param("connect4", 128, 6, 1000, 2, marks=_MISS("fp16 operands, 6 blocks: worst policy entry 1.26e-3 (p99.9 < 1e-3)")),
    pytest.param("ttt", 128, 6, 300, 2),
    pytest.param("hex7", 128, 4, 515, 2),
    pytest.param("reversi8", 128, 2, 256, 2),
    pytest.param("connect4", 128, 0, 77, 2),
    pytest.param("connect4", 128, 0, 77, 0),
    pytest.param("connect4", 128, 6, 1000, 0, marks=_MISS("bf16 operands (8 significand bits): worst entry 1e-2")),
    pytest.param("ttt", 128, 6, 300, 0, marks=_MISS("bf16 operands: worst entry 8e-3")),
    pytest.param("hex7", 128, 4, 515, 0, marks=_MISS("bf16 operands: worst entry 6e-3")),
    # width 512: the residual stream itself is 16-bit (all 512 TMEM columns are accumulators)
    pytest.param("gobang9", 512, 8, 200, 2),
    pytest.param("hex7", 512, 0, 130, 2),
    pytest.param("gobang9", 512, 1, 64, 2),
    pytest.param("hex7", 512, 8, 300, 2, marks=_MISS("fp16 16-bit residual stream, 8 blocks: worst policy entry 2.7e-3, p99.9 1.9e-3")),
    pytest.param("reversi8", 512, 8, 260, 2, marks=_MISS("fp16 16-bit residual stream, 8 blocks: worst policy entry 2.4e-3, p99.9 1.4e-3")),
    pytest.param("hex7", 512, 8, 300, 0, marks=_MISS("bf16 16-bit residual stream: worst entry 2e-2")),
]


@pytest.mark.parametrize("name,n,k,L,nn_mode", NORTH_STAR_CASES)
def test_tc_chain_meets_the_1e3_tolerance(name, n, k, L, nn_mode):
    ospec = oracle.Spec(*GAME_SPECS[name])
    pnet, onet = make_nets(GAME_SPECS[name], n, k, seed=3)
    ctx = ctx_for(name, 4, 8, n, k, nn_mode)
    ctx.set_weights(pnet)
    x = ospec.encode(random_positions(ospec, L, seed=9))
    logits, v = ctx.forward(x)
    ctx.close()
    fl, fv = onet.forward(x, mode=oracle.Net.FP32)
    dp, dv = np.abs(oracle.softmax(logits) - oracle.softmax(fl)), np.abs(v - fv)
    print(f"\n{name} {n}x{k} mode {nn_mode}: |Δpolicy| max {dp.max():.2e} p99.9 {np.quantile(dp, 0.999):.2e}; |Δvalue| max {dv.max():.2e}")
    assert dp.max() <= 1e-3 and dv.max() <= 1e-3
