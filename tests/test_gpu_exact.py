"""GPU parity of the TENSOR-CORE product paths against the fp32 oracle, bit for bit (-m gpu).

The tcgen05 network chain normally differs from the fp32 formula by operand rounding and accumulation order, so the per-ply
persistent kernel (`fused::ply_kernel` with its own thread-per-game descent / expansion and item-packed backup — the kernel the
headline sims/s number is measured on) could only be compared with the other CUDA path.  Here the network is an EXACT-ARITHMETIC one
(tests/helpers.py::make_exact_nets: integer activations below 2048, dyadic head weights): every operand is exactly representable in
fp16/bf16 and every partial sum in fp32, so any accumulation order in any format gives the oracle's fp32 bits, and the default
product mode (AGPU_NN_FP16_TC) must reproduce `oracle.selfplay(..., FP32)` / `oracle.Tree.mcts_single(..., FP32)` exactly:
samples, π̄, results, and whole tree tables.  Reference semantics at stake: mcts_gpu.jl:100-199 (kdescendTree!), :250-328
(expand, backUp), :376-462 (mcts_single), :477-579 (self-play loop)."""
import numpy as np
import pytest

import oracle
from conftest import GAME_SPECS
from helpers import assert_bits_equal, make_exact_nets, random_positions
from test_gpu_parity import compare_trees

pytestmark = pytest.mark.gpu
FP16_TC, BF16_TC = 2, 0


def ctx_for(name, R, L, n, k, nn_mode=FP16_TC):
    import alphagpu_b200 as ag
    g, N, nv = GAME_SPECS[name]
    return ag.Context(ag.GameSpec(g, N, nv), R, L, n, k, 0, nn_mode)


def assert_samples_equal(smp, osmp, res, ores, stats, ost, what):
    assert np.array_equal(res, ores), (what, res, ores)
    assert stats["faults"] == 0 and ost["faults"] == 0, what
    for key in ("sims", "positions", "plies", "total_length"):
        assert stats[key] == ost[key], (what, key)
    n = osmp.count
    assert len(smp["player"]) == n, what
    assert np.array_equal(smp["state"], osmp.state[:n]), what
    assert np.array_equal(smp["player"], osmp.player[:n]), what
    assert np.array_equal(smp["game"], osmp.game[:n]) and np.array_equal(smp["ply"], osmp.ply[:n]), what
    assert_bits_equal(smp["policy"], osmp.policy[:n], f"{what}: sample policy")
    assert_bits_equal(smp["value"], osmp.value[:n], f"{what}: sample value")
    assert np.array_equal(smp["fstate"], osmp.fstate[:n]), what


# games per CTA forced through AGPU_FUSED_MIN_GPC so that every variant of the per-ply kernel and both sides of every boundary run
# at a size the oracle finishes in a second: <= 32 (swapped, N = 32), 33..64 (swapped, N = 64), 65..128 (one tile, ordinary),
# 129..256 (two tiles); the last CTA of each grid holds a ragged remainder.
VARIANT_CASES = [(8, 150), (32, 100), (40, 100), (64, 200), (72, 200), (128, 300), (136, 300), (256, 600)]


@pytest.mark.parametrize("min_gpc,games", VARIANT_CASES)
def test_fused_ply_kernel_selfplay_bit_exact_vs_oracle(min_gpc, games, monkeypatch):
    """fused::ply_kernel (default product path, fp16 tcgen05 chain) == oracle self-play with the fp32 network, every kernel variant."""
    monkeypatch.setenv("AGPU_FUSED_MIN_GPC", str(min_gpc))
    name, R = "connect4", 64
    ospec = oracle.Spec(*GAME_SPECS[name])
    pnet, onet = make_exact_nets(GAME_SPECS[name], 128, 6, seed=11)
    ctx = ctx_for(name, R, games, 128, 6)
    ctx.set_weights(pnet)
    res, stats, smp = ctx.selfplay(R, games, cpuct=1.5, seed=2024, uid_base=77)
    ctx.close()
    assert stats["kernel_launches"] < 8 * stats["plies"] + 8, "the per-ply persistent kernel did not run (one search launch per ply expected)"
    osmp = oracle.Samples(ospec, games * ospec.maxLen)
    ores, ost = oracle.selfplay(ospec, onet, R, games, cpuct=1.5, seed=2024, uid_base=77, samples=osmp, nn_mode=oracle.Net.FP32)
    assert_samples_equal(smp, osmp, res, ores, stats, ost, f"min_gpc={min_gpc}")


@pytest.mark.parametrize("nn_mode", [FP16_TC, BF16_TC])
def test_fused_ply_kernel_config2_full_size_bit_exact_vs_oracle(nn_mode):
    """BASELINE config 2 at full size — Connect4, DenseNet 128x6, 64 rollouts, 32768 games — through the benchmarked kernel, held to
    the oracle's fp32 self-play bit for bit (all 580 k samples).  bf16 operands too: the exact net's activations stay <= 192."""
    name, R, games = "connect4", 64, 32768
    if nn_mode == BF16_TC:
        games = 4096
    ospec = oracle.Spec(*GAME_SPECS[name])
    pnet, onet = make_exact_nets(GAME_SPECS[name], 128, 6, seed=5)
    ctx = ctx_for(name, R, games, 128, 6, nn_mode)
    ctx.set_weights(pnet)
    res, stats, smp = ctx.selfplay(R, games, cpuct=1.5, seed=1, uid_base=0)
    ctx.close()
    osmp = oracle.Samples(ospec, games * ospec.maxLen)
    ores, ost = oracle.selfplay(ospec, onet, R, games, cpuct=1.5, seed=1, uid_base=0, samples=osmp, nn_mode=oracle.Net.FP32)
    assert_samples_equal(smp, osmp, res, ores, stats, ost, f"config 2, nn_mode {nn_mode}")


def test_fused_ply_kernel_config1_ttt_bit_exact_vs_oracle():
    """BASELINE config 1 (Gobang 3x3/3, 128x6, 64 rollouts, 1024 games) through the per-ply kernel (A = 9, 16-wide records)."""
    name, R, games = "ttt", 64, 1024
    ospec = oracle.Spec(*GAME_SPECS[name])
    pnet, onet = make_exact_nets(GAME_SPECS[name], 128, 6, seed=3)
    ctx = ctx_for(name, R, games, 128, 6)
    ctx.set_weights(pnet)
    res, stats, smp = ctx.selfplay(R, games, cpuct=1.5, seed=9)
    ctx.close()
    osmp = oracle.Samples(ospec, games * ospec.maxLen)
    ores, ost = oracle.selfplay(ospec, onet, R, games, cpuct=1.5, seed=9, samples=osmp, nn_mode=oracle.Net.FP32)
    assert_samples_equal(smp, osmp, res, ores, stats, ost, "config 1")


@pytest.mark.parametrize("min_gpc,L", [(8, 200), (64, 200), (128, 300), (136, 300)])
@pytest.mark.parametrize("training", [True, False])
def test_fused_search_tree_tables_bit_exact_vs_oracle(min_gpc, L, training, monkeypatch):
    """agpu_search (== mcts_single) through the per-ply kernel + agpu_get_tree: parent / action / child / order / expanded / states and
    the floats prior, q, visits of EVERY node, plus policy_final, equal the oracle's after a 64-rollout search from mid-game positions."""
    monkeypatch.setenv("AGPU_FUSED_MIN_GPC", str(min_gpc))
    name, R = "connect4", 64
    ospec = oracle.Spec(*GAME_SPECS[name])
    pnet, onet = make_exact_nets(GAME_SPECS[name], 128, 6, seed=21)
    pos = random_positions(ospec, L, seed=5, max_plies=30)
    uids = (np.arange(L) * 5 + 1).astype(np.uint32)
    ctx = ctx_for(name, R, L, 128, 6)
    ctx.set_weights(pnet)
    ctx.re_init(pos, uids)
    ctx.mcts_single(R, training=training, cpuct=1.5, seed=0xABCDEF0123, ply=7)
    t = oracle.Tree(ospec, R, L)
    t.reinit(pos, uids)
    t.mcts_single(onet, R, training, 1.5, seed=0xABCDEF0123, ply=7, nn_mode=oracle.Net.FP32)
    compare_trees(ctx.tree(), t.dump(), f"min_gpc={min_gpc}")
    pol, batch = ctx.roots()
    opol, obatch = t.roots()
    assert_bits_equal(pol, opol, "policy_final")
    assert np.array_equal(batch, obatch)
    ctx.close()


@pytest.mark.parametrize("name,n,k,R,L", [("hex7", 512, 8, 64, 300), ("gobang9", 512, 8, 128, 160), ("reversi8", 512, 8, 64, 200),
                                          ("hex5", 128, 4, 32, 300), ("gobang5", 128, 2, 32, 300), ("reversi6", 128, 2, 32, 200)])
def test_large_board_tc_search_bit_exact_vs_oracle(name, n, k, R, L):
    """Configs 3-5 (Hex 7, Gobang 9, Reversi 8 with 512x8 nets) and the width-128 chain on mid-size boards: the tensor-core search
    path (tc_mlp512 / tc_mlp128 + the search kernels for large action sets) vs the fp32 oracle, whole tree tables."""
    ospec = oracle.Spec(*GAME_SPECS[name])
    pnet, onet = make_exact_nets(GAME_SPECS[name], n, k, seed=8)
    pos = random_positions(ospec, L, seed=6, max_plies=max(2, ospec.maxLen // 2))
    ctx = ctx_for(name, R, L, n, k)
    ctx.set_weights(pnet)
    ctx.re_init(pos)
    ctx.mcts_single(R, training=True, cpuct=1.5, seed=99, ply=3)
    t = oracle.Tree(ospec, R, L)
    t.reinit(pos)
    t.mcts_single(onet, R, True, 1.5, seed=99, ply=3, nn_mode=oracle.Net.FP32)
    compare_trees(ctx.tree(), t.dump(), name)
    assert_bits_equal(ctx.roots()[0], t.roots()[0], "policy_final")
    ctx.close()


@pytest.mark.parametrize("name,n,k,R,games", [("hex7", 512, 8, 64, 96), ("gobang9", 512, 8, 128, 48), ("reversi8", 512, 8, 64, 64)])
def test_large_board_tc_selfplay_bit_exact_vs_oracle(name, n, k, R, games):
    """Whole self-play generations of configs 3-5 (reduced game counts) with the tensor-core network, bit-identical to the oracle."""
    ospec = oracle.Spec(*GAME_SPECS[name])
    pnet, onet = make_exact_nets(GAME_SPECS[name], n, k, seed=4)
    ctx = ctx_for(name, R, games, n, k)
    ctx.set_weights(pnet)
    res, stats, smp = ctx.selfplay(R, games, cpuct=1.5, seed=31)
    ctx.close()
    osmp = oracle.Samples(ospec, games * ospec.maxLen)
    ores, ost = oracle.selfplay(ospec, onet, R, games, cpuct=1.5, seed=31, samples=osmp, nn_mode=oracle.Net.FP32)
    assert_samples_equal(smp, osmp, res, ores, stats, ost, name)


def test_duel_tc_exact_nets_bit_exact_vs_oracle():
    """The duel loop (mcts_gpu.jl:581-651) through the per-ply kernel with two exact nets: the tally equals the oracle's."""
    ospec = oracle.Spec(*GAME_SPECS["connect4"])
    p1, o1 = make_exact_nets(GAME_SPECS["connect4"], 128, 6, seed=1)
    p2, o2 = make_exact_nets(GAME_SPECS["connect4"], 128, 6, seed=2)
    ctx = ctx_for("connect4", 32, 1024, 128, 6)
    ctx.set_weights(p1, 0)
    ctx.set_weights(p2, 1)
    res, st = ctx.duel(32, 1024, cpuct=2.0, seed=9)
    ores, ost = oracle.duel(ospec, o1, o2, 32, 1024, cpuct=2.0, seed=9, nn_mode=oracle.Net.FP32)
    assert np.array_equal(res, ores), (res, ores)
    assert st["positions"] == ost["positions"] and st["plies"] == ost["plies"]
    ctx.close()
