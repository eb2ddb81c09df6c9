"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs.  Bar: bit-exact for positions, legality, outcomes, tree tables AND for every float the search
produces (q, visits, prior, π̄), because oracle and kernels evaluate the same IEEE-fp32 operations in the
same order; the bf16 tensor-core network is the one place with a tolerance (tests/test_gpu_nn.py)."""
import numpy as np
import pytest

import oracle
from conftest import GAME_SPECS
from helpers import assert_bits_equal, make_nets, random_positions

pytestmark = pytest.mark.gpu
f32 = np.float32

ALL = ["connect4", "ttt", "gobang5", "gobang9", "hex5", "hex7", "reversi8", "reversi6"]


def ctx_for(name, R, L, n=128, k=2, nn_mode=1):
    import alphagpu_b200 as ag
    g, N, nv = GAME_SPECS[name]
    return ag.Context(ag.GameSpec(g, N, nv), R, L, n, k, 0, nn_mode)


@pytest.mark.parametrize("name", ALL)
def test_plugin_surface_bit_exact(name):
    ospec = oracle.Spec(*GAME_SPECS[name])
    ctx = ctx_for(name, 4, 8)
    assert np.array_equal(ctx.Position(3).tobytes(), ospec.position(3).tobytes())
    pos = random_positions(ospec, 300, seed=5)
    assert np.array_equal(ctx.canPlay(pos), ospec.legal(pos))
    over, res = ctx.isOver(pos)
    oover, ores = ospec.is_over(pos)
    assert np.array_equal(over, oover)
    assert np.array_equal(ctx.encode(pos), ospec.encode(pos))
    rng = np.random.default_rng(1)
    legal = ospec.legal(pos)
    acts = np.array([int(rng.choice(np.nonzero(l)[0])) + 1 for l in legal], np.int32)
    nxt, onxt = ctx.play(pos, acts), ospec.play(pos, acts)
    assert nxt.tobytes() == onxt.tobytes()
    over, res = ctx.isOver(nxt)
    oover, ores = ospec.is_over(onxt)
    assert np.array_equal(over, oover) and np.array_equal(res[over], ores[oover])
    assert over.any() or name in ("gobang9",)      # random play reaches terminal positions
    assert np.array_equal(ctx.canPlay(nxt), ospec.legal(onxt))
    ctx.close()


def test_canonical_exp_bit_exact():
    ctx = ctx_for("connect4", 4, 8)
    rng = np.random.default_rng(0)
    x = np.concatenate([-rng.uniform(0, 90, 100000), rng.uniform(-3, 3, 20000), [0.0, -87.0, -87.5, -1e-9, -30.0]]).astype(f32)
    assert_bits_equal(ctx.debug_expf(x), oracle.c_expf(x), "c_expf")
    assert_bits_equal(ctx.debug_expf(x, sigmoid=True), oracle.sigmoid(x), "sigmoid")
    ctx.close()


@pytest.mark.parametrize("name,n,k", [("connect4", 128, 6), ("ttt", 128, 6), ("hex7", 512, 2), ("reversi8", 256, 1)])
def test_forward_fp32_bit_exact(name, n, k):
    """AGPU_NN_FP32 evaluates DenseNet.jl:294-304 in the oracle's operation order: logits and σ(value) agree in every bit."""
    ospec = oracle.Spec(*GAME_SPECS[name])
    pnet, onet = make_nets(GAME_SPECS[name], n, k, seed=3)
    ctx = ctx_for(name, 4, 8, n, k)
    ctx.set_weights(pnet)
    x = ospec.encode(random_positions(ospec, 61, seed=9))
    logits, v = ctx.forward(x)
    ol, ov = onet.forward(x, mode=oracle.Net.FP32)
    assert_bits_equal(logits, ol, "logits")
    assert_bits_equal(v, ov, "value")
    ctx.close()


def compare_trees(d, od, what):
    for key in ("nnodes", "parent", "action", "child", "order", "nchild", "expanded"):
        assert np.array_equal(d[key], od[key]), f"{what}: {key} differs"
    assert d["states"].tobytes() == od["states"].tobytes(), f"{what}: states differ"
    for key in ("prior", "q", "visits"):
        assert_bits_equal(d[key], od[key], f"{what}: {key}")


@pytest.mark.parametrize("name", ALL)
@pytest.mark.parametrize("training", [True, False])
def test_search_injected_streams_bit_exact(name, training):
    """kdescendTree!/expand/backUp with identical injected (prob, prior, v): trees identical after every rollout."""
    ospec = oracle.Spec(*GAME_SPECS[name])
    L, R = 40, 24
    rng = np.random.default_rng(11)
    pos = random_positions(ospec, L, seed=21, max_plies=max(2, ospec.maxLen - 4))
    ctx = ctx_for(name, R, L)
    ctx.re_init(pos)
    ctx.search_begin()
    t = oracle.Tree(ospec, R, L)
    t.reinit(pos)
    t.search_begin()
    for r in range(R):
        prob = rng.uniform(0, 1, size=(L, ospec.maxLen)).astype(f32)
        prob[prob == 0] = 0.5
        prior = rng.uniform(0.01, 1, size=(L, ospec.A)).astype(f32)
        prior /= prior.sum(1, keepdims=True)
        v = rng.uniform(0, 1, size=L).astype(f32)
        last = r == R - 1
        ctx.select(r, 1.5, last=last, prob=prob)
        t.select(0, 1.5, prob=prob)          # the oracle indexes prob[(rollout*L+i)...]: slice per rollout -> rollout 0
        leaf, batch = ctx.leaves()
        oleaf, obatch = t.leaf_batch()
        assert np.array_equal(leaf, oleaf), (name, r)
        assert np.array_equal(batch, obatch)
        ctx.expand_backup(prior, v, training=training, last=last)
        t.expand_backup(prior, v, training)
        if r in (0, 1, R // 2, R - 1):
            compare_trees(ctx.tree(), t.dump(), f"{name} rollout {r}")
    t.finish_search()
    pol, batch = ctx.roots()
    opol, obatch = t.roots()
    assert_bits_equal(pol, opol, "policy_final")
    assert np.array_equal(batch, obatch)
    ctx.close()


@pytest.mark.parametrize("name,n,k", [("connect4", 128, 6), ("ttt", 128, 6), ("gobang9", 128, 2), ("hex7", 128, 2), ("reversi8", 128, 2)])
def test_mcts_single_philox_fp32_bit_exact(name, n, k):
    """agpu_search == mcts_single with the in-kernel Philox stream and the fp32 network: whole trees identical."""
    ospec = oracle.Spec(*GAME_SPECS[name])
    L, R = 48, 32
    pnet, onet = make_nets(GAME_SPECS[name], n, k, seed=4)
    pos = random_positions(ospec, L, seed=33, max_plies=max(2, ospec.maxLen // 2))
    uids = (np.arange(L) * 7 + 3).astype(np.uint32)
    ctx = ctx_for(name, R, L, n, k)
    ctx.set_weights(pnet)
    ctx.re_init(pos, uids)
    ctx.mcts_single(R, training=True, cpuct=1.5, seed=0xDEADBEEF12345, ply=5)
    t = oracle.Tree(ospec, R, L)
    t.reinit(pos, uids)
    t.mcts_single(onet, R, True, 1.5, seed=0xDEADBEEF12345, ply=5, nn_mode=oracle.Net.FP32)
    compare_trees(ctx.tree(), t.dump(), name)
    pol, batch = ctx.roots()
    opol, obatch = t.roots()
    assert_bits_equal(pol, opol, "policy_final")
    assert np.array_equal(batch, obatch)
    ctx.close()


def test_visits_below_capacity_and_R1():
    ospec = oracle.Spec(*GAME_SPECS["connect4"])
    pnet, onet = make_nets(GAME_SPECS["connect4"], 128, 1, seed=2)
    L = 16
    ctx = ctx_for("connect4", 16, L, 128, 1)
    ctx.set_weights(pnet)
    for visits in (1, 2, 7):
        ctx.re_init(ospec.position(L))
        ctx.mcts_single(visits, training=True, cpuct=1.5, seed=1)
        t = oracle.Tree(ospec, 16, L)
        t.reinit(ospec.position(L))
        t.mcts_single(onet, visits, True, 1.5, seed=1)
        assert_bits_equal(ctx.roots()[0], t.roots()[0], f"visits={visits}")
    ctx.close()


@pytest.mark.parametrize("name,games,R", [("ttt", 256, 16), ("connect4", 192, 16), ("hex5", 96, 12), ("reversi6", 64, 8), ("gobang5", 96, 12)])
def test_selfplay_fp32_end_to_end_bit_exact(name, games, R):
    """agpu_selfplay == mcts(actor, visits, ngames, buffer): every sample (state, π̄, player, value, fstate), in push
    order, and the win/draw/loss tally equal the oracle's, bit for bit, with the fp32 network in the loop."""
    ospec = oracle.Spec(*GAME_SPECS[name])
    pnet, onet = make_nets(GAME_SPECS[name], 128, 2, seed=6)
    ctx = ctx_for(name, R, games, 128, 2)
    ctx.set_weights(pnet)
    res, stats, smp = ctx.selfplay(R, games, cpuct=1.5, seed=77, uid_base=1000)
    osmp = oracle.Samples(ospec, games * ospec.maxLen)
    ores, ost = oracle.selfplay(ospec, onet, R, games, cpuct=1.5, seed=77, uid_base=1000, samples=osmp)
    assert np.array_equal(res, ores), (res, ores)
    assert stats["faults"] == 0 and ost["faults"] == 0
    for key in ("sims", "positions", "plies", "total_length"):
        assert stats[key] == ost[key], key
    n = osmp.count
    assert len(smp["player"]) == n
    assert np.array_equal(smp["state"], osmp.state[:n])
    assert np.array_equal(smp["player"], osmp.player[:n])
    assert np.array_equal(smp["game"], osmp.game[:n]) and np.array_equal(smp["ply"], osmp.ply[:n])
    assert_bits_equal(smp["policy"], osmp.policy[:n], "sample policy")
    assert_bits_equal(smp["value"], osmp.value[:n], "sample value")
    assert np.array_equal(smp["fstate"], osmp.fstate[:n])
    ctx.close()


def test_duel_fp32_bit_exact():
    ospec = oracle.Spec(*GAME_SPECS["connect4"])
    p1, o1 = make_nets(GAME_SPECS["connect4"], 128, 2, seed=1)
    p2, o2 = make_nets(GAME_SPECS["connect4"], 128, 2, seed=2)
    ctx = ctx_for("connect4", 8, 128, 128, 2)
    ctx.set_weights(p1, 0)
    ctx.set_weights(p2, 1)
    res, st = ctx.duel(8, 128, cpuct=2.0, seed=9)
    ores, ost = oracle.duel(ospec, o1, o2, 8, 128, cpuct=2.0, seed=9)
    assert np.array_equal(res, ores)
    assert st["positions"] == ost["positions"] and st["plies"] == ost["plies"]
    ctx.close()


def test_shard_invariance_and_determinism():
    """Games are keyed by uid: two shards reproduce the single run's samples game by game; a rerun is identical."""
    pnet, _ = make_nets(GAME_SPECS["connect4"], 128, 2, seed=8)
    ctx = ctx_for("connect4", 12, 256, 128, 2)
    ctx.set_weights(pnet)
    res, st, smp = ctx.selfplay(12, 256, cpuct=1.5, seed=5)
    res2, st2, smp2 = ctx.selfplay(12, 256, cpuct=1.5, seed=5)
    assert np.array_equal(res, res2) and all(np.array_equal(smp[k], smp2[k]) for k in smp)
    ra, _, sa = ctx.selfplay(12, 128, cpuct=1.5, seed=5, uid_base=0)
    rb, _, sb = ctx.selfplay(12, 128, cpuct=1.5, seed=5, uid_base=128)
    assert np.array_equal(ra + rb, res)

    def by_game(s):
        order = np.lexsort((s["ply"], s["game"]))
        return {k: v[order] for k, v in s.items()}
    whole = by_game(smp)
    parts = by_game({k: np.concatenate([sa[k], sb[k]]) for k in sa})
    for k in whole:
        assert np.array_equal(whole[k], parts[k]), k
    ctx.close()


def test_error_paths():
    import alphagpu_b200 as ag
    ctx = ctx_for("connect4", 8, 16)
    with pytest.raises(ag._lib.AlphaGPUError) as e:
        ctx.mcts_single(8, 16)                      # no reinit, no weights
    assert e.value.code == ag._lib.ERR_STATE
    pnet, _ = make_nets(GAME_SPECS["connect4"], 128, 2, seed=8)
    ctx.set_weights(pnet)
    ctx.re_init(ctx.Position(16))
    with pytest.raises(ag._lib.AlphaGPUError) as e:
        ctx.mcts_single(9)                          # visits > rollouts capacity
    assert e.value.code == ag._lib.ERR_INVALID
    with pytest.raises(ag._lib.AlphaGPUError):
        ag.Context(ag.GameSpec.named("gobang", 7, 5), 8, 16, 128, 2)      # size not compiled in -> loud
    with pytest.raises(ag._lib.AlphaGPUError):
        ag.Context(ag.GameSpec.named("connect4"), 300, 16, 128, 2)        # rollouts > 255
    ctx.close()


def test_fdiv_fast_matches_ieee_division():
    """The branch-free division of the α-solve (common.cuh: fdiv_fast) equals __fdiv_rn bit for bit on every operand pair it does not
    flag, over 2^30 random pairs drawn in and around its exponent box (2^34 in profiles/r01_fdiv_check.txt)."""
    import ctypes as C
    from alphagpu_b200 import _lib
    out = (C.c_uint64 * 2)()
    assert _lib.load().agpu_debug_fdiv_check(1 << 30, 12345, out) == 0
    mism, flagged = int(out[0]), int(out[1])
    assert mism == 0
    assert 0.05 * (1 << 30) < flagged < 0.4 * (1 << 30)      # the out-of-box band is exercised, the box is not empty


def _golden(name):
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name)) as f:
        return json.load(f)


def test_selfplay_matches_the_committed_digests():
    """The CUDA path against a COMMITTED fixture (tests/golden/selfplay_digests.json, frozen oracle outputs): whole self-play generations
    with the fp32 evaluator hash to the stored SHA-256 — every state, policy, player, value, fstate, game id and ply of every sample."""
    import os
    import sys
    import alphagpu_b200 as ag
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_selfplay_digests import digest
    for c in _golden("selfplay_digests.json"):
        name, games, R, n, k, net_seed, seed, uid_base, cpuct = c["case"]
        pnet, _ = make_nets(GAME_SPECS[name], n, k, seed=net_seed)
        ctx = ctx_for(name, R, games, n, k)
        ctx.set_weights(pnet)
        res, stats, smp = ctx.selfplay(R, games, cpuct=cpuct, seed=seed, uid_base=uid_base)
        ctx.close()
        assert [int(x) for x in res] == c["results"] and len(smp["player"]) == c["samples"], name
        assert digest(smp, res) == c["digest"], name


def test_plugin_surface_against_golden_traces():
    """tests/golden/kat_games.json replayed through the C ABI (Position / canPlay / play / isOver on the device): bitboards, side to move,
    round counter, legality and outcome after every move equal the committed vectors."""
    import alphagpu_b200 as ag
    ctxs = {}
    for case in _golden("kat_games.json"):
        key = tuple(case["spec"])
        if key not in ctxs:
            ctxs[key] = ag.Context(ag.GameSpec(*key), 4, 8, 128, 2, 0, 1)
        ctx = ctxs[key]
        pos = ctx.Position(1)
        for step in case["trace"]:
            if step["move"] is not None:
                assert ctx.canPlay(pos)[0, step["move"] - 1], case["name"]
                pos = ctx.play(pos, step["move"])
            p = pos[0]
            assert [hex(int(x)) for x in p["bplayer"]["chunks"]] == step["bplayer"], case["name"]
            assert [hex(int(x)) for x in p["bopponent"]["chunks"]] == step["bopponent"], case["name"]
            if step["legalplay"] is not None:
                assert [hex(int(x)) for x in p["legalplay"]["chunks"]] == step["legalplay"], case["name"]
            assert int(p["player"]) == step["player"]
            if pos.dtype.itemsize == 104:
                assert int(p["aux"]) == step["aux"]
            over, res = ctx.isOver(pos)
            assert bool(over[0]) == step["over"]
            if step["over"]:
                assert int(res[0]) == step["result"]
            assert [int(a) + 1 for a in np.nonzero(ctx.canPlay(pos)[0])[0]] == step["legal"], case["name"]
    for ctx in ctxs.values():
        ctx.close()
