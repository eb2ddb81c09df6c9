"""agpu_multi_*: several devices behind one call from one process (include/alphagpu.h; SURVEY §8(b)/(e), the caller is selfplay.jl:34).
The games are block-partitioned by uid and the RNG is keyed by uid, so the gathered samples must equal the single-context run's, game
by game, and arrive in ascending uid blocks.  On a one-GPU box the two contexts share device 0 (the logic under test — threads,
sharding, gather offsets — is the same); with two or more GPUs they sit on different devices."""
import numpy as np
import pytest

import oracle
from conftest import GAME_SPECS
from helpers import assert_bits_equal, make_exact_nets, make_nets

pytestmark = pytest.mark.gpu


def _devices(n):
    import torch
    have = torch.cuda.device_count()
    return [i % max(1, have) for i in range(n)]


def _by_game(s):
    order = np.lexsort((s["ply"], s["game"]))
    return {k: v[order] for k, v in s.items()}


@pytest.mark.parametrize("ngpus,games", [(2, 300), (3, 100), (2, 1)])
def test_multi_selfplay_reproduces_single_context_by_uid(ngpus, games):
    import alphagpu_b200 as ag
    spec = ag.GameSpec.named("connect4")
    pnet, _ = make_nets(GAME_SPECS["connect4"], 128, 2, seed=8)
    R = 12
    one = ag.Context(spec, R, games, 128, 2)
    one.set_weights(pnet)
    res1, st1, smp1 = one.selfplay(R, games, cpuct=1.5, seed=5, uid_base=40)
    one.close()
    m = ag.MultiContext(spec, R, games, 128, 2, ngpus, devices=_devices(ngpus))
    m.set_weights(pnet)
    res, st, smp = m.selfplay(R, games, cpuct=1.5, seed=5, uid_base=40)
    res_again, _, smp_again = m.selfplay(R, games, cpuct=1.5, seed=5, uid_base=40)      # the contexts are reused
    m.close()
    assert np.array_equal(res, res1) and np.array_equal(res_again, res1)
    for key in ("sims", "positions", "total_length", "faults"):
        assert st[key] == st1[key], key
    assert len(smp["player"]) == len(smp1["player"])
    a, b, c = _by_game(smp), _by_game(smp1), _by_game(smp_again)
    for k in a:
        assert np.array_equal(a[k], b[k]) and np.array_equal(c[k], b[k]), k
    # gathered order: ascending uid blocks (device order), push order within a block
    base, blocks = 0, []
    for r in range(ngpus):
        cnt = games // ngpus + (1 if r < games % ngpus else 0)
        blocks.append((40 + base, 40 + base + cnt))
        base += cnt
    g = smp["game"]
    edges = np.nonzero(np.diff(np.searchsorted([hi for _, hi in blocks], g, side="right")))[0]
    assert len(edges) <= ngpus - 1 and np.all(np.diff(np.searchsorted([hi for _, hi in blocks], g, side="right")) >= 0)


def test_multi_selfplay_tc_exact_net_vs_oracle():
    """The multi-device call through the per-ply tcgen05 kernel, held to the fp32 oracle bit for bit (exact-arithmetic net)."""
    import alphagpu_b200 as ag
    name, R, games = "connect4", 32, 500
    ospec = oracle.Spec(*GAME_SPECS[name])
    pnet, onet = make_exact_nets(GAME_SPECS[name], 128, 6, seed=2)
    m = ag.MultiContext(ag.GameSpec.named(name), R, games, 128, 6, 2, devices=_devices(2))
    m.set_weights(pnet)
    res, st, smp = m.selfplay(R, games, cpuct=1.5, seed=77)
    m.close()
    osmp = oracle.Samples(ospec, games * ospec.maxLen)
    ores, ost = oracle.selfplay(ospec, onet, R, games, cpuct=1.5, seed=77, samples=osmp, nn_mode=oracle.Net.FP32)
    assert np.array_equal(res, ores)
    n = osmp.count
    o = _by_game(dict(state=osmp.state[:n], policy=osmp.policy[:n], player=osmp.player[:n], value=osmp.value[:n], fstate=osmp.fstate[:n],
                      game=osmp.game[:n], ply=osmp.ply[:n]))
    a = _by_game(smp)
    for k in ("state", "player", "fstate", "game", "ply"):
        assert np.array_equal(a[k], o[k]), k
    assert_bits_equal(a["policy"], o["policy"], "policy")
    assert_bits_equal(a["value"], o["value"], "value")


def test_multi_duel_sums_the_tallies():
    import alphagpu_b200 as ag
    spec = ag.GameSpec.named("connect4")
    p1, _ = make_nets(GAME_SPECS["connect4"], 128, 2, seed=1)
    p2, _ = make_nets(GAME_SPECS["connect4"], 128, 2, seed=2)
    one = ag.Context(spec, 8, 128, 128, 2)
    one.set_weights(p1, 0); one.set_weights(p2, 1)
    r1, _ = one.duel(8, 128, cpuct=2.0, seed=9)
    one.close()
    m = ag.MultiContext(spec, 8, 128, 128, 2, 2, devices=_devices(2))
    m.set_weights(p1, 0); m.set_weights(p2, 1)
    r2, _ = m.duel(8, 128, cpuct=2.0, seed=9)
    m.close()
    assert np.array_equal(r1, r2)


def test_mcts_entry_point_with_ngpus():
    """mcts(actor, visits, ngames, buffer; ...) with ngpus=2 pushes the same number of samples as the single-context call."""
    import alphagpu_b200 as ag
    spec = ag.GameSpec.named("connect4")
    pnet, _ = make_nets(GAME_SPECS["connect4"], 128, 2, seed=3)
    b1, b2 = ag.PoolSample(spec, 10000), ag.PoolSample(spec, 10000)
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("mcts(ngpus=2) places its contexts on devices 0 and 1")
    o1 = ag.mcts(pnet, 8, 200, b1, spec=spec, cpuct=1.5, seed=4)
    o2 = ag.mcts(pnet, 8, 200, b2, spec=spec, cpuct=1.5, seed=4, ngpus=2)
    assert o1["valid"] and o2["valid"] and b1.length_buffer() == b2.length_buffer()
    assert np.array_equal(o1["stats"]["results"], o2["stats"]["results"])


def test_pinned_sample_buffers_stream_the_same_samples():
    """Context.pinned_samples() (agpu_host_alloc): with page-locked destinations agpu_selfplay streams every ply's rows out while the
    next ply searches; the arrays must equal the ones returned through ordinary (pageable) numpy arrays."""
    import alphagpu_b200 as ag
    spec = ag.GameSpec.named("connect4")
    net = ag.ressimplesf(2 * spec.VectorizedState, spec.maxActions, 128, 6, seed=3)
    ctx = ag.Context(spec, 32, 700, 128, 6)
    ctx.set_weights(net)
    res_a, st_a, smp_a = ctx.selfplay(32, 700, cpuct=1.5, seed=12, uid_base=9)
    res_b, st_b, smp_b = ctx.selfplay(32, 700, cpuct=1.5, seed=12, uid_base=9, out=ctx.pinned_samples())
    assert np.array_equal(res_a, res_b) and st_a["positions"] == st_b["positions"]
    for k in ("state", "policy", "player", "value", "fstate", "game", "ply"):
        assert np.array_equal(smp_a[k], np.array(smp_b[k])), k
    ctx.close()
