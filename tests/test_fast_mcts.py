"""Interactive-play mirror (alphagpu_b200/fast_mcts.py: FMCTS.MctsContext + the testvsordi drivers).
CPU part: move notation of the four drivers, root value, loud failure without a GPU.
GPU part: `MctsContext(pos, readout)` is `mcts_single` with one game and training=false — policy and tree statistics equal the
oracle's bit for bit with the fp32 evaluator; a whole engine-vs-script game replays to the same outcome through the oracle."""
import numpy as np
import pytest

import oracle
from conftest import GAME_SPECS
from helpers import assert_bits_equal, make_nets, random_positions


def spec_of(name):
    import alphagpu_b200 as ag
    g, N, nv = GAME_SPECS[name]
    return ag.GameSpec(g, N, nv)


def test_move_dictionaries_follow_the_drivers():
    from alphagpu_b200.fast_mcts import move_dictionaries
    fwd, inv = move_dictionaries(spec_of("reversi8"))          # testrev8.jl:1-13
    assert fwd["a1"] == 1 and fwd["a8"] == 8 and fwd["b1"] == 9 and fwd["h8"] == 64 and fwd["p"] == 65
    assert inv[65] == "pass" and inv[10] == "b2" and len(inv) == 65
    fwd, inv = move_dictionaries(spec_of("reversi6"))          # testrev6.jl: the same with a 6x6 board
    assert fwd["f6"] == 36 and fwd["p"] == 37 and inv[7] == "b1"
    fwd, inv = move_dictionaries(spec_of("hex7"))              # testHex.jl:5-17: column letter + row, c = N*(col-1)+row
    assert fwd["A1"] == 1 and fwd["A7"] == 7 and fwd["B1"] == 8 and fwd["G7"] == 49 and inv[9] == "B2"
    fwd, inv = move_dictionaries(spec_of("gobang9"))           # testgobang.jl:39-47: play = 10x+y -> c = N*x+y+1
    assert fwd["0"] == 1 and fwd["8"] == 9 and fwd["10"] == 10 and fwd["88"] == 81 and inv[11] == "11"
    for name in ("reversi8", "reversi6", "hex7", "gobang9", "ttt", "connect4"):
        s = spec_of(name)
        fwd, inv = move_dictionaries(s)
        assert sorted(inv) == list(range(1, s.maxActions + 1))
        assert all(fwd[t if t != "pass" else "p"] == a for a, t in inv.items())


def test_root_value_is_sum_w_over_n():
    from alphagpu_b200.fast_mcts import root_value
    q = np.array([0.25, 0.5, 0.0, 1.0], np.float32)
    n = np.array([4, 2, 0, 1], np.float32)
    # fast_mcts.jl:300: w = q*n per action, N = readout = sum(n) + 1 (the root's own first visit)
    assert root_value(q, n, 8) == pytest.approx((1.0 + 1.0 + 0.0 + 1.0) / 8)


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import alphagpu_b200 as ag
    from alphagpu_b200.fast_mcts import MctsContext, testvsordi
    s = spec_of("connect4")
    net = ag.ressimplesf(2 * s.VectorizedState, s.maxActions, 128, 2, seed=0)
    with pytest.raises(ag._lib.AlphaGPUError):
        MctsContext(1.5, net, s)
    with pytest.raises(ag._lib.AlphaGPUError):
        testvsordi(net, 16, spec=s, moves=[1, 2, 3], log=None)
    with pytest.raises(ValueError):
        MctsContext(1.5, net, s, readout_max=256)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["connect4", "ttt", "hex5", "reversi6", "reversi8", "gobang9"])
def test_mcts_context_is_mcts_single_with_one_game(name):
    import alphagpu_b200 as ag
    from alphagpu_b200.fast_mcts import MctsContext, root_value
    ospec = oracle.Spec(*GAME_SPECS[name])
    pnet, onet = make_nets(GAME_SPECS[name], 128, 2, seed=12)
    puct = MctsContext(1.5, pnet, spec_of(name), readout_max=48, nn_mode=ag._lib.NN_FP32, seed=99)
    positions = random_positions(ospec, 3, seed=4, max_plies=max(2, ospec.maxLen // 2))
    for call, readout in enumerate((48, 17, 1)):
        pos = positions[call:call + 1]
        p, v = puct(pos, readout)
        t = oracle.Tree(ospec, 48, 1)
        t.reinit(pos, np.asarray([call], np.uint32))
        t.mcts_single(onet, readout, False, 1.5, seed=99, ply=0, nn_mode=oracle.Net.FP32)
        assert_bits_equal(p, t.roots()[0][0], f"{name} policy, readout {readout}")
        d = t.dump()
        assert v == root_value(d["q"][0, 0], d["visits"][0, 0], readout)
        assert int(d["visits"][0, 0].sum()) == readout - 1           # the first read-out expands the root (fast_mcts.jl:73-91)
        legal = ospec.legal(pos)[0]
        assert np.all(p[~legal] == 0) and abs(float(p.sum()) - 1) < 2e-3                 # π̄ is a distribution over the legal moves
    with pytest.raises(ValueError):
        puct(positions[:1], 49)                                        # above the context's capacity
    with pytest.raises(ValueError):
        puct(positions[:2], 8)                                         # one position per call
    puct.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name,player", [("connect4", 1), ("ttt", -1), ("reversi6", -1), ("hex5", 1)])
def test_testvsordi_plays_a_legal_game_to_the_end(name, player):
    """Engine (tensor-core evaluator, the product mode) against a scripted opponent that plays the first legal move; the history replays
    through the oracle's plugin functions to a terminal position with the reported winner."""
    from alphagpu_b200.fast_mcts import move_dictionaries, testvsordi
    ospec = oracle.Spec(*GAME_SPECS[name])
    pnet, _ = make_nets(GAME_SPECS[name], 128, 2, seed=5)
    _, inv = move_dictionaries(spec_of(name))

    def first_legal(game):
        a = int(np.nonzero(ospec.legal(game)[0])[0][0]) + 1
        return inv[a] if inv[a] != "pass" else "p"                     # through the text notation, as readline() would deliver it

    lines = []
    history, w = testvsordi(pnet, 24, player, spec=spec_of(name), moves=first_legal, log=lines.append)
    pos = ospec.position(1)
    for k, c in enumerate(history):
        assert ospec.legal(pos)[0, c - 1], (k, c)
        assert not ospec.is_over(pos)[0][0]
        pos = ospec.play(pos, c)
    over, res = ospec.is_over(pos)
    assert over[0] and int(res[0]) == w
    assert sum(s.startswith("coup: ") for s in lines) >= 1 and lines[-1] in ("winner: puct", "match nul", "winner: internet")
