"""The five BASELINE.json configurations as parity cases (-m gpu): each game with its configured network shape and rollout
count, played end to end on the GPU (fp32 evaluator = bit-exact parity mode) and compared sample by sample with the oracle.
Config 1 runs at its full size (1024 games); configs 3-5 at the reference shape but a reduced number of games so that the CPU
oracle finishes in seconds (sims/s on CPU is size-independent; the full game counts only change how many tiles are in flight)."""
import time

import numpy as np
import pytest

import oracle
from conftest import GAME_SPECS
from helpers import assert_bits_equal, make_nets

pytestmark = pytest.mark.gpu

CONFIGS = [
    # name, width, blocks, rollouts, games
    ("ttt", 128, 6, 64, 1024),         # config 1: Gobang N=3 Nvict=3, 128x6, rollout 64, samples 1024 (full size)
    ("connect4", 128, 6, 64, 512),     # config 2 (shape), the metric configuration
    ("hex7", 512, 8, 64, 24),          # config 3
    ("gobang9", 512, 8, 128, 12),      # config 4
    ("reversi8", 512, 8, 64, 16),      # config 5
]


@pytest.mark.parametrize("name,n,k,R,games", CONFIGS)
def test_baseline_config_selfplay_bit_exact(name, n, k, R, games):
    import alphagpu_b200 as ag
    g, N, nv = GAME_SPECS[name]
    ospec = oracle.Spec(g, N, nv)
    pnet, onet = make_nets(GAME_SPECS[name], n, k, seed=13)
    ctx = ag.Context(ag.GameSpec(g, N, nv), R, games, n, k, 0, ag._lib.NN_FP32)
    ctx.set_weights(pnet)
    t0 = time.time()
    res, st, smp = ctx.selfplay(R, games, cpuct=1.5, seed=5, uid_base=7)
    t_gpu = time.time() - t0
    osmp = oracle.Samples(ospec, games * ospec.maxLen)
    t0 = time.time()
    ores, ost = oracle.selfplay(ospec, onet, R, games, cpuct=1.5, seed=5, uid_base=7, samples=osmp)
    t_cpu = time.time() - t0
    print(f"\n{name} {n}x{k} R={R} games={games}: sims {st['sims']} plies {st['plies']} results {res.tolist()} gpu {t_gpu:.2f}s cpu {t_cpu:.2f}s")
    assert np.array_equal(res, ores) and st["faults"] == 0
    for key in ("sims", "positions", "plies", "total_length"):
        assert st[key] == ost[key], key
    m = osmp.count
    assert np.array_equal(smp["state"], osmp.state[:m]) and np.array_equal(smp["player"], osmp.player[:m])
    assert_bits_equal(smp["policy"], osmp.policy[:m], "policy")
    assert_bits_equal(smp["value"], osmp.value[:m], "value")
    assert np.array_equal(smp["fstate"], osmp.fstate[:m])
    ctx.close()


def test_duel_tc_runs_and_is_deterministic():
    """duelnetwork through the tensor-core path: both nets resident, 32 rollouts x 1024 games as selfplay.jl:56 calls it."""
    import alphagpu_b200 as ag
    name = "connect4"
    p1, _ = make_nets(GAME_SPECS[name], 128, 6, seed=1)
    p2, _ = make_nets(GAME_SPECS[name], 128, 6, seed=2)
    spec = ag.GameSpec(*GAME_SPECS[name])
    a = ag.duelnetwork(p1, p2, 32, 1024, spec=spec, seed=3)
    b = ag.duelnetwork(p1, p2, 32, 1024, spec=spec, seed=3)
    assert a == b and sum(a) == 1024
