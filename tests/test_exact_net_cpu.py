"""CPU check of the premise behind tests/test_gpu_exact.py: with an exact-arithmetic network (helpers.make_exact_nets) the oracle's
fp32 evaluator and its operand-faithful 16-bit modes (fp16 / bf16 operands, 16-bit residual stream) agree in every bit — on random
positions and on every position a whole self-play generation visits — so a tensor-core kernel can be held to the fp32 oracle exactly."""
import numpy as np
import pytest

import oracle
from conftest import GAME_SPECS
from helpers import make_exact_nets, random_positions


@pytest.mark.parametrize("name,n,k", [("connect4", 128, 6), ("ttt", 128, 6), ("hex7", 512, 8), ("gobang9", 512, 8), ("reversi8", 512, 8)])
def test_exact_net_is_format_independent(name, n, k):
    ospec = oracle.Spec(*GAME_SPECS[name])
    pnet, onet = make_exact_nets(GAME_SPECS[name], n, k, seed=1)
    x = ospec.encode(random_positions(ospec, 300, seed=3))
    ref_l, ref_v = onet.forward(x, mode=oracle.Net.FP32)
    modes = [oracle.Net.F16, oracle.Net.F16_RESID] + ([oracle.Net.BF16, oracle.Net.BF16_RESID] if n == 128 else [])
    for mode in modes:
        l, v = onet.forward(x, mode=mode)
        assert np.array_equal(l.view(np.uint32), ref_l.view(np.uint32)) and np.array_equal(v.view(np.uint32), ref_v.view(np.uint32)), mode
    p = oracle.softmax(ref_l)
    assert (p.max(1) - p.min(1)).mean() > 0.05          # the policy is not degenerate: the search sees varied priors
    assert len(np.unique(ref_v)) > 20


def test_exact_net_selfplay_is_format_independent():
    ospec = oracle.Spec(*GAME_SPECS["connect4"])
    pnet, onet = make_exact_nets(GAME_SPECS["connect4"], 128, 6, seed=11)
    out = []
    for mode in (oracle.Net.FP32, oracle.Net.F16, oracle.Net.BF16):
        smp = oracle.Samples(ospec, 64 * ospec.maxLen)
        res, st = oracle.selfplay(ospec, onet, 32, 64, cpuct=1.5, seed=5, samples=smp, nn_mode=mode)
        out.append((res.copy(), smp.count, smp.policy[:smp.count].copy(), smp.state[:smp.count].copy()))
    for o in out[1:]:
        assert np.array_equal(o[0], out[0][0]) and o[1] == out[0][1]
        assert np.array_equal(o[2].view(np.uint32), out[0][2].view(np.uint32)) and np.array_equal(o[3], out[0][3])
