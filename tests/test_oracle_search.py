"""Oracle search semantics (mcts_gpu.jl:100-339, 376-462): the α-solve known answer of SURVEY
Appendix B.6, a step-by-step numpy-float32 re-derivation of one descent, and structural invariants
of mcts_single with injected and network evaluators."""
import struct

import numpy as np
import pytest

import oracle
from conftest import GAME_SPECS

f32 = np.float32


def alpha_solve_numpy(prior, q, visits, order, cpuct):
    """kdescendTree! lines 116-169 in numpy float32 scalars (each op rounds to binary32)."""
    A = len(prior)
    has_child = [False] * A
    for a in order:
        has_child[a - 1] = True
    n, prior_rem, Acount = f32(1), f32(0), f32(0)
    for k in range(A):
        n = f32(n + visits[k])
        if not has_child[k]:
            prior_rem = f32(prior_rem + prior[k])
        if prior[k] > 0:
            Acount = f32(Acount + f32(1))
    lam = f32(f32(f32(cpuct) * np.sqrt(n, dtype=f32)) / f32(Acount + n))
    alpha = f32(0)
    prior_rem = f32(prior_rem * lam)
    for k in range(A):
        gap = max(f32(lam * prior[k]), f32(1e-4))
        alpha = max(alpha, f32(q[k] + gap))
    err = f32(np.inf)
    iters = 0
    for _ in range(100):
        iters += 1
        S = f32(prior_rem / alpha)
        g = f32(f32(-prior_rem) / f32(alpha * alpha))
        for a in order:
            top = f32(lam * prior[a - 1])
            bot = f32(alpha - q[a - 1])
            S = f32(S + f32(top / bot))
            g = f32(g + f32(f32(-top) / f32(bot * bot)))
        newerr = f32(S - f32(1))
        if newerr < f32(0.001) or newerr == err:
            break
        alpha = f32(alpha - f32(newerr / g))
        err = newerr
    pol = [f32(f32(lam * prior[k]) / f32(alpha - q[k])) for k in range(A)]
    return lam, alpha, iters, pol


def test_alpha_solve_known_answer():
    """SURVEY.md Appendix B.6."""
    spec = oracle.Spec(oracle.CONNECT4)
    prior = np.array([.10, .05, .20, .30, .15, 0, .20], f32)
    q = np.array([.40, 0, .55, .62, 0, 0, .35], f32)
    visits = np.array([3, 0, 5, 9, 0, 0, 2], f32)
    order = [4, 3, 1, 7]
    t = oracle.Tree(spec, 8, 1)
    t.reinit(spec.position(1))
    t.search_begin()
    t.poke(0, 1, prior, q, visits, order)
    prob = np.full((1, 1, spec.maxLen), 0.5, f32)
    t.select(0, 1.5, prob=prob)
    pol = t.dump()["policy"][0, 0]
    lam, alpha, iters, want = alpha_solve_numpy(prior, q, visits, order, 1.5)
    assert struct.pack("<f", lam).hex() == "9b19843e"
    assert struct.pack("<f", alpha).hex() == "a7a3453f"
    assert iters == 5
    assert [struct.pack("<f", x).hex() for x in pol] == ["3e088e3d", "dde2883c", "d6fc6d3e", "6956023f", "4c544d3d", "00000000", "e468fa3d"]
    assert [struct.pack("<f", x).hex() for x in want] == [struct.pack("<f", x).hex() for x in pol]
    assert abs(float(np.sum(pol, dtype=np.float64)) - 1.0000018) < 2e-7
    c = t.counters()
    assert c["newton_solves"] == 1 and c["newton_iters"] == 5
    # cumulative sampling: u=0.5 falls in action 4 (cum 0.0694,0.0861,0.3185,0.8276)
    leaf, _ = t.leaf_batch()
    d = t.dump()
    assert d["action"][0, leaf[0] - 1] == 4


def fmcts_policy_numpy(prior, w, n, c):
    """The reference's OTHER statement of the same solve: FMCTS.newton + bestChild/extractRoot (fast_mcts.jl:41-69, 218-223, 293-302),
    numpy float32, every action summed individually (no lumped prior_rem), q = w/n formed on the fly."""
    A = len(prior)
    visits = f32(1) + f32(np.sum(n))                                   # node.visits: own first visit + one per action visit
    nact = int(np.sum(prior > 0))                                      # getActionNumber(state) = number of legal actions
    lam = f32(f32(f32(c) * np.sqrt(visits, dtype=f32)) / f32(f32(nact) + visits))
    alpha = f32(0)
    for k in range(A):
        gap = max(f32(lam * prior[k]), f32(1e-4))
        alpha = max(alpha, gap if n[k] == 0 else f32(f32(w[k] / n[k]) + gap))
    err = f32(np.inf)
    for _ in range(100):
        S, g = f32(0), f32(0)
        for k in range(A):
            top = f32(lam * prior[k])
            bot = alpha if n[k] == 0 else f32(alpha - f32(w[k] / n[k]))
            S = f32(S + f32(top / bot))
            g = f32(g + f32(f32(-top) / f32(bot * bot)))
        newerr = f32(S - f32(1))
        if newerr < f32(0.001) or newerr == err:
            break
        alpha = f32(alpha - f32(newerr / g))
        err = newerr
    return alpha, [f32(top_k / (alpha if n[k] == 0 else f32(alpha - f32(w[k] / n[k]))))
                   for k, top_k in enumerate(f32(lam) * prior.astype(f32))]


def test_alpha_solve_agrees_with_the_references_cpu_statement():
    """mcts_gpu.jl's kdescendTree! solve (oracle) and fast_mcts.jl's newton are two texts of one fixed point: same λ, same start, same
    stopping rule, different summation grouping — α and π̄ agree to rounding on random node statistics (A = 7, 9, 49, 81; boards where
    every action can be played from Position(), so that any subset can be given children)."""
    rng = np.random.default_rng(17)
    worst = 0.0
    for A, name in ((7, "connect4"), (9, "ttt"), (49, "hex7"), (81, "gobang9")):
        spec = oracle.Spec(*GAME_SPECS[name])
        for trial in range(40):
            prior = rng.uniform(0.01, 1, A).astype(f32)
            prior[rng.uniform(size=A) < 0.2] = 0                        # illegal actions
            if not np.any(prior > 0):
                prior[0] = 1
            prior = (prior / prior.sum(dtype=f32)).astype(f32)
            legal = np.nonzero(prior > 0)[0]
            order = [int(a) + 1 for a in rng.permutation(legal)[:rng.integers(0, len(legal) + 1)]]
            visits = np.zeros(A, f32)
            q = np.zeros(A, f32)
            for a in order:
                visits[a - 1] = f32(rng.integers(1, 30))
                q[a - 1] = f32(rng.uniform(0, 1))
            t = oracle.Tree(spec, A + 2, 1)
            t.reinit(spec.position(1))
            t.search_begin()
            t.poke(0, 1, prior, q, visits, order)
            t.select(0, 1.5, prob=np.full((1, 1, spec.maxLen), 0.5, f32))
            pol = t.dump()["policy"][0, 0]
            lam, alpha, iters, want = alpha_solve_numpy(prior, q, visits, order, 1.5)
            assert np.array_equal(np.asarray(want, f32).view(np.uint32), pol.view(np.uint32))      # the oracle is the mcts_gpu text, bit for bit
            alpha2, pol2 = fmcts_policy_numpy(prior, q * visits, visits, 1.5)
            assert abs(float(alpha2) - float(alpha)) < 2e-5 * max(1.0, float(alpha))
            worst = max(worst, float(np.max(np.abs(np.asarray(pol2, f32) - pol))))
    assert worst < 1e-4, worst


def rand_net(spec, n, k, seed):
    rng = np.random.default_rng(seed)
    glorot = lambda o, i: rng.uniform(-1, 1, size=(o, i)).astype(f32) * f32(np.sqrt(6.0 / (o + i)))
    return oracle.Net(glorot(n, 2 * spec.VS), [glorot(n, n) for _ in range(k)], glorot(spec.A, n), np.zeros(spec.A, f32),
                      glorot(1, n), np.zeros(1, f32))


@pytest.mark.parametrize("name", ["connect4", "ttt", "hex5", "reversi6"])
def test_mcts_single_invariants(name):
    spec = oracle.Spec(*GAME_SPECS[name])
    L, R = 24, 32
    net = rand_net(spec, 32, 2, 3)
    t = oracle.Tree(spec, R, L)
    t.reinit(spec.position(L), np.arange(L, dtype=np.uint32))
    t.mcts_single(net, R, True, 1.5, seed=11, ply=0)
    d = t.dump()
    pol, batch = t.roots()
    legal = spec.legal(spec.position(L))
    assert np.all(d["nnodes"] == R)                       # rollout 1 expands the root, every later one adds a node (no terminals at ply 0)
    assert np.all(d["parent"][:, 0] == 0)
    for g in range(L):
        for nd in range(1, R):
            p, a = d["parent"][g, nd], d["action"][g, nd]
            assert 1 <= p <= nd and d["child"][g, p - 1, a - 1] == nd + 1
        # root visit counts: R-1 backups passed through the root
        assert d["visits"][g, 0].sum() == R - 1
        # children in slot order are the order of creation (ascending node id)
        nc = d["nchild"][g, 0]
        ids = [d["child"][g, 0, a - 1] for a in d["order"][g, 0, :nc]]
        assert ids == sorted(ids) and len(set(ids)) == nc
    assert np.all((pol > 0) == legal)
    assert np.allclose(pol.sum(1), 1.0, atol=2e-3)
    assert np.array_equal(batch, spec.encode(spec.position(L)))
    # q is a mean of values in [0,1]
    assert np.all(d["q"] >= 0) and np.all(d["q"] <= 1)
    # determinism and uid-keyed RNG: a permuted slot order gives the permuted result
    perm = np.random.default_rng(0).permutation(L)
    t2 = oracle.Tree(spec, R, L)
    t2.reinit(spec.position(L), perm.astype(np.uint32))
    t2.mcts_single(net, R, True, 1.5, seed=11, ply=0)
    pol2, _ = t2.roots()
    assert np.array_equal(pol2, pol[perm])


def test_policy_final_is_pre_last_backup():
    """A.7: policy_final = root π̄ at the start of the last rollout; with R=1 it is the mixed prior."""
    spec = oracle.Spec(oracle.CONNECT4)
    L = 4
    rng = np.random.default_rng(5)
    pri = rng.uniform(0.05, 1, size=(1, L, spec.A)).astype(f32)
    pri /= pri.sum(-1, keepdims=True)
    v = rng.uniform(0, 1, size=(1, L)).astype(f32)
    t = oracle.Tree(spec, 4, L)
    t.reinit(spec.position(L))
    t.mcts_single(None, 1, True, 1.5, inj_prior=pri, inj_v=v)
    pol, _ = t.roots()
    want = (f32(0.75) * pri[0] / pri[0].sum(-1, keepdims=True, dtype=f32)).astype(f32) + f32(0.25) / f32(7)
    assert np.allclose(pol, want, atol=1e-6)
    t.reinit(spec.position(L))      # re_init resets `expanded` (mcts_gpu.jl:371); mcts_single itself does not
    t.mcts_single(None, 1, False, 1.5, inj_prior=pri, inj_v=v)
    pol, _ = t.roots()
    assert np.allclose(pol, pri[0], atol=1e-6)


def test_terminal_leaf_backup_values():
    """Drive a Connect4 game to a position with an immediate win and check terminal backups (A.6)."""
    spec = oracle.Spec(oracle.CONNECT4)
    pos = spec.position(1)
    for m in [4, 1, 4, 1, 4, 1]:      # player +1 to move with three in column 4
        pos = spec.play(pos, m)
    R = 16
    t = oracle.Tree(spec, R, 1)
    t.reinit(pos)
    pri = np.full((R, 1, 7), 1 / 7, f32)
    v = np.full((R, 1), 0.5, f32)
    t.mcts_single(None, R, False, 1.5, inj_prior=pri, inj_v=v, seed=1)
    d = t.dump()
    # the child reached by action 4 is terminal: never expanded, and q(root, 4) == 1 (mover wins)
    c = d["child"][0, 0, 3]
    assert c > 0 and d["expanded"][0, c - 1] == 0
    assert d["q"][0, 0, 3] == 1.0
    assert d["visits"][0, 0, 3] >= 1


def test_selfplay_and_duel_complete():
    spec = oracle.Spec(*GAME_SPECS["ttt"])
    net = rand_net(spec, 32, 2, 1)
    smp = oracle.Samples(spec, 64 * 9)
    res, st = oracle.selfplay(spec, net, 16, 64, cpuct=1.5, seed=3, samples=smp)
    assert res.sum() == 64 and st["faults"] == 0
    assert smp.count == st["positions"] and st["sims"] == st["positions"] * 16
    n = smp.count
    assert set(np.unique(smp.value[:n])) <= {0.0, 0.5, 1.0}
    assert np.all(np.abs(smp.policy[:n].sum(1) - 1) < 2e-3)
    # value is from the sample's own perspective: winner's samples 1, loser's 0
    first = smp.ply[:n] == 0
    assert np.all(smp.player[:n][first] == 1)
    # fstate = final board seen from the sample's player (±1 everywhere)
    assert set(np.unique(smp.fstate[:n])) <= {-1, 1}
    res2, st2 = oracle.selfplay(spec, net, 16, 64, cpuct=1.5, seed=3)
    assert np.array_equal(res, res2) and st == st2
    # shard invariance: two halves with uid_base give the same totals
    ra, _ = oracle.selfplay(spec, net, 16, 32, cpuct=1.5, seed=3, uid_base=0)
    rb, _ = oracle.selfplay(spec, net, 16, 32, cpuct=1.5, seed=3, uid_base=32)
    assert np.array_equal(ra + rb, res)
    dres, dst = oracle.duel(spec, net, rand_net(spec, 32, 2, 2), 8, 32, seed=5)
    assert dres.sum() == 32


def test_selfplay_digests_match_the_committed_fixture():
    """tests/golden/selfplay_digests.json (made by tests/golden/make_selfplay_digests.py): whole oracle self-play generations, frozen as
    SHA-256 over every sample array.  Any change to the oracle that moves a single bit of a single sample shows up here; the GPU twin of
    this test (tests/test_gpu_parity.py) holds the CUDA path to the same fixture."""
    import json
    import os
    import sys
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, golden)
    import make_selfplay_digests as m
    with open(os.path.join(golden, "selfplay_digests.json")) as f:
        cases = json.load(f)
    assert len(cases) >= 5
    for c in cases:
        got = m.oracle_case(tuple(c["case"]))
        assert got["results"] == c["results"] and got["samples"] == c["samples"], c["case"]
        assert got["digest"] == c["digest"], c["case"]
