"""Pins the CPU oracle's game plugins (oracle/oracle.cpp) three ways, since the reference ships no
vectors: (1) hand-checked constants of SURVEY Appendix B, (2) the committed
tests/golden/kat_games.json traces (made by tests/golden/make_kats.py from oracle/pyref.py),
(3) per-ply agreement over random games with oracle/pyref.py (big-int restatement of the Julia
text) and tests/naive_rules.py (plain grid rules: Connect4 / k-in-a-row / Hex BFS / Othello)."""
import json
import os
import random

import numpy as np
import pytest

import oracle
from oracle import pyref
from conftest import GAME_SPECS
import naive_rules

HERE = os.path.dirname(os.path.abspath(__file__))


def chunks(p, field):
    return tuple(int(x) for x in p[field]["chunks"])


def replay(spec, moves):
    pos = spec.position(1)
    for m in moves:
        assert spec.can_play(pos, m)[0]
        pos = spec.play(pos, m)
    return pos


# ---------------- (1) Appendix B constants ----------------
def test_appendix_b_connect4():
    s = oracle.Spec(oracle.CONNECT4)
    p = replay(s, [4, 4, 5, 5, 6, 6, 7])
    assert chunks(p[0], "bplayer") == (0x410400000, 0, 0)
    assert chunks(p[0], "bopponent") == (0x20820800000, 0, 0)
    assert p[0]["player"] == -1
    over, res = s.is_over(p)
    assert over[0] and res[0] == 1
    p = replay(s, [1, 2, 1, 2, 1, 2, 1])
    assert chunks(p[0], "bplayer") == (0xE00, 0, 0) and chunks(p[0], "bopponent") == (0x3C, 0, 0)
    over, res = s.is_over(p)
    assert over[0] and res[0] == 1
    p = replay(s, [4, 4, 4, 4, 4, 4, 3])
    assert chunks(p[0], "bplayer") == (0x540000, 0, 0) and chunks(p[0], "bopponent") == (0xAA0000, 0, 0)
    over, res = s.is_over(p)
    assert not over[0]
    assert [a + 1 for a in np.nonzero(s.legal(p)[0])[0]] == [1, 2, 3, 5, 6, 7]


def test_appendix_b_gobang():
    s = oracle.Spec(oracle.GOBANG, 3, 3)
    p = replay(s, [1, 2, 5, 3, 9])
    assert chunks(p[0], "bopponent") == (0x111, 0, 0) and chunks(p[0], "bplayer") == (0x6, 0, 0)
    over, res = s.is_over(p)
    assert over[0] and res[0] == 1
    p = replay(s, [1, 2, 3, 5, 4, 6, 8, 7, 9])
    assert chunks(p[0], "bplayer") == (0x72, 0, 0) and chunks(p[0], "bopponent") == (0x18D, 0, 0)
    over, res = s.is_over(p)
    assert over[0] and res[0] == 0
    s = oracle.Spec(oracle.GOBANG, 9, 5)
    p = replay(s, [1, 10, 2, 11, 3, 12, 4, 13, 5])
    assert chunks(p[0], "bopponent") == (0x1F, 0, 0) and chunks(p[0], "bplayer") == (0x1E00, 0, 0)
    over, res = s.is_over(p)
    assert over[0] and res[0] == 1
    p = replay(s, [37, 1, 45, 2, 53, 3, 61, 4, 29])   # vertical wrap guard
    assert chunks(p[0], "bopponent") == (0x1010101010000000, 0, 0)
    over, res = s.is_over(p)
    assert not over[0]


def test_appendix_b_hex():
    s = oracle.Spec(oracle.HEX, 7)
    p0 = s.position(1)
    assert chunks(p0[0], "bplayer") == (0xFC, 0, 0)
    assert chunks(p0[0], "bopponent") == (0x101010101010000, 0, 0)
    assert p0[0]["player"] == 1 and p0[0]["aux"] == 49
    moves = [1, 7, 8, 14, 15, 21, 22, 28, 29, 35, 36, 42, 43]
    pos = p0
    for i, m in enumerate(moves):
        pos = s.play(pos, m)
        over, res = s.is_over(pos)
        assert bool(over[0]) == (i == len(moves) - 1)
    assert chunks(pos[0], "bplayer") == (0x181818181818000, 0, 0)
    assert chunks(pos[0], "bopponent") == (0x2020202020202FC, 0, 0)
    assert res[0] == 1


def test_appendix_b_reversi():
    s = oracle.Spec(oracle.REVERSI8)
    p0 = s.position(1)
    assert chunks(p0[0], "bplayer") == (0x810000000, 0, 0)
    assert chunks(p0[0], "bopponent") == (0x1008000000, 0, 0)
    assert chunks(p0[0], "legalplay") == (0x102004080000, 0, 0)
    assert [a + 1 for a in np.nonzero(s.legal(p0)[0])[0]] == [20, 27, 38, 45]
    p = replay(s, [20, 19])
    assert chunks(p[0], "bplayer") == (0x810080000, 0, 0) and chunks(p[0], "bopponent") == (0x1008040000, 0, 0)
    assert p[0]["player"] == 1
    assert [a + 1 for a in np.nonzero(s.legal(p)[0])[0]] == [18, 27, 38, 45]
    s6 = oracle.Spec(oracle.REVERSI6)
    q0 = s6.position(1)
    assert chunks(q0[0], "bplayer") == (0x108000, 0, 0) and chunks(q0[0], "bopponent") == (0x204000, 0, 0)
    assert chunks(q0[0], "legalplay") == (0x8402100, 0, 0)


def test_game_constants():
    want = {"connect4": (7, 42, 42, 42, 104), "ttt": (9, 9, 9, 9, 104), "gobang9": (81, 81, 81, 81, 104),
            "hex7": (49, 64, 64, 49, 104), "reversi8": (65, 64, 64, 70, 152), "reversi6": (37, 36, 36, 50, 152)}
    for name, w in want.items():
        s = oracle.Spec(*GAME_SPECS[name])
        assert (s.A, s.VS, s.FS, s.maxLen, s.pos_bytes) == w, name


# ---------------- (2) committed golden traces ----------------
def test_golden_traces():
    with open(os.path.join(HERE, "golden", "kat_games.json")) as f:
        cases = json.load(f)
    assert len(cases) >= 15
    for case in cases:
        s = oracle.Spec(*case["spec"])
        pos = s.position(1)
        for step in case["trace"]:
            if step["move"] is not None:
                assert s.can_play(pos, step["move"])[0]
                pos = s.play(pos, step["move"])
            assert [hex(c) for c in chunks(pos[0], "bplayer")] == step["bplayer"], case["name"]
            assert [hex(c) for c in chunks(pos[0], "bopponent")] == step["bopponent"]
            if step["legalplay"] is not None:
                assert [hex(c) for c in chunks(pos[0], "legalplay")] == step["legalplay"]
            assert int(pos[0]["player"]) == step["player"]
            if s.pos_bytes == 104:
                assert int(pos[0]["aux"]) == step["aux"]
            over, res = s.is_over(pos)
            assert bool(over[0]) == step["over"]
            if step["over"]:
                assert int(res[0]) == step["result"]
            assert [a + 1 for a in np.nonzero(s.legal(pos)[0])[0]] == step["legal"]


# ---------------- (3) random games vs pyref and naive rules ----------------
@pytest.mark.parametrize("name", ["connect4", "ttt", "gobang9", "gobang5", "hex7", "hex5", "reversi8", "reversi6"])
def test_random_games_three_way(name):
    spec_t = GAME_SPECS[name]
    s = oracle.Spec(*spec_t)
    g = pyref.make(*spec_t)
    rng = random.Random(1234 + sum(map(ord, name)))
    ngames = 12 if name in ("reversi8", "gobang9", "hex7") else 25
    for _ in range(ngames):
        pos, pp, nv = s.position(1), g.position(), naive_rules.make(*spec_t)
        for ply in range(400):
            # positions agree bit for bit
            assert chunks(pos[0], "bplayer") == pp.bplayer.chunks()
            assert chunks(pos[0], "bopponent") == pp.bopponent.chunks()
            if pp.legalplay is not None:
                assert chunks(pos[0], "legalplay") == pp.legalplay.chunks()
            assert int(pos[0]["player"]) == pp.player == nv.to_move
            over, res = s.is_over(pos)
            over2, res2 = g.is_over(pp)
            over3, res3 = nv.over()
            assert bool(over[0]) == bool(over2) == bool(over3), (name, ply)
            if over[0]:
                assert int(res[0]) == int(res2) == int(res3), (name, ply)
                break
            legal = [int(a) + 1 for a in np.nonzero(s.legal(pos)[0])[0]]
            assert legal == [a for a in range(1, g.A + 1) if g.can_play(pp, a)]
            assert legal == [a for a in range(1, g.A + 1) if nv.legal(a)]
            assert legal, "no legal move in a non-terminal position"
            enc = s.encode(pos)[0]
            assert list(enc) == g.encode(pp)
            m = rng.choice(legal)
            pos, pp = s.play(pos, m), g.play(pp, m)
            nv.play(m)
        else:
            raise AssertionError("game did not end")


def test_bitboard_ops_vs_bigint():
    rng = random.Random(7)
    for (h, w) in [(6, 7), (3, 3), (9, 9), (8, 8), (13, 13), (10, 10), (6, 6)]:
        ln = h * w
        for _ in range(40):
            bits = rng.getrandbits(ln)
            bb = np.zeros(1, dtype=oracle.BB)
            bb["chunks"][0] = [bits & (2**64 - 1), (bits >> 64) & (2**64 - 1), (bits >> 128) & (2**64 - 1)]
            bb["len"], bb["dims"][0] = ln, [h, w]
            ref = pyref.BB(bits, h, w)
            n = rng.randrange(0, 15)
            for op, want in (("<<", ref.shl(n)), (">>>", ref.shr(n)), ("right", ref.right()), ("left", ref.left()),
                             ("down", ref.down()), ("up", ref.up()), ("~", ref.inv())):
                got = oracle.bb_op(op, bb[0], n)
                assert tuple(int(x) for x in got["chunks"]) == want.chunks(), (op, h, w, n)
    # shifts by >= 64 (Bitboard.jl:89-98): '<<' moves whole chunks, then masks
    bb = np.zeros(1, dtype=oracle.BB)
    bb["chunks"][0] = [0x8000000000000001, 0x3, 0]
    bb["len"], bb["dims"][0] = 169, [13, 13]
    got = oracle.bb_op("<<", bb[0], 64)
    assert tuple(int(x) for x in got["chunks"]) == (0, 0x8000000000000001, 0x3)
    got = oracle.bb_op("<<", bb[0], 65)
    assert tuple(int(x) for x in got["chunks"]) == (0, 0x2, 0x7)


def test_philox_known_answer():
    # Random123 kat_vectors: philox4x32-10, counter 0 key 0; all-ones; pi digits
    assert oracle.philox([0, 0, 0, 0], 0, 0) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert oracle.philox([0xFFFFFFFF] * 4, 0xFFFFFFFF, 0xFFFFFFFF) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert oracle.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], 0xA4093822, 0x299F31D0) == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    us = [oracle.uniform(0, g, 0, r, d) for g in range(4) for r in range(4) for d in range(8)]
    assert all(0.0 < u <= 1.0 for u in us) and len(set(us)) == len(us)
