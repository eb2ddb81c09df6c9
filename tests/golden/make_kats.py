"""Regenerates tests/golden/kat_games.json — known-answer vectors for the game plugins.

The reference ships no test vectors and Julia is not available, so these are
DERIVED vectors: the move sequences of SURVEY.md Appendix B replayed through oracle/pyref.py
(the big-integer restatement of the Julia text).  tests/test_oracle_games.py then requires
oracle.cpp (and, on the GPU, the CUDA kernels) to reproduce them bit for bit, and checks the
hand-verified Appendix-B constants directly.

    python tests/golden/make_kats.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pyref  # noqa: E402

CASES = [
    ("connect4", (0, 0, 0), [4, 4, 5, 5, 6, 6, 7]),
    ("connect4", (0, 0, 0), [1, 2, 1, 2, 1, 2, 1]),
    ("connect4", (0, 0, 0), [4, 4, 4, 4, 4, 4, 3]),
    ("connect4", (0, 0, 0), [1, 2, 2, 3, 3, 4, 3, 4, 4, 6, 4]),
    ("ttt", (1, 3, 3), [1, 2, 5, 3, 9]),
    ("ttt", (1, 3, 3), [1, 2, 3, 5, 4, 6, 8, 7, 9]),
    ("gobang9", (1, 9, 5), [1, 10, 2, 11, 3, 12, 4, 13, 5]),
    ("gobang9", (1, 9, 5), [37, 1, 45, 2, 53, 3, 61, 4, 29]),
    ("gobang9", (1, 9, 5), [1, 81, 11, 80, 21, 79, 31, 78, 41]),
    ("hex7", (2, 7, 0), [1, 7, 8, 14, 15, 21, 22, 28, 29, 35, 36, 42, 43]),
    ("hex7", (2, 7, 0), [25, 1, 2, 3, 4, 5, 6, 7, 8]),
    ("reversi8", (3, 0, 0), [20, 19]),
    ("reversi8", (3, 0, 0), "lowest:60"),     # always the lowest legal action, up to 60 plies (exercises passes at the end)
    ("reversi8", (3, 0, 0), "highest:64"),
    ("reversi6", (4, 0, 0), "lowest:40"),
    ("hex7", (2, 7, 0), "lowest:49"),
    ("connect4", (0, 0, 0), "lowest:42"),
]


def expand(g, moves):
    if not isinstance(moves, str):
        return moves
    kind, n = moves.split(":")
    pos, out = g.position(), []
    for _ in range(int(n)):
        if g.is_over(pos)[0]:
            break
        legal = [a for a in range(1, g.A + 1) if g.can_play(pos, a)]
        m = legal[0] if kind == "lowest" else legal[-1]
        out.append(m)
        pos = g.play(pos, m)
    return out


def main():
    out = []
    for name, spec, moves in CASES:
        g = pyref.make(*spec)
        moves = expand(g, moves)
        pos = g.position()
        trace = []
        for m in [None] + moves:
            if m is not None:
                assert g.can_play(pos, m), (name, moves, m)
                pos = g.play(pos, m)
            over, res = g.is_over(pos)
            trace.append(dict(
                move=m, bplayer=[hex(c) for c in pos.bplayer.chunks()], bopponent=[hex(c) for c in pos.bopponent.chunks()],
                legalplay=None if pos.legalplay is None else [hex(c) for c in pos.legalplay.chunks()],
                player=pos.player, aux=pos.aux, over=bool(over), result=int(res) if over else None,
                legal=[a for a in range(1, g.A + 1) if g.can_play(pos, a)]))
        out.append(dict(name=name, spec=list(spec), moves=moves, trace=trace))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat_games.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path, len(out), "cases")


if __name__ == "__main__":
    main()
