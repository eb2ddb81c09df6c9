"""Second, independent restatement of the reference game plugins — TEST INFRASTRUCTURE ONLY.

Pure Python, one arbitrary-precision integer per bitboard (bit i-1 of the integer = Julia
linear index i).  Written from the Julia text (Bitboard.jl, 4IARow.jl, Gobang.jl, Hex.jl,
Reversi6x6.jl, Reversi8x8.jl) without looking at oracle.cpp's chunked arithmetic, so that the
two restatements pin each other (the reference ships no vectors of its own).
Small cases only (pure-Python loops).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple


@dataclass(frozen=True)
class BB:
    """Bitboard.jl:5-9 — bits + len + dims."""
    bits: int
    h: int   # dims[1]
    w: int   # dims[2]

    @property
    def len(self) -> int:
        return self.h * self.w

    def _mask(self) -> int:                      # Bitboard.jl:33-41
        return (1 << self.len) - 1

    def get(self, i: int) -> bool:               # :47-52
        return (self.bits >> (i - 1)) & 1 == 1

    def get2(self, r: int, c: int) -> bool:      # :54-57
        return self.get(self.h * (c - 1) + r)

    def set(self, x: bool, i: int) -> "BB":      # :60-74
        b = self.bits | (1 << (i - 1)) if x else self.bits & ~(1 << (i - 1))
        return BB(b, self.h, self.w)

    def set2(self, x: bool, r: int, c: int) -> "BB":
        return self.set(x, self.h * (c - 1) + r)

    def shl(self, n: int) -> "BB":               # :85-107 (for n < 64 this is a plain masked shift of the 192-bit value)
        assert 0 <= n < 64
        return BB(((self.bits << n) & ((1 << 192) - 1)) & self._mask(), self.h, self.w)

    def shr(self, n: int) -> "BB":               # :110-133 (n < 64)
        assert 0 <= n < 64
        return BB((self.bits >> n) & self._mask(), self.h, self.w)

    def right(self) -> "BB":                     # :135-138
        return self.shl(self.h)

    def left(self) -> "BB":                      # :141-144
        return self.shr(self.h)

    def down(self) -> "BB":                      # :146-160
        b = self.shl(1).bits
        for i in range(1, self.len + 1, self.h):
            b &= ~(1 << (i - 1))
        return BB(b, self.h, self.w)

    def up(self) -> "BB":                        # :162-176
        b = self.shr(1).bits
        for i in range(self.h, self.len + 1, self.h):
            b &= ~(1 << (i - 1))
        return BB(b, self.h, self.w)

    def num_bit(self) -> int:                    # :177-180
        return bin(self.bits).count("1")

    def inv(self) -> "BB":                       # :182-187
        return BB((~self.bits) & self._mask(), self.h, self.w)

    def __and__(self, o): return BB(self.bits & o.bits, self.h, self.w)
    def __or__(self, o): return BB(self.bits | o.bits, self.h, self.w)
    def __xor__(self, o): return BB(self.bits ^ o.bits, self.h, self.w)

    def chunks(self) -> Tuple[int, int, int]:
        m = (1 << 64) - 1
        return (self.bits & m, (self.bits >> 64) & m, (self.bits >> 128) & m)


@dataclass(frozen=True)
class Position:
    bplayer: BB
    bopponent: BB
    player: int
    aux: int = 0              # round / lp
    legalplay: BB = None      # Reversi only


class Game:
    name = ""
    A = VS = FS = maxLen = 0
    def position(self) -> Position: ...
    def can_play(self, pos: Position, a: int) -> bool: ...
    def play(self, pos: Position, a: int) -> Position: ...
    def is_over(self, pos: Position) -> Tuple[bool, int]: ...

    def encode(self, pos: Position):             # mcts_gpu.jl:202-223
        return [1.0 if pos.bplayer.get(j) else 0.0 for j in range(1, self.VS + 1)] + \
               [1.0 if pos.bopponent.get(j) else 0.0 for j in range(1, self.VS + 1)]


def _row_test(board: BB, nvict: int) -> bool:    # 4IARow.jl:47-78 / Gobang.jl:36-67
    b = board
    for _ in range(nvict - 1):
        b = b & b.right()
    if b.num_bit():
        return True
    b = board
    for _ in range(nvict - 1):
        b = b & b.down()
    if b.num_bit():
        return True
    b = board
    for _ in range(nvict - 1):
        b = b & b.right().down()
    if b.num_bit():
        return True
    b = board
    for _ in range(nvict - 1):
        b = b & b.down().left()
    return b.num_bit() != 0


class Connect4(Game):                            # 4IARow.jl
    name = "connect4"
    A, VS, FS, maxLen = 7, 42, 42, 42

    def position(self):
        return Position(BB(0, 6, 7), BB(0, 6, 7), 1, 1)

    def can_play(self, pos, col):
        return (not pos.bplayer.get2(1, col)) and (not pos.bopponent.get2(1, col))

    def play(self, pos, col):
        free = 1
        empty = (pos.bplayer | pos.bopponent).inv()
        for i in range(1, 7):
            if empty.get2(i, col):
                free = i
            else:
                break
        c = 6 * (col - 1) + free
        return Position(pos.bopponent, pos.bplayer.set(True, c), -pos.player, pos.aux + 1)

    def is_over(self, pos):
        if _row_test(pos.bopponent, 4):
            return True, -pos.player
        return pos.bplayer.num_bit() + pos.bopponent.num_bit() == 42, 0


class Gobang(Game):                              # Gobang.jl
    name = "gobang"

    def __init__(self, N, Nvict):
        self.N, self.Nvict = N, Nvict
        self.A = self.VS = self.FS = self.maxLen = N * N

    def position(self):
        return Position(BB(0, self.N, self.N), BB(0, self.N, self.N), 1, 0)

    def can_play(self, pos, col):
        return (not pos.bplayer.get(col)) and (not pos.bopponent.get(col))

    def play(self, pos, col):
        return Position(pos.bopponent, pos.bplayer.set(True, col), -pos.player, pos.aux + 1)

    def is_over(self, pos):
        if _row_test(pos.bopponent, self.Nvict):
            return True, -pos.player
        return pos.bplayer.num_bit() + pos.bopponent.num_bit() == self.N * self.N, 0


class Hex(Game):                                 # Hex.jl
    name = "hex"

    def __init__(self, N):
        self.N = N
        self.A = self.maxLen = N * N
        self.VS = self.FS = (N + 1) * (N + 1)

    def position(self):
        N = self.N
        sx = BB(0, N + 1, N + 1)
        so = BB(0, N + 1, N + 1)
        for i in range(3, N + 2):
            sx = sx.set2(True, i, 1)
            so = so.set2(True, 1, i)
        return Position(sx, so, 1, N * N)

    def _idx(self, col):
        N = self.N
        x = (col - 1) // N
        y = col - N * x
        return (N + 1) * (x + 1) + y + 1

    def can_play(self, pos, col):
        i = self._idx(col)
        return (not pos.bplayer.get(i)) and (not pos.bopponent.get(i))

    def play(self, pos, col):
        return Position(pos.bopponent, pos.bplayer.set(True, self._idx(col)), -pos.player, pos.aux - 1)

    def is_over(self, pos):
        N = self.N
        a = pos.bopponent
        for j in range(1, 2 * N - 1):
            b = a.up()
            c = b.right()
            a = ((a & (b | c)) | (b & c)).down()
            if pos.player == 1:
                for k in range(3 + j, N + 2):
                    a = a.set2(True, 1, k)
        return a.get2(N + 1, N + 1), -pos.player


class Reversi(Game):                             # Reversi8x8.jl / Reversi6x6.jl
    name = "reversi"

    def __init__(self, n):
        self.n = n
        self.VS = self.FS = n * n
        self.A = n * n + 1
        self.maxLen = 70 if n == 8 else 50
        e = BB(0, n, n)
        if n == 8:
            self.starto = e.set2(True, 4, 5).set2(True, 5, 4)
            self.startp = e.set2(True, 5, 5).set2(True, 4, 4)
        else:
            self.starto = e.set2(True, 4, 3).set2(True, 3, 4)
            self.startp = e.set2(True, 3, 3).set2(True, 4, 4)

    _dirs_legal = ("up", "down", "left", "right", "diaghg", "diagbg", "diaghd", "diagbd")
    _dirs_flip = ("up", "down", "left", "right", "diaghd", "diaghg", "diagbd", "diagbg")

    @staticmethod
    def _dir(name, x: BB) -> BB:
        if name == "up": return x.up()
        if name == "down": return x.down()
        if name == "left": return x.left()
        if name == "right": return x.right()
        if name == "diaghd": return x.right().up()
        if name == "diaghg": return x.left().up()
        if name == "diagbd": return x.right().down()
        if name == "diagbg": return x.left().down()
        raise KeyError(name)

    def _legal_play(self, tj, ta, d):
        vide = tj.inv() & ta.inv()
        moves = BB(0, self.n, self.n)
        cand = self._dir(d, tj) & ta
        while cand.num_bit():
            moves = moves | (vide & self._dir(d, cand))
            cand = ta & self._dir(d, cand)
        return moves

    def legalplay(self, tj, ta):
        m = BB(0, self.n, self.n)
        for d in self._dirs_legal:
            m = m | self._legal_play(tj, ta, d)
        return m

    def _flippar(self, tj, ta, play, d):
        cand = self._dir(d, play) & ta
        toflip = cand
        while cand.num_bit():
            cand = ta & self._dir(d, cand)
            toflip = toflip | cand
        if (self._dir(d, toflip) & tj).num_bit():
            return toflip
        return BB(0, self.n, self.n)

    def _flip(self, tj, ta, play):
        test = BB(0, self.n, self.n).set(True, play)
        h = BB(0, self.n, self.n)
        for d in self._dirs_flip:
            h = h | self._flippar(tj, ta, test, d)
        return h

    def position(self):
        return Position(self.starto, self.startp, 1, 0, self.legalplay(self.starto, self.startp))

    def can_play(self, pos, c):
        if c == self.A:
            return pos.legalplay.num_bit() == 0
        return pos.legalplay.get(c)

    def play(self, pos, c):
        tj, ta = pos.bplayer, pos.bopponent
        if c == self.A:
            return Position(pos.bopponent, pos.bplayer, -pos.player, 0, self.legalplay(ta, tj))
        h = self._flip(tj, ta, c)
        tj = tj ^ h
        ta = ta ^ h
        tj = tj.set(True, c)
        return Position(ta, tj, -pos.player, 0, self.legalplay(ta, tj))

    def is_over(self, pos):
        test = pos.bplayer.num_bit() - pos.bopponent.num_bit()
        over = pos.legalplay.num_bit() == 0 and self.legalplay(pos.bopponent, pos.bplayer).num_bit() == 0
        sg = (test > 0) - (test < 0)
        if self.n == 6 and not over:
            return False, 0                      # Reversi6x6.jl:110-111
        return over, sg * pos.player


def make(game: int, N: int = 0, Nvict: int = 0) -> Game:
    return {0: lambda: Connect4(), 1: lambda: Gobang(N, Nvict), 2: lambda: Hex(N), 3: lambda: Reversi(8), 4: lambda: Reversi(6)}[game]()
