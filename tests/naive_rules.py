"""Third check on game semantics: plain rule implementations on 2-D grids (no bitboards).

Cells are addressed the way the reference's linear indices map to (row, col):
index i (1-based) -> row = (i-1) % H + 1, col = (i-1) // H + 1 (column-major, Bitboard.jl:54-57).
Each class tracks stones by colour (+1 first player, -1 second) and exposes
legal(a), play(a), over() -> (bool, winner colour or 0), to_move.
"""
from collections import deque


class NaiveConnect4:
    H, W, A = 6, 7, 7

    def __init__(self):
        self.g = [[0] * (self.W + 1) for _ in range(self.H + 1)]  # g[row][col], row 1 = top
        self.to_move = 1
        self.last = None

    def legal(self, col):
        return self.g[1][col] == 0

    def play(self, col):
        row = max(r for r in range(1, self.H + 1) if all(self.g[k][col] == 0 for k in range(1, r + 1)))
        self.g[row][col] = self.to_move
        self.last = (row, col)
        self.to_move = -self.to_move

    def _wins(self, colour, k):
        H, W = self.H, self.W
        for r in range(1, H + 1):
            for c in range(1, W + 1):
                for dr, dc in ((0, 1), (1, 0), (1, 1), (1, -1)):
                    ok = True
                    for t in range(k):
                        rr, cc = r + dr * t, c + dc * t
                        if not (1 <= rr <= H and 1 <= cc <= W) or self.g[rr][cc] != colour:
                            ok = False
                            break
                    if ok:
                        return True
        return False

    def over(self):
        mover = -self.to_move
        if self._wins(mover, 4):
            return True, mover
        full = all(self.g[r][c] != 0 for r in range(1, self.H + 1) for c in range(1, self.W + 1))
        return full, 0

    def cell(self, i):
        r, c = (i - 1) % self.H + 1, (i - 1) // self.H + 1
        return self.g[r][c]


class NaiveGobang(NaiveConnect4):
    def __init__(self, N, k):
        self.H = self.W = N
        self.A = N * N
        self.k = k
        super().__init__()

    def legal(self, a):
        return self.cell(a) == 0

    def play(self, a):
        r, c = (a - 1) % self.H + 1, (a - 1) // self.H + 1
        self.g[r][c] = self.to_move
        self.to_move = -self.to_move

    def over(self):
        mover = -self.to_move
        if self._wins(mover, self.k):
            return True, mover
        full = all(self.g[r][c] != 0 for r in range(1, self.H + 1) for c in range(1, self.W + 1))
        return full, 0


class NaiveHex:
    """Action c -> x=(c-1)//N, y=c-N*x.  Neighbours (0,±1),(±1,0),(+1,-1),(-1,+1).
    First player (+1) joins x=0 with x=N-1; second (-1) joins y=1 with y=N."""

    def __init__(self, N):
        self.N, self.A = N, N * N
        self.s = {}
        self.to_move = 1

    def legal(self, c):
        return c not in self.s

    def play(self, c):
        self.s[c] = self.to_move
        self.to_move = -self.to_move

    def _xy(self, c):
        x = (c - 1) // self.N
        return x, c - self.N * x

    def _connected(self, colour):
        N = self.N
        cells = {self._xy(c) for c, v in self.s.items() if v == colour}
        if colour == 1:
            start = [p for p in cells if p[0] == 0]
            goal = lambda p: p[0] == N - 1
        else:
            start = [p for p in cells if p[1] == 1]
            goal = lambda p: p[1] == N
        seen, dq = set(start), deque(start)
        while dq:
            p = dq.popleft()
            if goal(p):
                return True
            for dx, dy in ((0, 1), (0, -1), (1, 0), (-1, 0), (1, -1), (-1, 1)):
                nb = (p[0] + dx, p[1] + dy)
                if nb in cells and nb not in seen:
                    seen.add(nb)
                    dq.append(nb)
        return False

    def over(self):
        mover = -self.to_move
        return self._connected(mover), mover


class NaiveReversi:
    def __init__(self, n):
        self.n, self.A = n, n * n + 1
        self.g = [[0] * (n + 2) for _ in range(n + 2)]
        # the reference's start: side to move (+1) owns starto (Reversi8x8.jl:10-14,82)
        if n == 8:
            mine, theirs = [(4, 5), (5, 4)], [(5, 5), (4, 4)]
        else:
            mine, theirs = [(4, 3), (3, 4)], [(3, 3), (4, 4)]
        for r, c in mine:
            self.g[r][c] = 1
        for r, c in theirs:
            self.g[r][c] = -1
        self.to_move = 1

    def _flips(self, r, c, colour):
        if self.g[r][c] != 0:
            return []
        out = []
        for dr in (-1, 0, 1):
            for dc in (-1, 0, 1):
                if dr == 0 and dc == 0:
                    continue
                line, rr, cc = [], r + dr, c + dc
                while 1 <= rr <= self.n and 1 <= cc <= self.n and self.g[rr][cc] == -colour:
                    line.append((rr, cc))
                    rr, cc = rr + dr, cc + dc
                if line and 1 <= rr <= self.n and 1 <= cc <= self.n and self.g[rr][cc] == colour:
                    out += line
        return out

    def _moves(self, colour):
        return [(r, c) for r in range(1, self.n + 1) for c in range(1, self.n + 1) if self._flips(r, c, colour)]

    def legal(self, a):
        if a == self.A:
            return not self._moves(self.to_move)
        r, c = (a - 1) % self.n + 1, (a - 1) // self.n + 1
        return bool(self._flips(r, c, self.to_move))

    def play(self, a):
        if a != self.A:
            r, c = (a - 1) % self.n + 1, (a - 1) // self.n + 1
            for rr, cc in self._flips(r, c, self.to_move):
                self.g[rr][cc] = self.to_move
            self.g[r][c] = self.to_move
        self.to_move = -self.to_move

    def over(self):
        if self._moves(1) or self._moves(-1):
            return False, 0
        d = sum(self.g[r][c] for r in range(1, self.n + 1) for c in range(1, self.n + 1))
        return True, (d > 0) - (d < 0)


def make(game, N=0, Nvict=0):
    return {0: lambda: NaiveConnect4(), 1: lambda: NaiveGobang(N, Nvict), 2: lambda: NaiveHex(N), 3: lambda: NaiveReversi(8),
            4: lambda: NaiveReversi(6)}[game]()
