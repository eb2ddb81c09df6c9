"""Interactive play: the `FMCTS.MctsContext` call surface of fast_mcts.jl and the `testvsordi` drivers of
testHex.jl / testgobang.jl / testrev6.jl / testrev8.jl, served by the batched GPU search with one game.

The reference plays its interactive games with a CPU tree (`fast_mcts.jl:262-289`) and keeps the GPU
variant as a commented-out line in every driver (`testrev8.jl:24`, `testgobang.jl:25`, `testHex.jl:39`):

    _,p = mcts_gpu.mcts_single(actor, readout, 256, vnodes, vnodesStats, leaf, newindex, 1, training=false, cpuct=1.5, …)

That line is what runs here: `re_init` with the one position, `mcts_single` with `L = 1` and `training=false`,
`policy_final` as `p`.  `v` is the root's mean backed-up value Σ_a q[a]·n[a] / readout — `extractRoot`'s
`sum(action.w)/N` (`fast_mcts.jl:293-302`; the root itself is visited once without an action, so N = readout).
Differences from the CPU tree that follow from using `mcts_single`: at most 255 read-outs (node ids are 8 bit),
`p` is the regularised policy π̄ the last descent computed at the root (`copy_pol`, `mcts_gpu.jl:330-339`) rather
than one re-solved after the last backup, and `komi` is accepted and ignored as the reference's `evaluate` ignores it
(`fast_mcts.jl:126-142`).  There is no CPU path: without the CUDA library the constructor raises.
"""
from __future__ import annotations

from typing import Callable, Iterable, Iterator, List, Optional, Tuple, Union

import numpy as np

from . import _lib
from .densenet import SNetwork2
from .game import GameSpec
from .mcts_gpu import Context

MAX_READOUT = 255


class MctsContext:
    """`FMCTS.MctsContext(c, nn, prealloc)` (fast_mcts.jl:262-266): `ctx(pos, readout) -> (p, v)`."""

    def __init__(self, c: float, nn: SNetwork2, spec: GameSpec, *, readout_max: int = MAX_READOUT, device: int = 0,
                 nn_mode: int = _lib.NN_FP16_TC, seed: int = 0):
        if not 1 <= readout_max <= MAX_READOUT:
            raise ValueError(f"readout_max must be 1..{MAX_READOUT}")
        self.c, self.spec, self.readout_max, self.seed = float(c), spec, readout_max, seed
        self.ctx = Context(spec, readout_max, 1, nn.width, nn.blocks, device, nn_mode)
        self.ctx.set_weights(nn, 0)
        self.calls = 0

    def close(self):
        self.ctx.close()

    def _one(self, pos) -> np.ndarray:
        a = np.ascontiguousarray(np.asarray(pos, self.spec.position_dtype)).reshape(-1)
        if a.shape[0] != 1:
            raise ValueError("MctsContext searches one position per call")
        return a

    def __call__(self, pos, readout: int, komi: float = 0) -> Tuple[np.ndarray, float]:
        """The call operator of fast_mcts.jl:270-289: `readout` simulations from `pos`, then `extractRoot`."""
        if not 1 <= readout <= self.readout_max:
            raise ValueError(f"readout must be 1..{self.readout_max} (capacity of this context)")
        a = self._one(pos)
        over, _ = self.ctx.isOver(a)
        if over[0]:
            raise ValueError("position is terminal: nothing to search (the drivers test isOver first)")
        self.ctx.re_init(a, np.asarray([self.calls], np.uint32))        # a fresh random stream per call
        self.ctx.mcts_single(readout, 1, training=False, cpuct=self.c, seed=self.seed, ply=0)
        self.calls += 1
        p, _ = self.ctx.roots(1)
        t = self.ctx.tree()
        return p[0], root_value(t["q"][0, 0], t["visits"][0, 0], readout)


def root_value(q: np.ndarray, n: np.ndarray, readout: int) -> float:
    """`sum(action.w for action in node.actions)/N` (fast_mcts.jl:300) from the mean values and visit counts of the root."""
    return float(np.sum(q.astype(np.float64) * n.astype(np.float64)) / readout)


# ------------------------------------------------------------------------------------------------
# move notation of the four drivers
# ------------------------------------------------------------------------------------------------
_LETTERS = "abcdefgh"


def move_dictionaries(spec: GameSpec):
    """(text -> action, action -> text), actions 1-based.
    Reversi: `dic_coups` / `dic_coups_inverse` of testrev8.jl:1-13 and testrev6.jl (letter = row block, digit = offset, "p" = pass);
    Hex: `generate_dict` of testHex.jl:5-17 (column letter A.. + row); Gobang: the two-digit `x y` code of testgobang.jl:39-47
    (c = N·x + y + 1); Connect4 (no driver in the reference): the column number."""
    g, N = spec.game, spec.N
    fwd, inv = {}, {}
    if g in (_lib.REVERSI8, _lib.REVERSI6):
        side = 8 if g == _lib.REVERSI8 else 6
        for case in range(1, side + 1):
            for zone in range(1, side + 1):
                fwd[f"{_LETTERS[case - 1]}{zone}"] = side * (case - 1) + zone
                inv[zone + side * (case - 1)] = f"{_LETTERS[case - 1]}{zone}"
        fwd["p"] = side * side + 1
        inv[side * side + 1] = "pass"
    elif g == _lib.HEX:
        for c in range(1, N * N + 1):
            col = (c - 1) // N + 1
            row = c - N * (col - 1)
            coup = f"{chr(ord('A') + col - 1)}{row}"
            inv[c], fwd[coup] = coup, c
    elif g == _lib.GOBANG:
        for x in range(N):
            for y in range(N):
                fwd[str(10 * x + y)] = N * x + y + 1
                inv[N * x + y + 1] = str(10 * x + y)
    else:
        for c in range(1, spec.maxActions + 1):
            fwd[str(c)], inv[c] = c, str(c)
    return fwd, inv


MoveSource = Union[Iterable[Union[int, str]], Callable[[np.ndarray], Union[int, str]]]


def testvsordi(actor: SNetwork2, readout: int, player: int = -1, *, spec: GameSpec, pos=None, moves: MoveSource = (), cpuct: float = 1.5,
               log: Optional[Callable[[str], None]] = print, device: int = 0, nn_mode: int = _lib.NN_FP16_TC, seed: int = 0):
    """`testvsordi(actor, readout, joueur; pos)` (testrev8.jl:14-58, testgobang.jl:9-59, testHex.jl:21-66): the engine plays the side
    `player` (±1, `game.player == joueur`) with `argmax(p)` after `readout` simulations at cpuct 1.5; the other side's moves come from
    `moves` — an iterable of actions or of move texts in the driver's notation, or a callable `moves(game)` — where the reference
    calls `readline()`.  Returns (history of 1-based actions, winner as `isOver(game)[2]`)."""
    fwd, inv = move_dictionaries(spec)
    puct = MctsContext(cpuct, actor, spec, readout_max=max(1, min(MAX_READOUT, readout)), device=device, nn_mode=nn_mode, seed=seed)
    say = log or (lambda s: None)
    it: Optional[Iterator] = None if callable(moves) else iter(moves)
    try:
        ctx = puct.ctx
        game = ctx.Position(1) if pos is None else puct._one(pos)
        history: List[int] = []
        while True:
            over, result = ctx.isOver(game)
            if over[0]:
                break
            if int(game["player"][0]) == player:
                p, v = puct(game, readout)
                c = int(np.argmax(p)) + 1
                say(f"coup: {inv[c]}")
                say(f"situation: {v}")
            else:
                say("coups d internet")
                m = moves(game) if it is None else next(it)
                c = fwd[m] if isinstance(m, str) else int(m)
            if not ctx.canPlay(game)[0, c - 1]:
                raise ValueError(f"coup non valide: {c}")
            history.append(c)
            game = ctx.play(game, c)
        w = int(result[0])
        say("winner: puct" if w * player > 0 else "match nul" if w == 0 else "winner: internet")
        return history, w
    finally:
        puct.close()
