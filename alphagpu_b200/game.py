"""Host-side mirror of the reference's game-plugin surface (README.md:65-74; 4IARow.jl:2):
`Position`, `canPlay`, `play`, `isOver` and the constants `VectorizedState`, `FeatureSize`,
`maxActions`, `maxLengthGame`.  Positions are numpy structured arrays with the Julia isbits layout
(the bytes a Julia `Vector{Position}` holds), and every operation runs on the GPU through the C ABI.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib

# Bitboard.jl:5-9 / 4IARow.jl:16-21 / Reversi8x8.jl:73-78
BITBOARD = np.dtype([("chunks", "<u8", (3,)), ("len", "<i8"), ("dims", "<i8", (2,))])
POSITION2 = np.dtype([("bplayer", BITBOARD), ("bopponent", BITBOARD), ("player", "i1"), ("aux", "i1"), ("pad", "i1", (6,))])
POSITION3 = np.dtype([("bplayer", BITBOARD), ("bopponent", BITBOARD), ("legalplay", BITBOARD), ("player", "i1"), ("pad", "i1", (7,))])
assert POSITION2.itemsize == 104 and POSITION3.itemsize == 152

GAME_IDS = {"4IARow": _lib.CONNECT4, "connect4": _lib.CONNECT4, "GoBang": _lib.GOBANG, "gobang": _lib.GOBANG, "Hex": _lib.HEX, "hex": _lib.HEX,
            "Reversi8x8": _lib.REVERSI8, "reversi8": _lib.REVERSI8, "Reversi6x6": _lib.REVERSI6, "reversi6": _lib.REVERSI6}


@dataclass(frozen=True)
class GameSpec:
    """One game plugin: module name of the reference + the Main.N / Main.Nvict constants it reads."""
    game: int
    N: int = 0
    Nvict: int = 0

    @staticmethod
    def named(name: str, N: int = 0, Nvict: int = 0) -> "GameSpec":
        return GameSpec(GAME_IDS[name], N, Nvict)

    def _info(self):
        gi = _lib.GameInfo()
        rc = _lib.load().agpu_game_info_get(self.game, self.N, self.Nvict, C.byref(gi))
        if rc != _lib.OK:
            raise ValueError(f"unsupported game spec {self}")
        return gi

    @property
    def maxActions(self): return self._info().max_actions
    @property
    def VectorizedState(self): return self._info().vectorized_state
    @property
    def FeatureSize(self): return self._info().feature_size
    @property
    def maxLengthGame(self): return self._info().max_length_game
    @property
    def position_dtype(self): return POSITION2 if self._info().position_bytes == 104 else POSITION3
