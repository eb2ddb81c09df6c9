"""alphagpu_b200 — B200-native self-play MCTS engine behind AlphaGPU's plugin surface.

The compute lives in libalphagpu.so (CUDA, sm_100a; C ABI in include/alphagpu.h); this package is the
Python host-side mirror of the reference's `mcts_gpu` module and game-plugin interface.
"""
from . import _lib  # noqa: F401
from .densenet import NetworkF, SNetwork2, convert_back, load_network, ressimplesf, ressimplesf_full, save_network  # noqa: F401
from .game import GameSpec  # noqa: F401
from .mcts_gpu import Context, MultiContext, PoolSample, duelnetwork, init, mcts, mcts_duel  # noqa: F401
from .train import Trainer, traininPipe  # noqa: F401
from .selfplay import elo_update, trainingPipeline  # noqa: F401
from .fast_mcts import MctsContext, move_dictionaries, testvsordi  # noqa: F401
