import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


# (game id, N, Nvict) — the five BASELINE.json configs + Reversi6
GAME_SPECS = {
    "connect4": (0, 0, 0),
    "ttt": (1, 3, 3),
    "gobang9": (1, 9, 5),
    "hex7": (2, 7, 0),
    "reversi8": (3, 0, 0),
    "reversi6": (4, 0, 0),
    "gobang5": (1, 5, 4),
    "hex5": (2, 5, 0),
}


@pytest.fixture(params=["connect4", "ttt", "gobang9", "hex7", "reversi8", "reversi6"])
def game_name(request):
    return request.param
