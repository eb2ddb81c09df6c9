"""Host-side mirror of selfplay.jl: `trainingPipeline` — one generation = self-play with the current best net, training of
`trainingnet` on the sample pool, a 1024-game duel at 32 rollouts, the Elo update and the checkpoint (selfplay.jl:1-109).
Every GPU phase goes through the C ABI (alphagpu.h: self-play, duel; alphagpu_train.h: training)."""
from __future__ import annotations

import math
import os
from typing import Optional

from . import _lib
from .densenet import NetworkF, convert_back, save_network
from .game import GameSpec
from .mcts_gpu import PoolSample, duelnetwork, mcts
from .train import traininPipe


def elo_update(duel, currentelo: float, ngames: int = 1024) -> float:
    """EA = 1024/(v + 0.5 n); newelo = -400 log10(EA - 1) + currentelo (selfplay.jl:63-64), IEEE semantics at the edges
    (all wins -> +inf, no points -> -inf) as Julia's Float64 arithmetic gives them."""
    pts = duel[0] + 0.5 * duel[1]
    ea = math.inf if pts == 0 else ngames / pts
    x = ea - 1.0
    if x <= 0.0:
        return math.inf if x == 0.0 else math.nan
    return -400.0 * math.log10(x) + currentelo if math.isfinite(x) else -math.inf


def trainingPipeline(net: NetworkF, trainingnet: NetworkF, buffer: PoolSample, generation: int, currentelo: float = -1000.0, *, spec: GameSpec,
                     game: str = "", cpuct: float = 2.0, noise: float = 0.1, samplesNumber: int = 32000, rollout: int = 64, iteration: int = 100,
                     batchsize: int = 4096, lr: float = 0.001, epoch: int = 1, seed: int = 0, device: int = 0, nn_mode: int = _lib.NN_FP16_TC,
                     duel_games: int = 1024, duel_rollout: int = 32, save_dir: Optional[str] = None, verbose: bool = True):
    """trainingPipeline(net, trainingnet, buffer, generation, currentelo; ...) -> (net, trainingnet, passing, currentelo)
    (selfplay.jl:1-109).  `sizein/sizeout/fsize` of the reference are carried by `spec`."""
    passing = False
    i = generation
    if verbose:
        print(f"iteration: {i}")
    mcts(convert_back(net), rollout, samplesNumber, buffer, spec=spec, cpuct=cpuct, noise=noise, seed=seed + 7919 * i, device=device,
         nn_mode=nn_mode)
    if verbose:
        print("fin de la première volée")
        print("taille du buffer: ", buffer.length_buffer())
    trainingnet, report = traininPipe(batchsize, trainingnet, buffer, epoch=epoch, lr=lr, seed=seed + 104729 * i, device=device, verbose=verbose)
    index = (i - 1) % 1000 + 1
    duel = duelnetwork(convert_back(trainingnet), convert_back(net), duel_rollout, duel_games, spec=spec, seed=seed + 15485863 * i, device=device,
                       nn_mode=nn_mode)
    tot = max(1, sum(duel))
    if verbose:
        print("résultat du duel: ", [100.0 * d / tot for d in duel])
    newelo = elo_update(duel, currentelo, duel_games)
    if newelo > currentelo:
        currentelo = newelo
        passing = True
        net = trainingnet.copy()
    if save_dir is not None:
        os.makedirs(save_dir, exist_ok=True)
        save_network(os.path.join(save_dir, f"reseau{index}.agpu"), trainingnet, meta=dict(game=game, generation=i, elo=currentelo))
    return net, trainingnet, passing, currentelo
