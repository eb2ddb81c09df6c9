"""Command line of the reference's main*.jl (main4IARow.jl:86-143 and its siblings): generations of self-play -> training -> duel.

    python -m alphagpu_b200.main --game 4IARow --samples 32768 --rollout 64 --generation 100 --batchsize 8192 --cpuct 1.5 --noise 0.2857

Options and defaults are the reference's; `--game` replaces the choice of main file (4IARow, Gobang, Hex, Reversi6x6, Reversi8x8)
and `--width/--blocks` the hard-coded `ressimplesf(..., 512, 4)`.
"""
from __future__ import annotations

import argparse

from .densenet import ressimplesf_full
from .game import GameSpec
from .mcts_gpu import PoolSample
from .selfplay import trainingPipeline

# main file -> (plugin name, N, Nvict, n_filter, n_tower) as each main*.jl builds its net
GAMES = {"4IARow": ("connect4", 0, 0, 512, 4), "Gobang": ("gobang", 9, 5, 512, 6), "Hex": ("hex", 7, 0, 512, 8),
         "Reversi6x6": ("reversi6", 0, 0, 512, 4), "Reversi8x8": ("reversi8", 0, 0, 512, 8)}


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(prog="alphagpu_b200.main", description=__doc__.split("\n")[0])
    ap.add_argument("--game", default="4IARow", choices=sorted(GAMES))
    ap.add_argument("--samples", type=int, default=32 * 1024, help="number of selfplay games per generation")
    ap.add_argument("--rollout", type=int, default=64, help="number of rollouts")
    ap.add_argument("--generation", type=int, default=100, help="number of generations")
    ap.add_argument("--batchsize", type=int, default=2 * 4096, help="batchsize for training")
    ap.add_argument("--cpuct", type=float, default=1.5, help="cpuct (exploration coefficient in cpuct formula)")
    ap.add_argument("--noise", type=float, default=None, help="uniform noise at the root, default to 2/maxActions")
    ap.add_argument("--width", type=int, default=None, help="n_filter of ressimplesf (default: the main file's)")
    ap.add_argument("--blocks", type=int, default=None, help="n_tower of ressimplesf (default: the main file's)")
    ap.add_argument("--buffer", type=int, default=2_000_000, help="PoolSample capacity (main4IARow.jl:125)")
    ap.add_argument("--board", type=int, default=None, help="board size N for Gobang / Hex")
    ap.add_argument("--nvict", type=int, default=None, help="stones in a row to win (Gobang)")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--save-dir", default=None, help="write reseau<index>.agpu checkpoints here (selfplay.jl:86-98)")
    return ap


def main(argv=None):
    args = build_parser().parse_args(argv)
    name, N, nv, width, blocks = GAMES[args.game]
    spec = GameSpec.named(name, args.board if args.board is not None else N, args.nvict if args.nvict is not None else nv)
    width = args.width or width
    blocks = args.blocks if args.blocks is not None else blocks
    noise = args.noise if args.noise is not None else 2.0 / spec.maxActions
    net = ressimplesf_full(2 * spec.VectorizedState, spec.maxActions, spec.FeatureSize, width, blocks, seed=args.seed)
    trainingnet = net.copy()
    buffer = PoolSample(spec, args.buffer)
    best, currentelo = 1, -1000.0
    for i in range(1, args.generation + 1):
        net, trainingnet, passing, currentelo = trainingPipeline(
            net, trainingnet, buffer, i, currentelo, spec=spec, game=args.game, cpuct=args.cpuct, noise=noise, samplesNumber=args.samples,
            rollout=args.rollout, iteration=1, batchsize=args.batchsize, seed=args.seed, device=args.device, save_dir=args.save_dir)
        if passing:
            best = i
        print(f"meilleur réseau: {best}")
        print(f"elo actuel: {currentelo}, generation: {i}")
    return net, trainingnet, currentelo


if __name__ == "__main__":
    main()
