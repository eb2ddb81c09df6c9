"""One full generation (selfplay.jl trainingPipeline) on N GPUs, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/generation_dp.py --game gobang9 --samples 4096

self-play sharded by game uid (no collective) -> NCCL all-gather of the sample blocks -> data-parallel training with one gradient
all-reduce per step -> sharded duel with summed results -> Elo.  Prints one JSON line with the phase times and checks that every
rank ends the generation with the same parameters and the same buffer.
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alphagpu_b200 as ag  # noqa: E402

GAMES = {"connect4": ("connect4", 0, 0), "gobang9": ("gobang", 9, 5), "hex7": ("hex", 7, 0), "reversi8": ("reversi8", 0, 0), "ttt": ("gobang", 3, 3)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--game", default="gobang9"); ap.add_argument("--width", type=int, default=512); ap.add_argument("--blocks", type=int, default=7)
    ap.add_argument("--samples", type=int, default=4096); ap.add_argument("--rollout", type=int, default=128); ap.add_argument("--batchsize", type=int, default=2048)
    ap.add_argument("--duel-games", type=int, default=256); ap.add_argument("--duel-rollout", type=int, default=16)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, dev = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{dev}"))
    spec = ag.GameSpec.named(*GAMES[a.game])
    net = ag.ressimplesf_full(2 * spec.VectorizedState, spec.maxActions, spec.FeatureSize, a.width, a.blocks, seed=0)
    trainingnet = net.copy()
    buf = ag.PoolSample(spec, 2_000_000)
    times = {}
    import alphagpu_b200.selfplay as sp
    # time the three phases by wrapping the functions trainingPipeline calls
    def timed(name, fn):
        def w(*args, **kw):
            torch.cuda.synchronize(); t = time.perf_counter()
            r = fn(*args, **kw)
            torch.cuda.synchronize(); times[name] = times.get(name, 0.0) + time.perf_counter() - t
            return r
        return w
    sp.mcts, sp.traininPipe, sp.duelnetwork = timed("selfplay_s", sp.mcts), timed("train_s", sp.traininPipe), timed("duel_s", sp.duelnetwork)
    t0 = time.perf_counter()
    net2, tn2, passing, elo = ag.trainingPipeline(net, trainingnet, buf, 1, -1000.0, spec=spec, game=a.game, cpuct=1.5, samplesNumber=a.samples, rollout=a.rollout,
                                                  batchsize=a.batchsize, duel_games=a.duel_games, duel_rollout=a.duel_rollout, device=dev, verbose=False)
    total = time.perf_counter() - t0
    h = hashlib.sha256()
    for arr in tn2.arrays():
        h.update(np.ascontiguousarray(arr).tobytes())
    n = buf.length_buffer()
    h.update(buf.state[:n].tobytes()); h.update(buf.policy[:n].tobytes()); h.update(buf.value[:n].tobytes())
    digest = h.hexdigest()
    same = True
    if world > 1:
        all_d = [None] * world
        dist.all_gather_object(all_d, digest)
        same = len(set(all_d)) == 1
    if rank == 0:
        print(json.dumps(dict(game=a.game, net=f"{a.width}x{a.blocks + 1}", world=world, games=a.samples, rollout=a.rollout, samples_in_buffer=n,
                              batchsize=a.batchsize, **{k: round(v, 3) for k, v in times.items()}, total_s=round(total, 3), elo=elo, passing=passing,
                              trained=not np.array_equal(tn2.base, net.base), ranks_agree=same)))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    assert same


if __name__ == "__main__":
    main()
