"""Times the training step (agpu_trainer_step) on synthetic batches: device ms from the library's CUDA events, end-to-end ms through
the Python shim with host batches, and the fp32 FLOP rate of the step (6 FLOP per weight per sample: forward, dX, dW).

    python scripts/train_bench.py [--game connect4|gobang9|hex7|reversi8] [--width 128] [--blocks 5] [--batch 8192] [--steps 20]
    torchrun --nproc-per-node 2 scripts/train_bench.py --dp        # data-parallel: checks against the single-GPU step as well
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alphagpu_b200 as ag  # noqa: E402

GAMES = {"connect4": ("connect4", 0, 0), "gobang9": ("gobang", 9, 5), "hex7": ("hex", 7, 0), "reversi8": ("reversi8", 0, 0), "ttt": ("gobang", 3, 3)}


def synth(spec, B, seed):
    rng = np.random.default_rng(seed)
    state = (rng.random((B, 2 * spec.VectorizedState)) < 0.3).astype(np.int8)
    pol = rng.random((B, spec.maxActions)).astype(np.float32)
    pol /= pol.sum(1, keepdims=True)
    return state, pol, rng.choice(np.array([0, 0.5, 1], np.float32), size=B), rng.integers(-1, 2, size=(B, spec.FeatureSize)).astype(np.int8)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--game", default="connect4"); ap.add_argument("--width", type=int, default=128); ap.add_argument("--blocks", type=int, default=5)
    ap.add_argument("--batch", type=int, default=8192); ap.add_argument("--steps", type=int, default=20); ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--dp", action="store_true")
    a = ap.parse_args()
    spec = ag.GameSpec.named(*GAMES[a.game])
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = int(os.environ.get("LOCAL_RANK", 0))
    if a.dp:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(dev)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{dev}"))
    net = ag.ressimplesf_full(2 * spec.VectorizedState, spec.maxActions, spec.FeatureSize, a.width, a.blocks, seed=0)
    per = a.batch // world
    tr = ag.Trainer.for_network(net, per, device=dev)
    batches = [synth(spec, a.batch, s) for s in range(4)]
    sl = ag.train.dp_slice(a.batch, rank, world)
    step = (lambda b: tr.step_dp(*[x[sl] for x in b])) if a.dp else (lambda b: tr.step(*b))
    for i in range(a.warmup):
        step(batches[i % 4])
    dev_ms, t0 = [], time.perf_counter()
    for i in range(a.steps):
        loss = step(batches[i % 4])
        dev_ms.append(sum(tr.last_ms()))
    wall = (time.perf_counter() - t0) / a.steps * 1e3
    nw = net.width * net.in_features + net.blocks * net.width ** 2 + (net.actions + 1 + net.fsize) * net.width
    flops = 6.0 * nw * per
    out = dict(game=a.game, net=f"{a.width}x{a.blocks + 1}", batch=a.batch, world=world, device_ms_per_step=float(np.median(dev_ms)),
               e2e_ms_per_step=wall, samples_per_s=a.batch / (wall * 1e-3), tflops_fp32_device=flops / (np.median(dev_ms) * 1e-3) / 1e12,
               loss=[float(x) for x in loss])
    if a.dp:
        # the data-parallel run must land where a single-GPU run over the full batches lands (summation order differs: tolerance)
        ref = ag.Trainer.for_network(net, a.batch, device=dev)
        for i in range(a.warmup):
            ref.step(*batches[i % 4])
        for i in range(a.steps):
            ref.step(*batches[i % 4])
        p, q = tr.get_params(), ref.get_params()
        out["dp_vs_single_max_abs_diff"] = float(max(np.abs(x - y).max() for x, y in zip(p.arrays(), q.arrays())))
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
