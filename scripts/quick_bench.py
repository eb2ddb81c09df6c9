"""Quick device-side timing of one self-play generation (development aid, not the bench contract)."""
import argparse
import json
import sys
import time
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import alphagpu_b200 as ag

ap = argparse.ArgumentParser()
ap.add_argument("--game", default="connect4")
ap.add_argument("--n", type=int, default=0)
ap.add_argument("--nvict", type=int, default=0)
ap.add_argument("--games", type=int, default=32768)
ap.add_argument("--rollout", type=int, default=64)
ap.add_argument("--width", type=int, default=128)
ap.add_argument("--blocks", type=int, default=6)
ap.add_argument("--nn-mode", type=int, default=2)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--profile", type=int, default=1)
a = ap.parse_args()

spec = ag.GameSpec.named(a.game, a.n, a.nvict)
net = ag.ressimplesf(2 * spec.VectorizedState, spec.maxActions, a.width, a.blocks, seed=0)
ctx = ag.Context(spec, a.rollout, a.games, a.width, a.blocks, 0, a.nn_mode)
ctx.set_weights(net)
print("layout", ctx.layout())
for rep in range(a.reps):
    t = time.time()
    res, st, _ = ctx.selfplay(a.rollout, a.games, cpuct=1.5, seed=rep, want_samples=False)
    dt = time.time() - t
    print(json.dumps(dict(rep=rep, wall_s=round(dt, 4), device_ms=round(st["device_ms"], 3), sims=st["sims"], plies=st["plies"],
                          sims_per_s=round(st["sims"] / (st["device_ms"] / 1e3)), results=res.tolist(), launches=st["kernel_launches"],
                          mean_len=st["positions"] / a.games)))
if a.profile:
    ctx.profile(True)
    ctx.kernel_times(reset=True)
    res, st, _ = ctx.selfplay(a.rollout, a.games, cpuct=1.5, seed=0, want_samples=False)
    kt = ctx.kernel_times()
    tot = sum(v["ms"] for k, v in kt.items() if isinstance(v, dict))
    for k, v in kt.items():
        if isinstance(v, dict) and v["launches"]:
            print(f"  {k:14s} launches {v['launches']:6d}  ms {v['ms']:10.3f}  share {v['ms'] / tot:6.1%}  us/launch {1e3 * v['ms'] / v['launches']:8.2f}")
    print("  d_bar", kt["nodes_traversed"] / max(1, kt["descents"]), "device_ms(profiled)", st["device_ms"])
