"""Development: the DUAL variant of the per-ply kernel (one tile, 256 threads x 128 registers, two CTAs per SM; fused.cuh FCfg::DUAL)
against the stand-alone kernels (bit for bit) and against the default variant selection (time), in one process.

    python scripts/dual_experiment.py [--skip-parity] [--games 32768] [--reps 3]

The environment switches are read when a context is created, so each case builds its own context."""
import argparse
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import alphagpu_b200 as ag

KEYS = ("AGPU_FUSED", "AGPU_FUSED_DUAL", "AGPU_FUSED_DUAL_MIN", "AGPU_FUSED_DUAL_STAGGER_US")


def context(env, games, R, blocks=6):
    for k in KEYS:
        os.environ.pop(k, None)
    os.environ.update(env)
    spec = ag.GameSpec.named("connect4")
    net = ag.ressimplesf(84, 7, 128, blocks, seed=0)
    ctx = ag.Context(spec, R, games, 128, blocks)
    ctx.set_weights(net)
    return ctx


def digest(env, games, R, seed):
    ctx = context(env, games, R, blocks=5)
    res, st, smp = ctx.selfplay(R, games, cpuct=1.5, seed=seed)
    ctx.close()
    h = hashlib.sha256()
    for k in sorted(smp):
        h.update(np.ascontiguousarray(smp[k]).tobytes())
    return h.hexdigest()[:16], [int(x) for x in res], int(st["positions"])


ap = argparse.ArgumentParser()
ap.add_argument("--skip-parity", action="store_true")
ap.add_argument("--games", type=int, default=32768)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()

bad = 0
if not a.skip_parity:
    # DUAL_MIN=1 forces the dual kernel for every launch (any number of games per CTA); the default threshold is exercised by the
    # large cases (more than 128 games per SM)
    for games, R in [(1, 1), (9, 5), (129, 6), (520, 9), (148 * 64 + 1, 4), (148 * 128 + 9, 3), (148 * 130, 3), (30000, 5), (40000, 3)]:
        ref = digest({"AGPU_FUSED": "0"}, games, R, 5)
        forced = digest({"AGPU_FUSED_DUAL": "1", "AGPU_FUSED_DUAL_MIN": "1"}, games, R, 5)
        dflt = digest({"AGPU_FUSED_DUAL": "1"}, games, R, 5)
        ok = ref == forced == dflt
        bad += not ok
        print(json.dumps(dict(parity=dict(games=games, R=R, ok=ok, ref=ref[0], forced=forced[0], dual=dflt[0], results=ref[1]))), flush=True)
    print(json.dumps(dict(parity_mismatches=bad)), flush=True)

cases = [("default", {}), ("dual_min129", {"AGPU_FUSED_DUAL": "1"}),
         ("dual_min129_stagger12us", {"AGPU_FUSED_DUAL": "1", "AGPU_FUSED_DUAL_STAGGER_US": "12"}),
         ("dual_min65", {"AGPU_FUSED_DUAL": "1", "AGPU_FUSED_DUAL_MIN": "65"}),
         ("dual_min33", {"AGPU_FUSED_DUAL": "1", "AGPU_FUSED_DUAL_MIN": "33"}), ("default_again", {})]
for name, env in cases:
    ctx = context(env, a.games, 64)
    ms = []
    for rep in range(a.reps + 1):
        res, st, _ = ctx.selfplay(64, a.games, cpuct=1.5, seed=rep, want_samples=False)
        if rep:
            ms.append(round(st["device_ms"], 3))
    ctx.profile(True)
    ctx.kernel_times(reset=True)
    ctx.selfplay(64, a.games, cpuct=1.5, seed=0, want_samples=False)
    kt = ctx.kernel_times()
    ctx.close()
    print(json.dumps(dict(case=name, env=env, device_ms=ms, best_ms=min(ms), sims=int(st["sims"]), sims_per_s=round(st["sims"] / (min(ms) * 1e-3)),
                          ply_fused_ms=round(kt["ply_fused"]["ms"], 3), launches=kt["ply_fused"]["launches"])), flush=True)
sys.exit(1 if bad else 0)
