"""Development: a library built with extra compile-time switches (python -m alphagpu_b200.build with AGPU_VARIANT / AGPU_EXTRA_NVCC)
against the default library: identical self-play output (digest over every sample array) and device time per generation.

    python scripts/lib_variant_experiment.py alphagpu_b200/libalphagpu_a.so [alphagpu_b200/libalphagpu_b.so ...] [--games 32768] [--reps 3]

AGPU_LIB is read when alphagpu_b200 is imported, so each library runs in its own process."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys, json, hashlib
sys.path.insert(0, %(root)r)
import numpy as np, alphagpu_b200 as ag
spec = ag.GameSpec.named("connect4")
out = {"digests": [], "ms": []}
for games, R in [(9, 5), (520, 9), (148 * 64 + 1, 4), (148 * 130, 3), (30000, 5)]:
    net = ag.ressimplesf(84, 7, 128, 5, seed=0)
    ctx = ag.Context(spec, R, games, 128, 5); ctx.set_weights(net)
    res, st, smp = ctx.selfplay(R, games, cpuct=1.5, seed=5)
    ctx.close()
    h = hashlib.sha256()
    for k in sorted(smp): h.update(np.ascontiguousarray(smp[k]).tobytes())
    out["digests"].append(h.hexdigest()[:16])
net = ag.ressimplesf(84, 7, 128, 6, seed=0)
ctx = ag.Context(spec, 64, %(games)d, 128, 6); ctx.set_weights(net)
for rep in range(%(reps)d + 1):
    res, st, _ = ctx.selfplay(64, %(games)d, cpuct=1.5, seed=rep, want_samples=False)
    if rep: out["ms"].append(round(st["device_ms"], 3))
ctx.profile(True); ctx.kernel_times(reset=True)
ctx.selfplay(64, %(games)d, cpuct=1.5, seed=0, want_samples=False)
kt = ctx.kernel_times(); ctx.close()
out["ply_fused_ms"] = round(kt["ply_fused"]["ms"], 3); out["sims"] = int(st["sims"])
print(json.dumps(out))
"""

ap = argparse.ArgumentParser()
ap.add_argument("lib", nargs="+")
ap.add_argument("--games", type=int, default=32768)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
res = {}
cases = [("default", None)] + [(os.path.basename(l).replace("libalphagpu_", "").replace(".so", ""), os.path.abspath(l)) for l in a.lib] + [("default_again", None)]
for name, lib in cases:
    env = dict(os.environ)
    env.pop("AGPU_LIB", None)
    if lib:
        env["AGPU_LIB"] = lib
    p = subprocess.run([sys.executable, "-c", CHILD % dict(root=ROOT, games=a.games, reps=a.reps)], env=env, capture_output=True, text=True)
    if p.returncode != 0:
        print(json.dumps(dict(case=name, error=p.stderr[-1500:])), flush=True)
        continue
    res[name] = json.loads(p.stdout.strip().splitlines()[-1])
    same = res[name]["digests"] == res["default"]["digests"] if "default" in res else None
    print(json.dumps(dict(case=name, lib=lib, identical_output=same, **res[name], best_ms=min(res[name]["ms"]))), flush=True)
ok = all(r["digests"] == res["default"]["digests"] for r in res.values()) if "default" in res else False
print(json.dumps(dict(all_identical=ok)))
sys.exit(0 if ok else 1)
