"""Development: wall time per ply of one self-play generation (AGPU_TRACE_PLIES), for one or more builds of the library side by side.

    python scripts/ply_profile.py [alphagpu_b200/libalphagpu_x.so ...] [--games 32768]"""
import argparse
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import sys
sys.path.insert(0, %(root)r)
import alphagpu_b200 as ag
spec = ag.GameSpec.named("connect4")
net = ag.ressimplesf(84, 7, 128, 6, seed=0)
ctx = ag.Context(spec, 64, %(games)d, 128, 6); ctx.set_weights(net)
ctx.selfplay(64, %(games)d, cpuct=1.5, seed=1, want_samples=False)
print("MARK", file=sys.stderr, flush=True)
res, st, _ = ctx.selfplay(64, %(games)d, cpuct=1.5, seed=0, want_samples=False)
print("device_ms", st["device_ms"], file=sys.stderr)
"""
ap = argparse.ArgumentParser()
ap.add_argument("lib", nargs="*")
ap.add_argument("--games", type=int, default=32768)
a = ap.parse_args()
cols = {}
for name, lib in [("default", None)] + [(os.path.basename(l).replace("libalphagpu_", "").replace(".so", ""), os.path.abspath(l)) for l in a.lib]:
    env = dict(os.environ, AGPU_TRACE_PLIES="1")
    env.pop("AGPU_LIB", None)
    if lib:
        env["AGPU_LIB"] = lib
    p = subprocess.run([sys.executable, "-c", CHILD % dict(root=ROOT, games=a.games)], env=env, capture_output=True, text=True)
    err = p.stderr.split("MARK")[-1]
    rows = [(int(m.group(1)), int(m.group(2)), float(m.group(3))) for m in re.finditer(r"ply (\d+) L (\d+) ms ([\d.]+)", err)]
    dm = re.search(r"device_ms ([\d.]+)", err)
    cols[name] = (rows, float(dm.group(1)) if dm else float("nan"))
names = list(cols)
print("ply      L  " + "  ".join(f"{n:>10s}" for n in names))
base = cols[names[0]][0]
for i, (ply, L, _) in enumerate(base):
    print(f"{ply:3d} {L:6d}  " + "  ".join(f"{cols[n][0][i][2]:10.3f}" if i < len(cols[n][0]) else " " * 10 for n in names))
print("sum        " + "  ".join(f"{sum(r[2] for r in cols[n][0]):10.3f}" for n in names))
print("device_ms  " + "  ".join(f"{cols[n][1]:10.3f}" for n in names))
