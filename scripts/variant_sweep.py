"""Development: fused per-ply kernel (all variants, chosen by games per CTA) against the stand-alone kernels, bit for bit, over game
counts that straddle the variant boundaries (64/65 and 128/129 games per CTA need L around 148 x those)."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

def run(fused, games, R, seed):
    code = f"""
import sys; sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})
import numpy as np, hashlib, alphagpu_b200 as ag
spec = ag.GameSpec.named("connect4")
net = ag.ressimplesf(84, 7, 128, 5, seed=0)
ctx = ag.Context(spec, {R}, {games}, 128, 5)
ctx.set_weights(net)
res, st, smp = ctx.selfplay({R}, {games}, cpuct=1.5, seed={seed})
h = hashlib.sha256()
for k in sorted(smp): h.update(np.ascontiguousarray(smp[k]).tobytes())
print(h.hexdigest(), list(map(int, res)), st["positions"])
"""
    env = dict(os.environ, AGPU_FUSED=str(fused))
    return subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True).stdout.strip()

bad = 0
for games, R in [(9, 5), (65, 7), (129, 6), (520, 9), (148 * 64, 4), (148 * 64 + 1, 4), (148 * 65, 3), (148 * 128, 3), (148 * 128 + 9, 3), (148 * 130, 3), (30000, 5)]:
    a, b = run(1, games, R, 5), run(0, games, R, 5)
    ok = a == b and len(a) > 0
    bad += not ok
    print(games, R, "OK" if ok else "MISMATCH", a[:16], b[:16], a.split("]")[-1], flush=True)
print("mismatches:", bad)
sys.exit(1 if bad else 0)
