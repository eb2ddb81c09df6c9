"""Development: a few small self-play generations through the per-ply kernel (for compute-sanitizer runs).
    [AGPU_FUSED_MIN_GPC=136] compute-sanitizer --tool racecheck python scripts/sanitize_small.py [games]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alphagpu_b200 as ag
G = int(sys.argv[1]) if len(sys.argv) > 1 else 300
for name, args, games, R in (("connect4", (), G, 6), ("gobang", (3, 3), G, 5)):
    spec = ag.GameSpec.named(name, *args)
    net = ag.ressimplesf(2 * spec.VectorizedState, spec.maxActions, 128, 5, seed=0)
    ctx = ag.Context(spec, R, games, 128, 5)
    ctx.set_weights(net)
    res, st, _ = ctx.selfplay(R, games, cpuct=1.5, seed=3, want_samples=True)
    print(name, games, list(map(int, res)), st["plies"], st["faults"], flush=True)
    ctx.close()
