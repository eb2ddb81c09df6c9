"""Development: phase cycles inside the per-ply kernel, per CTA and per warp (needs the trace build of the library:
    AGPU_VARIANT=trace AGPU_EXTRA_NVCC=-DAG_TRACE=1 python -m alphagpu_b200.build      # or -DAG_TRACE=2 for thread 0's fine-grained stamps
    AGPU_LIB=alphagpu_b200/libalphagpu_trace.so python scripts/fused_trace.py [--plies P] L [L ...]
--plies P: the searches start from positions reached by P random plies (games still running), instead of the empty board."""
import argparse, ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import alphagpu_b200 as ag
ap = argparse.ArgumentParser()
ap.add_argument("L", nargs="*", type=int, default=[32768, 8192, 1024, 128])
ap.add_argument("--plies", type=int, default=0)
a = ap.parse_args()
spec = ag.GameSpec.named("connect4")
net = ag.ressimplesf(84, 7, 128, 6, seed=0)
lib = ag._lib.load()
lib.agpu_debug_tc_trace.argtypes = [C.c_void_p]


def positions(ctx, L, plies):
    pos = ctx.Position(L)
    if plies == 0:
        return pos
    rng = np.random.default_rng(1)
    out = []
    while sum(len(o) for o in out) < L:
        cur = ctx.Position(L)
        for _ in range(plies):
            legal = ctx.canPlay(cur)
            act = np.array([rng.choice(np.nonzero(l)[0]) + 1 if l.any() else 1 for l in legal], np.int32)
            nxt = ctx.play(cur, act)
            over, _ = ctx.isOver(nxt)
            cur = nxt[~over]
            if len(cur) == 0:
                break
        if len(cur):
            out.append(cur)
    return np.concatenate(out)[:L]


for L in a.L:
    ctx = ag.Context(spec, 64, L, 128, 6, 0, 2)
    ctx.set_weights(net)
    pos = positions(ctx, L, a.plies)
    ctx.re_init(pos)
    ctx.mcts_single(64, cpuct=1.5, seed=1)
    buf = torch.zeros(128 * 512, dtype=torch.int64, device="cuda")
    lib.agpu_debug_tc_trace(C.c_void_p(buf.data_ptr()))
    ctx.re_init(pos)
    ctx.mcts_single(64, cpuct=1.5, seed=2, ply=a.plies)
    lib.agpu_debug_tc_trace(None)
    t = buf.cpu().numpy().reshape(-1, 128)
    t = t[t[:, 6] > 0]
    if len(t) == 0:
        print(f"L={L}: no trace recorded (is AGPU_LIB a -DAG_TRACE build?)")
        ctx.close()
        continue
    ph = t[:, :5].mean(0) / 64
    pw, dw = t[:, 32:48].mean(0) / 63, t[:, 48:64].mean(0) / 64
    print(f"L={L} (start: ply {a.plies}): CTAs {len(t)} games/CTA {t[:,5].mean():.0f}  cycles per rollout, thread 0 between barriers: search pool {ph[2]:.0f} descent {ph[3]:.0f} "
          f"network {ph[4]:.0f}  total {ph.sum():.0f} ({ph.sum()/1.965e3:.1f} us)  [a barrier blocks at the next consumer, not at issue: the split between "
          f"adjacent phases is approximate, the total and the per-warp times below are not]")
    print("      per-warp busy cycles in the search pool:", " ".join(f"{x:.0f}" for x in pw), f"(max {pw.max():.0f})")
    print("      per-warp cycles in the descent:", " ".join(f"{x:.0f}" for x in dw), f"(max {dw.max():.0f})")
    print(f"      => network phase ~ total - max pool - max descent = {ph.sum() - pw.max() - dw.max():.0f}")
    print("      descents that started while a pool unit was still running (must be 0):", int(t[:, 7].sum()))
    ly = t[:, 24:29].mean(0) / 64 / 6            # per trunk layer (6 per rollout for a 128x6 net)
    if ly.sum() > 0:
        print(f"      trunk layer (thread 0, AG_TRACE=2): wait weights {ly[0]:.0f}, issue MMAs + commits {ly[1]:.0f}, wait done {ly[2]:.0f}, epilogue {ly[3]:.0f}, barrier {ly[4]:.0f}  = {ly.sum():.0f} cycles")
    ob = t[:, 64:90].mean(0) / 64
    if ob.sum() > 0:
        print(f"      network phase seen by warp 8 (AG_TRACE=2): encode + first barrier {ob[0]:.0f}; per layer [wait for MMAs | epilogue | barrier]: "
              + " ".join(f"[{ob[2 + l]:.0f}|{ob[10 + l]:.0f}|{ob[18 + l]:.0f}]" for l in range(8)) + f"; head + end-of-rollout barrier {ob[1]:.0f}")
    x = t[:, 8:21].sum(0).astype(float)
    if x[2] > 0:
        print(f"      thread 0 (AG_TRACE=2): backup item load+update {x[0]/x[2]:.0f} cyc, solve {x[1]/x[2]:.0f} cyc ({x[2]/len(t)/64:.2f} items/rollout); "
              f"descent {x[4]/len(t)/64:.0f} cyc at warp-max depth, own depth {x[5]/len(t)/64:.2f}, first level {x[6]/len(t)/64:.0f} cyc; newton loop {x[3]/x[2]:.0f} cyc")
    ctx.close()
