"""Development: search-phase vs network-phase cycles inside the fused per-ply kernel, per CTA."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import alphagpu_b200 as ag
spec = ag.GameSpec.named("connect4")
net = ag.ressimplesf(84, 7, 128, 6, seed=0)
lib = ag._lib.load()
lib.agpu_debug_tc_trace.argtypes = [C.c_void_p]
for L in [int(x) for x in sys.argv[1:]] or [32768, 8192, 1024, 128]:
    ctx = ag.Context(spec, 64, L, 128, 6, 0, 2)
    ctx.set_weights(net)
    ctx.re_init(ctx.Position(L))
    ctx.mcts_single(64, cpuct=1.5, seed=1)
    buf = torch.zeros(64 * 512, dtype=torch.int64, device="cuda")
    lib.agpu_debug_tc_trace(C.c_void_p(buf.data_ptr()))
    ctx.re_init(ctx.Position(L))
    ctx.mcts_single(64, cpuct=1.5, seed=2)
    lib.agpu_debug_tc_trace(None)
    t = buf.cpu().numpy().reshape(-1, 64)
    t = t[t[:, 6] > 0]
    ph = t[:, :5].mean(0) / 64
    print(f"L={L}: CTAs {len(t)} games/CTA {t[:,5].mean():.0f}  cycles per rollout: search pool (backup + expand) {ph[2]:.0f} descent {ph[3]:.0f} "
          f"network {ph[4]:.0f}  total {ph.sum():.0f} ({ph.sum()/1.965e3:.1f} us)")
    ly = t[:, 24:29].mean(0) / 64 / 6            # per trunk layer (6 per rollout for a 128x6 net)
    print(f"      trunk layer (issuer thread): wait weights {ly[0]:.0f}, issue MMAs + commits {ly[1]:.0f}, wait done {ly[2]:.0f}, epilogue {ly[3]:.0f}, barrier {ly[4]:.0f}  = {ly.sum():.0f} cycles")
    x = t[:, 8:].sum(0).astype(float)
    if x[2] > 0:
        print(f"      thread 0: backup item load+update {x[0]/x[2]:.0f} cyc, solve {x[1]/x[2]:.0f} cyc ({x[2]/len(t)/64:.2f} items/rollout); "
              f"select total {x[4]/len(t)/64:.0f} cyc at warp-max depth, own depth {x[5]/len(t)/64:.2f}, first level {x[6]/len(t)/64:.0f} cyc; "
              f"newton loop {x[3]/x[2]:.0f} cyc; select level 0: loads {x[8]/len(t)/64:.0f}, +philox {x[9]/len(t)/64:.0f}, +scan {x[10]/len(t)/64:.0f}; philox alone {x[11]/len(t)/64:.0f}, entry->philox {x[12]/len(t)/64:.0f}")
    print("      descents that started while a pool unit was still running (must be 0):", int(t[:, 7].sum()))
    pw, dw = t[:, 32:48].mean(0) / 63, t[:, 48:64].mean(0) / 64
    print("      per-warp busy cycles in the search pool:", " ".join(f"{x:.0f}" for x in pw), f"(max {pw.max():.0f})")
    print("      per-warp cycles in the descent:", " ".join(f"{x:.0f}" for x in dw), f"(max {dw.max():.0f})")
    ctx.close()
