"""Development: per-layer clock64 trace of the tensor-core chain for a few CTAs."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import alphagpu_b200 as ag
spec = ag.GameSpec.named("connect4")
net = ag.ressimplesf(84, 7, 128, 6, seed=0)
L = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
ctx = ag.Context(spec, 8, L, 128, 6, 0, 2)
ctx.set_weights(net)
ctx.re_init(ctx.Position(L)); ctx.search_begin(); ctx.select(0, 1.5)
ncta = (L + 255) // 256
buf = torch.zeros(ncta * 2 * 16 * 4, dtype=torch.int64, device="cuda")
lib = ag._lib.load()
lib.agpu_debug_tc_trace.argtypes = [C.c_void_p]
for it in range(3):
    ctx.eval(fetch=False)
lib.agpu_debug_tc_trace(C.c_void_p(buf.data_ptr()))
import time
torch.cuda.synchronize(); t0=time.perf_counter(); ctx.eval(fetch=False); print('eval wall us', (time.perf_counter()-t0)*1e6)
lib.agpu_debug_tc_trace(None)
t = buf.cpu().numpy().reshape(ncta, 2, 16, 4)
for cta in (0, ncta // 2):
    for tile in (0, 1):
        base = t[cta, tile, 0, 0]
        print(f"cta {cta} tile {tile}: layer: start, issue_done, mma_done_seen, epilogue_done (cycles since layer-0 start); deltas")
        pr = t[cta, tile, 12:14].reshape(-1)
        print("   prologue: kernel_start->tmem_alloc_done %d, ->encode_done %d, layer0 start at %d; loop end->teardown: end_of_loop %d, after dealloc %d (since kernel start)" % (
            pr[1] - pr[0], pr[2] - pr[0], t[cta, tile, 0, 0] - pr[0], pr[3] - pr[0], pr[4] - pr[0]))
        for l in range(8):
            a = t[cta, tile, l] - base
            print(f"   l{l}: {a.tolist()}  issue {a[1]-a[0]}  mma_wait {a[2]-a[1]}  epilogue {a[3]-a[2] if a[3]>0 else None}")
