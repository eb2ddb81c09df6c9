"""Development: self-play generations over several uid bases and seeds (what ranks > 0 of a multi-GPU run play), each under a
watchdog: a kernel that does not come back within the limit is reported instead of hanging the caller."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faulthandler; faulthandler.enable()
import alphagpu_b200 as ag
spec = ag.GameSpec.named("connect4")
net = ag.ressimplesf(84, 7, 128, 5, seed=0)
games = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
ctx = ag.Context(spec, 64, games, 128, 5)
ctx.set_weights(net)
for uid_base in (0, 32768, 65536, 7 * 32768):
    for seed in (1000, 1001, 1002, 0, 1, 2, 77):
        t = time.perf_counter()
        done = threading.Event()
        def watchdog(u=uid_base, s=seed):
            if not done.wait(20):
                print(f"HANG uid_base {u} seed {s}", flush=True); faulthandler.dump_traceback(); os._exit(3)
        threading.Thread(target=watchdog, daemon=True).start()
        res, st, _ = ctx.selfplay(64, games, cpuct=1.5, seed=seed, uid_base=uid_base, want_samples=False)
        done.set()
        print(uid_base, seed, list(res), st["plies"], f"{time.perf_counter() - t:.3f}s", flush=True)
print("all ok")
