#!/bin/bash
# round 2, final build: compute-sanitizer (memcheck, racecheck) on small generations through every shape of the per-ply kernel and the
# large-board kernels, and the per-phase clock traces of the -DAG_TRACE=2 build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r03t
{
for gpc in 8 40 72 136; do
  echo "== memcheck, AGPU_FUSED_MIN_GPC=$gpc"; AGPU_FUSED_MIN_GPC=$gpc timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python scripts/sanitize_small.py 300 2>&1 | tail -4
  echo "== racecheck, AGPU_FUSED_MIN_GPC=$gpc"; AGPU_FUSED_MIN_GPC=$gpc timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python scripts/sanitize_small.py 300 2>&1 | tail -4
done
echo "== memcheck, large boards (hex 7 512x8, reversi8 512x8; stand-alone kernels)"
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python - <<'PY' 2>&1 | tail -5
import sys; sys.path.insert(0, ".")
import alphagpu_b200 as ag
for name, args in (("hex", (7,)), ("reversi8", ())):
    spec = ag.GameSpec.named(name, *args)
    net = ag.ressimplesf(2 * spec.VectorizedState, spec.maxActions, 512, 2, seed=0)
    ctx = ag.Context(spec, 8, 64, 512, 2); ctx.set_weights(net)
    res, st, _ = ctx.selfplay(8, 64, cpuct=1.5, seed=3)
    print(name, list(map(int, res)), st["plies"], st["faults"], flush=True); ctx.close()
PY
} > gpurun_out/${T}_sanitizer.txt 2>&1
cat gpurun_out/${T}_sanitizer.txt | grep -v "^$" | tail -40
AGPU_LIB=$PWD/alphagpu_b200/libalphagpu_trace.so timeout 900 python scripts/fused_trace.py 32768 > gpurun_out/${T}_trace_ply0.txt 2>&1; cat gpurun_out/${T}_trace_ply0.txt
AGPU_LIB=$PWD/alphagpu_b200/libalphagpu_trace.so timeout 900 python scripts/fused_trace.py --plies 20 16384 4096 1024 > gpurun_out/${T}_trace_ply20.txt 2>&1; cat gpurun_out/${T}_trace_ply20.txt
