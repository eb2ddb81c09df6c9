#!/bin/bash
# round 2: the bench contract at 8 and 4 GPUs of one box (torchrun, one rank per GPU) + the single-process multi-device call
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r03w
nvidia-smi -L > gpurun_out/${T}_gpus.txt
for n in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/${T}_bench_${n}gpu.json 2> gpurun_out/${T}_bench_${n}gpu.err; tail -c 400 gpurun_out/${T}_bench_${n}gpu.json; echo
done
timeout 600 python - > gpurun_out/${T}_multi_timing.txt 2>&1 <<'PY'
import time, numpy as np, alphagpu_b200 as ag
spec = ag.GameSpec.named("connect4")
net = ag.ressimplesf(84, 7, 128, 6, seed=0)
for ngpus, games in ((8, 262144), (8, 32768)):
    m = ag.MultiContext(spec, 64, games, 128, 6, ngpus)
    m.set_weights(net)
    pin = m.pinned_samples()
    m.selfplay(64, games, cpuct=1.5, seed=1, out=pin)
    t = time.perf_counter(); res, st, smp = m.selfplay(64, games, cpuct=1.5, seed=2, out=pin); dt = time.perf_counter() - t
    print(f"agpu_multi_selfplay: {ngpus} GPU(s), {games} games: device_ms(max) {st['device_ms']:.2f}  wall incl. sample gather {1e3*dt:.1f} ms  sims/s (wall) {st['sims']/dt:.4g}  samples {len(smp['player'])}  (page-locked arrays)")
    m.close()
PY
cat gpurun_out/${T}_multi_timing.txt
