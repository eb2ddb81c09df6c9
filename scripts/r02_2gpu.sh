#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r03r_gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r03r_multi_tests_2gpu.log 2>&1; tail -4 gpurun_out/r03r_multi_tests_2gpu.log
AGPU_BENCH_DEBUG=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r03r_bench_2gpu.json 2> gpurun_out/r03r_bench_2gpu.err; tail -c 1500 gpurun_out/r03r_bench_2gpu.json; tail -3 gpurun_out/r03r_bench_2gpu.err
timeout 300 python - > gpurun_out/r03r_multi_timing.txt 2>&1 <<'PY'
import time, numpy as np, alphagpu_b200 as ag
spec = ag.GameSpec.named("connect4")
net = ag.ressimplesf(84, 7, 128, 6, seed=0)
for ngpus, games in ((1, 32768), (2, 65536), (2, 32768)):
    m = ag.MultiContext(spec, 64, games, 128, 6, ngpus)
    m.set_weights(net)
    m.selfplay(64, games, cpuct=1.5, seed=1, want_samples=False)
    t = time.perf_counter(); res, st, smp = m.selfplay(64, games, cpuct=1.5, seed=2); dt = time.perf_counter() - t
    print(f"agpu_multi_selfplay: {ngpus} GPU(s), {games} games: device_ms(max) {st['device_ms']:.2f}  wall incl. sample gather {1e3*dt:.1f} ms  sims/s (wall) {st['sims']/dt:.4g}  samples {len(smp['player'])}  (pageable numpy arrays)")
    pin = m.pinned_samples()
    m.selfplay(64, games, cpuct=1.5, seed=1, out=pin)
    t = time.perf_counter(); res, st, smp = m.selfplay(64, games, cpuct=1.5, seed=2, out=pin); dt = time.perf_counter() - t
    print(f"agpu_multi_selfplay: {ngpus} GPU(s), {games} games: device_ms(max) {st['device_ms']:.2f}  wall incl. sample gather {1e3*dt:.1f} ms  sims/s (wall) {st['sims']/dt:.4g}  samples {len(smp['player'])}  (page-locked arrays: MultiContext.pinned_samples)")
    m.close()
PY
cat gpurun_out/r03r_multi_timing.txt
