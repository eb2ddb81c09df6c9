#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_exact.py tests/test_gpu_configs.py -x -q > gpurun_out/r02o_tests.log 2>&1; tail -8 gpurun_out/r02o_tests.log
for cfg in "hex 7 0 16384 64" "gobang 9 5 16384 128" "reversi8 0 0 32768 64"; do set -- $cfg
  timeout 600 python scripts/quick_bench.py --game $1 --n $2 --nvict $3 --games $4 --rollout $5 --width 512 --blocks 8 --reps 2 > gpurun_out/r02o_quick_$1.txt 2>&1; grep -E "rep|select|nn " gpurun_out/r02o_quick_$1.txt | cut -c1-220
done
