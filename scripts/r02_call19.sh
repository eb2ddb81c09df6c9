#!/bin/bash
# large boards after the shared-memory staging of the ordered sums: full GPU suite + the three configs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02u_tests.log 2>&1; tail -5 gpurun_out/r02u_tests.log
timeout 600 python scripts/quick_bench.py --game hex --n 7 --games 16384 --width 512 --blocks 8 --reps 2 > gpurun_out/r02u_hex.txt 2>&1; tail -12 gpurun_out/r02u_hex.txt
timeout 600 python scripts/quick_bench.py --game gobang --n 9 --nvict 5 --rollout 128 --games 16384 --width 512 --blocks 8 --reps 1 > gpurun_out/r02u_gobang.txt 2>&1; tail -11 gpurun_out/r02u_gobang.txt
timeout 600 python scripts/quick_bench.py --game reversi8 --games 32768 --width 512 --blocks 8 --reps 1 > gpurun_out/r02u_reversi8.txt 2>&1; tail -11 gpurun_out/r02u_reversi8.txt
