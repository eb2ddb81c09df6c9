#!/bin/bash
# round 2, final validation on one GPU: whole GPU suite, smoke, bench, large boards, ncu evidence
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r03f
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; tail -4 gpurun_out/${T}_tests.log
timeout 600 python __graft_entry__.py --smoke > gpurun_out/${T}_smoke.txt 2>&1; tail -2 gpurun_out/${T}_smoke.txt
timeout 1200 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 1500 gpurun_out/${T}_bench.json
timeout 600 python scripts/quick_bench.py --game hex --n 7 --games 16384 --width 512 --blocks 8 --reps 2 > gpurun_out/${T}_hex.txt 2>&1; tail -11 gpurun_out/${T}_hex.txt
timeout 600 python scripts/quick_bench.py --game gobang --n 9 --nvict 5 --rollout 128 --games 16384 --width 512 --blocks 8 --reps 2 > gpurun_out/${T}_gobang.txt 2>&1; tail -11 gpurun_out/${T}_gobang.txt
timeout 600 python scripts/quick_bench.py --game reversi8 --games 32768 --width 512 --blocks 8 --reps 2 > gpurun_out/${T}_reversi8.txt 2>&1; tail -11 gpurun_out/${T}_reversi8.txt
timeout 600 python scripts/quick_bench.py --games 32768 --reps 2 > gpurun_out/${T}_quick.txt 2>&1; tail -9 gpurun_out/${T}_quick.txt
