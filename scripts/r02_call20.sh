#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
AGPU_SEGMENTS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:^step_kernel -s 160 -c 1 -o gpurun_out/r02v_step_hex7_full -f python scripts/quick_bench.py --game hex --n 7 --games 16384 --rollout 64 --width 512 --blocks 8 --reps 1 --profile 0 > gpurun_out/r02v_ncu_step.out 2>&1
tail -3 gpurun_out/r02v_ncu_step.out
