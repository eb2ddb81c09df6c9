#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:step_seg_kernel -s 160 -c 1 -o gpurun_out/r02n_step_hex7 -f python scripts/quick_bench.py --game hex --n 7 --games 16384 --rollout 64 --width 512 --blocks 8 --reps 1 --profile 0 > gpurun_out/r02n_ncu_step.out 2>&1; tail -3 gpurun_out/r02n_ncu_step.out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:tc_mlp512 -s 160 -c 1 -o gpurun_out/r02n_mlp512_hex7 -f python scripts/quick_bench.py --game hex --n 7 --games 16384 --rollout 64 --width 512 --blocks 8 --reps 1 --profile 0 > gpurun_out/r02n_ncu_mlp.out 2>&1; tail -3 gpurun_out/r02n_ncu_mlp.out
ls -la gpurun_out/*.ncu-rep
