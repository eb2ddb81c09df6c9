#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r03h
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^ply_kernel -s 3 -c 1 -o gpurun_out/${T}_ply_full -f python scripts/quick_bench.py --games 32768 --reps 1 --profile 0 > gpurun_out/${T}_ncu_full.out 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^ply_kernel -s 30 -c 1 -o gpurun_out/${T}_ply_tail -f python scripts/quick_bench.py --games 32768 --reps 1 --profile 0 > gpurun_out/${T}_ncu_tail.out 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^ply_kernel -s 18 -c 1 -o gpurun_out/${T}_ply_mid -f python scripts/quick_bench.py --games 32768 --reps 1 --profile 0 > gpurun_out/${T}_ncu_mid.out 2>&1
ls -la gpurun_out/${T}_ply*
