#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_exact.py tests/test_gpu_nn.py -x -q > gpurun_out/r03c_tests.log 2>&1; tail -3 gpurun_out/r03c_tests.log
timeout 600 python scripts/ply_profile.py alphagpu_b200/libalphagpu_prev.so > gpurun_out/r03c_ply_profile.txt 2>&1; tail -4 gpurun_out/r03c_ply_profile.txt; head -8 gpurun_out/r03c_ply_profile.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^ply_kernel -s 3 -c 1 -o gpurun_out/r03c_ply_full -f python scripts/quick_bench.py --games 32768 --reps 1 --profile 0 > gpurun_out/r03c_ncu_full.out 2>&1; tail -2 gpurun_out/r03c_ncu_full.out
