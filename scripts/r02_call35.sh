#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_exact.py -x -q -k "fused or duel" > gpurun_out/r03k_tests.log 2>&1; tail -3 gpurun_out/r03k_tests.log
timeout 900 python scripts/ply_profile.py > gpurun_out/r03k_ply_profile.txt 2>&1; tail -3 gpurun_out/r03k_ply_profile.txt
bash scripts/r02_ncu.sh r03k > gpurun_out/r03k_ncu.log 2>&1; tail -12 gpurun_out/r03k_ncu.log
