#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_exact.py -x -q -k "fused or duel" > gpurun_out/r03l_tests.log 2>&1; tail -3 gpurun_out/r03l_tests.log
timeout 900 python scripts/ply_profile.py alphagpu_b200/libalphagpu_prev.so > gpurun_out/r03l_ply_profile.txt 2>&1; cat gpurun_out/r03l_ply_profile.txt
