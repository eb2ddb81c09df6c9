#!/bin/bash
# Connect4: register residual + select-free epilogue (default) vs HEAD~ (prev)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_exact.py -x -q -k "fused or duel or paired" > gpurun_out/r03b_tests.log 2>&1; tail -3 gpurun_out/r03b_tests.log
timeout 600 python scripts/ply_profile.py alphagpu_b200/libalphagpu_prev.so > gpurun_out/r03b_ply_profile.txt 2>&1; cat gpurun_out/r03b_ply_profile.txt
