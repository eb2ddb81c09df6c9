#!/bin/bash
# Connect4: fp32 residual of the ordinary epilogue in registers (default) vs in tensor memory (prev = HEAD before the change)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_exact.py -x -q -k "fused or duel" > gpurun_out/r02z_tests.log 2>&1; tail -3 gpurun_out/r02z_tests.log
timeout 600 python scripts/ply_profile.py alphagpu_b200/libalphagpu_prev.so > gpurun_out/r02z_ply_profile.txt 2>&1; cat gpurun_out/r02z_ply_profile.txt
