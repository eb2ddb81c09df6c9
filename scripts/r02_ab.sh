#!/bin/bash
# development: exact-net parity of the per-ply kernel + per-ply profile of the default build against alphagpu_b200/libalphagpu_prev.so
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_exact.py tests/test_gpu_nn.py -x -q > gpurun_out/ab_tests.log 2>&1; tail -3 gpurun_out/ab_tests.log
timeout 900 python scripts/ply_profile.py alphagpu_b200/libalphagpu_prev.so > gpurun_out/ab_ply_profile.txt 2>&1; awk 'NR<10 || NR%4==0' gpurun_out/ab_ply_profile.txt; tail -2 gpurun_out/ab_ply_profile.txt
