#!/bin/bash
# development: exact-net parity of the per-ply kernel + per-ply profile of the default build [against alphagpu_b200/libalphagpu_prev.so]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_exact.py -x -q -k "fused or duel" > gpurun_out/ab_tests.log 2>&1; tail -3 gpurun_out/ab_tests.log
P=""; [ -f alphagpu_b200/libalphagpu_prev.so ] && P=alphagpu_b200/libalphagpu_prev.so
timeout 900 python scripts/ply_profile.py $P > gpurun_out/ab_ply_profile.txt 2>&1; awk 'NR<6 || NR%4==0' gpurun_out/ab_ply_profile.txt; tail -2 gpurun_out/ab_ply_profile.txt
