#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r03o_tests.log 2>&1; tail -4 gpurun_out/r03o_tests.log
timeout 900 python scripts/ply_profile.py alphagpu_b200/libalphagpu_prev.so > gpurun_out/r03o_ply_profile.txt 2>&1; head -8 gpurun_out/r03o_ply_profile.txt; tail -3 gpurun_out/r03o_ply_profile.txt
