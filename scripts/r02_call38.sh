#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_exact.py -x -q -k "fused or duel" > gpurun_out/r03m_tests.log 2>&1; tail -3 gpurun_out/r03m_tests.log
AGPU_LIB=$PWD/alphagpu_b200/libalphagpu_st2.so timeout 900 python -m pytest tests/test_gpu_exact.py -x -q -k "fused_ply_kernel_selfplay" > gpurun_out/r03m_tests_st2.log 2>&1; tail -3 gpurun_out/r03m_tests_st2.log
timeout 900 python scripts/ply_profile.py alphagpu_b200/libalphagpu_prev.so alphagpu_b200/libalphagpu_st2.so > gpurun_out/r03m_ply_profile.txt 2>&1; cat gpurun_out/r03m_ply_profile.txt
