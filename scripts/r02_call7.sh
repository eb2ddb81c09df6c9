#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python scripts/ply_profile.py alphagpu_b200/libalphagpu_c3.so > gpurun_out/r02j_ply_profile.txt 2>&1; cat gpurun_out/r02j_ply_profile.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02j_tests.log 2>&1; tail -3 gpurun_out/r02j_tests.log
