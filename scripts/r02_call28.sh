#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python scripts/ply_profile.py alphagpu_b200/libalphagpu_prev.so alphagpu_b200/libalphagpu_v1.so alphagpu_b200/libalphagpu_v2.so > gpurun_out/r03d_ply_profile.txt 2>&1; cat gpurun_out/r03d_ply_profile.txt
