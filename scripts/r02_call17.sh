#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
AGPU_LIB=$PWD/alphagpu_b200/libalphagpu_trace2.so timeout 900 python scripts/fused_trace.py --plies 20 16384 4096 1024 > gpurun_out/r02q_trace2_ply20.txt 2>&1; cat gpurun_out/r02q_trace2_ply20.txt
AGPU_LIB=$PWD/alphagpu_b200/libalphagpu_trace2.so timeout 900 python scripts/fused_trace.py 32768 > gpurun_out/r02q_trace2_ply0.txt 2>&1; cat gpurun_out/r02q_trace2_ply0.txt
