#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
AGPU_LIB=$PWD/alphagpu_b200/libalphagpu_trace.so timeout 600 python scripts/fused_trace.py 32768 16384 4096 1024 > gpurun_out/r02p_trace_ply0.txt 2>&1; cat gpurun_out/r02p_trace_ply0.txt
AGPU_LIB=$PWD/alphagpu_b200/libalphagpu_trace.so timeout 900 python scripts/fused_trace.py --plies 20 16384 4096 1024 > gpurun_out/r02p_trace_ply20.txt 2>&1; cat gpurun_out/r02p_trace_ply20.txt
