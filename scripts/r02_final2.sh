#!/bin/bash
# round 2, final evidence on one GPU: ncu captures (scripts/r02_ncu.sh), traces, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash scripts/r02_ncu.sh r03p > gpurun_out/r03p_ncu.log 2>&1; tail -3 gpurun_out/r03p_ncu.log
timeout 1200 python bench.py > gpurun_out/r03p_bench.json 2> gpurun_out/r03p_bench.err; tail -c 600 gpurun_out/r03p_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r03p_bench_reference.json 2> gpurun_out/r03p_bench_reference.err; tail -c 600 gpurun_out/r03p_bench_reference.json
timeout 900 python scripts/ply_profile.py > gpurun_out/r03p_ply_profile.txt 2>&1; tail -3 gpurun_out/r03p_ply_profile.txt
