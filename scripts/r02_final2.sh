#!/bin/bash
# round 2, final evidence on one GPU: ncu captures (scripts/r02_ncu.sh), bench, reference arm, per-ply profile
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r03v}
bash scripts/r02_ncu.sh $T > gpurun_out/${T}_ncu.log 2>&1; tail -3 gpurun_out/${T}_ncu.log
timeout 1200 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 300 gpurun_out/${T}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; tail -c 300 gpurun_out/${T}_bench_reference.json
timeout 900 python scripts/ply_profile.py > gpurun_out/${T}_ply_profile.txt 2>&1; tail -3 gpurun_out/${T}_ply_profile.txt
