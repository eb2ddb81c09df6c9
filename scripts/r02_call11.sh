#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python scripts/ply_profile.py alphagpu_b200/libalphagpu_c3.so > gpurun_out/r02k_ply_profile.txt 2>&1; tail -3 gpurun_out/r02k_ply_profile.txt
timeout 600 python -m pytest tests/test_gpu_exact.py tests/test_gpu_nn.py -x -q > gpurun_out/r02k_tests.log 2>&1; tail -3 gpurun_out/r02k_tests.log
AGPU_BENCH_DEBUG=1 timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err; tail -c 3000 gpurun_out/r02k_bench.json; tail -5 gpurun_out/r02k_bench.err
