#!/bin/bash
# round 2, GPU call 3: the rewritten per-ply kernel (work pool, level lists, alternating tiles, node cache): tests, ms per generation, traces
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_exact.py tests/test_gpu_nn.py tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q > gpurun_out/r02c_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02c_tests.log
tail -15 gpurun_out/r02c_tests.log
timeout 600 python scripts/quick_bench.py > gpurun_out/r02c_quick.txt 2>&1; tail -5 gpurun_out/r02c_quick.txt
timeout 300 python scripts/fused_trace.py 32768 16384 4096 1024 > gpurun_out/r02c_trace.txt 2>&1; cat gpurun_out/r02c_trace.txt
