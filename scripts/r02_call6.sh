#!/bin/bash
# round 2, GPU call 6: issue warp for the two-tile kernel, warp-aggregated item lists; racecheck on small cases
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_exact.py tests/test_gpu_nn.py tests/test_gpu_parity.py -x -q > gpurun_out/r02f_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02f_tests.log
tail -6 gpurun_out/r02f_tests.log
timeout 600 python scripts/quick_bench.py > gpurun_out/r02f_quick.txt 2>&1; head -4 gpurun_out/r02f_quick.txt
timeout 300 python scripts/fused_trace.py 32768 16384 4096 1024 > gpurun_out/r02f_trace.txt 2>&1; cat gpurun_out/r02f_trace.txt
AGPU_FUSED_MIN_GPC=136 timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis python scripts/sanitize_small.py 300 > gpurun_out/r02f_racecheck_two_tile.txt 2>&1; tail -4 gpurun_out/r02f_racecheck_two_tile.txt
timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis python scripts/sanitize_small.py 40 > gpurun_out/r02f_racecheck_small.txt 2>&1; tail -4 gpurun_out/r02f_racecheck_small.txt
