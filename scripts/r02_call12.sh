#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r02l_multi_tests.log 2>&1; tail -15 gpurun_out/r02l_multi_tests.log
