#!/bin/bash
# Round 1, call e: the default library at HEAD — smoke, whole GPU suite, bench line.
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r01e_smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/r01e_smoke.txt
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r01e_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r01e_gpu_tests.log
timeout 200 python bench.py --steps 5 --warmup 3 > gpurun_out/r01e_bench.json 2> gpurun_out/r01e_bench.err
tail -2 gpurun_out/r01e_smoke.txt; tail -3 gpurun_out/r01e_gpu_tests.log; head -c 400 gpurun_out/r01e_bench.json
