#!/bin/bash
# Round 1, call d: root record handed to the descent through shared memory (-DAG_ROOT_SMEM=1 development library).
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "start $(date +%s)" > gpurun_out/r01d_timeline.txt
timeout 150 python scripts/lib_variant_experiment.py alphagpu_b200/libalphagpu_root.so > gpurun_out/r01d_root_variant.txt 2> gpurun_out/r01d_root_variant.err; echo "rc=$?" >> gpurun_out/r01d_root_variant.txt
echo "variant done $(date +%s)" >> gpurun_out/r01d_timeline.txt
AGPU_LIB=$PWD/alphagpu_b200/libalphagpu_root.so timeout 300 python -m pytest tests -m gpu -q > gpurun_out/r01d_gpu_tests_root.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r01d_gpu_tests_root.log
echo "pytest done $(date +%s)" >> gpurun_out/r01d_timeline.txt
AGPU_LIB=$PWD/alphagpu_b200/libalphagpu_root.so timeout 60 python scripts/fused_trace.py 32768 4096 > gpurun_out/r01d_trace_root.txt 2>&1
cat gpurun_out/r01d_root_variant.txt; tail -3 gpurun_out/r01d_gpu_tests_root.log; cat gpurun_out/r01d_trace_root.txt
