#!/bin/bash
# round 2, GPU call 2: step-A changes (single-word headers, global copies only on the last rollout, root state + Philox ahead of time) and
# the K-split / tree-cache variants: exact-net parity per library, ms per generation, per-phase traces
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_exact.py tests/test_gpu_nn.py tests/test_gpu_parity.py -x -q > gpurun_out/r02b_tests_default.log 2>&1; echo "rc=$?" >> gpurun_out/r02b_tests_default.log
tail -3 gpurun_out/r02b_tests_default.log
for v in k4 tree48k4; do
  AGPU_LIB=$PWD/alphagpu_b200/libalphagpu_$v.so timeout 600 python -m pytest tests/test_gpu_exact.py -x -q -k "fused" > gpurun_out/r02b_tests_$v.log 2>&1; echo "rc=$?" >> gpurun_out/r02b_tests_$v.log
  tail -3 gpurun_out/r02b_tests_$v.log
done
timeout 900 python scripts/lib_variant_experiment.py alphagpu_b200/libalphagpu_{k2,k4,tree48,tree48k4}.so > gpurun_out/r02b_variants.txt 2> gpurun_out/r02b_variants.err
cut -c1-330 gpurun_out/r02b_variants.txt
timeout 200 python scripts/fused_trace.py 32768 16384 4096 1024 > gpurun_out/r02b_trace_default.txt 2>&1
for v in k4 tree48k4; do
  AGPU_LIB=$PWD/alphagpu_b200/libalphagpu_$v.so timeout 200 python scripts/fused_trace.py 4096 1024 > gpurun_out/r02b_trace_$v.txt 2>&1
done
cat gpurun_out/r02b_trace_*.txt
