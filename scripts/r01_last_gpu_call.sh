#!/bin/bash
# The last GPU call of round 1 (11 GPU-minutes left): most valuable first, every step under its own timeout.
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "start $(date +%s)" > gpurun_out/r01b_timeline.txt
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/r01b_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r01b_gpu_tests.log
echo "pytest done $(date +%s)" >> gpurun_out/r01b_timeline.txt
AGPU_DEBUG=1 timeout 240 python scripts/dual_experiment.py > gpurun_out/r01b_dual_experiment.txt 2> gpurun_out/r01b_dual_experiment.err; echo "rc=$?" >> gpurun_out/r01b_dual_experiment.txt
echo "dual done $(date +%s)" >> gpurun_out/r01b_timeline.txt
timeout 200 python bench.py --steps 5 --warmup 3 > gpurun_out/r01b_bench.json 2> gpurun_out/r01b_bench.err
echo "bench done $(date +%s)" >> gpurun_out/r01b_timeline.txt
AGPU_FUSED_DUAL=1 timeout 90 python scripts/fused_trace.py 32768 > gpurun_out/r01b_trace_dual.txt 2>&1
echo "trace done $(date +%s)" >> gpurun_out/r01b_timeline.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 640 -c 230 --csv --log-file gpurun_out/r01b_ncu_launch_list.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r01b_ncu_launch_list.out 2>&1
echo "ncu list done $(date +%s)" >> gpurun_out/r01b_timeline.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:ply_kernel -s 3 -c 1 -f -o gpurun_out/r01b_ply_full \
    python scripts/quick_bench.py --games 32768 --reps 1 --profile 0 > gpurun_out/r01b_ncu_full.out 2>&1
echo "ncu full done $(date +%s)" >> gpurun_out/r01b_timeline.txt
tail -3 gpurun_out/r01b_gpu_tests.log; cat gpurun_out/r01b_dual_experiment.txt | tail -12; head -c 600 gpurun_out/r01b_bench.json
