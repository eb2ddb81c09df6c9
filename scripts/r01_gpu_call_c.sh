#!/bin/bash
# Round 1, call c: L2 evict_last variant against the default library; ncu full captures of the two per-ply kernel variants at HEAD.
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "start $(date +%s)" > gpurun_out/r01c_timeline.txt
timeout 150 python scripts/lib_variant_experiment.py alphagpu_b200/libalphagpu_l2.so > gpurun_out/r01c_l2_variant.txt 2> gpurun_out/r01c_l2_variant.err; echo "rc=$?" >> gpurun_out/r01c_l2_variant.txt
echo "l2 done $(date +%s)" >> gpurun_out/r01c_timeline.txt
timeout 100 ncu --set full --clock-control none --import-source on -k regex:^ply_kernel -s 3 -c 1 -f -o gpurun_out/r01c_ply_full \
    python scripts/quick_bench.py --games 32768 --reps 1 --profile 0 > gpurun_out/r01c_ncu_full.out 2>&1
echo "ncu full done $(date +%s)" >> gpurun_out/r01c_timeline.txt
timeout 100 ncu --set full --clock-control none --import-source on -k regex:^ply_kernel -s 30 -c 1 -f -o gpurun_out/r01c_ply_tail \
    python scripts/quick_bench.py --games 32768 --reps 1 --profile 0 > gpurun_out/r01c_ncu_tail.out 2>&1
echo "ncu tail done $(date +%s)" >> gpurun_out/r01c_timeline.txt
AGPU_LIB=$PWD/alphagpu_b200/libalphagpu_l2.so timeout 100 ncu --set full --clock-control none -k regex:^ply_kernel -s 3 -c 1 -f -o gpurun_out/r01c_ply_full_l2 \
    python scripts/quick_bench.py --games 32768 --reps 1 --profile 0 > gpurun_out/r01c_ncu_full_l2.out 2>&1
echo "ncu l2 done $(date +%s)" >> gpurun_out/r01c_timeline.txt
timeout 90 python scripts/fused_trace.py 32768 16384 4096 1024 > gpurun_out/r01c_trace.txt 2>&1
echo "trace done $(date +%s)" >> gpurun_out/r01c_timeline.txt
cat gpurun_out/r01c_l2_variant.txt; cat gpurun_out/r01c_trace.txt | head -20
