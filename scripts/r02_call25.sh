#!/bin/bash
# Connect4: paired configuration (two 256-thread CTAs per SM) at full load / also in the mid region
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_exact.py -x -q -k "paired or fused_ply_kernel_selfplay" > gpurun_out/r03a_tests.log 2>&1; tail -3 gpurun_out/r03a_tests.log
for pm in none 136 72; do
  if [ $pm != none ]; then export AGPU_FUSED_PAIR_MIN=$pm; else unset AGPU_FUSED_PAIR_MIN; fi
  echo "pair_min $pm"
  timeout 600 python scripts/ply_profile.py > gpurun_out/r03a_ply_profile_$pm.txt 2>&1; tail -3 gpurun_out/r03a_ply_profile_$pm.txt
done
paste gpurun_out/r03a_ply_profile_none.txt gpurun_out/r03a_ply_profile_136.txt gpurun_out/r03a_ply_profile_72.txt | cut -c1-120
