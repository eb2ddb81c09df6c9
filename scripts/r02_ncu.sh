#!/bin/bash
# round 2: ncu evidence for profiles/ — launch list of the bench command (gpu__time_duration), full captures of the per-ply kernel at full
# load (ply 3, 32768 games), in the middle (ply 18) and in the tail (ply 30), and of the two large-board kernels (Hex 7, 512x8)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r03i}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 640 -c 260 --csv --log-file gpurun_out/${T}_ncu_launch_list.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${T}_ncu_launch_list.out 2>&1
for p in full:3 mid:18 tail:30; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^ply_kernel -s ${p#*:} -c 1 -o gpurun_out/${T}_ply_${p%:*} -f python scripts/quick_bench.py --games 32768 --reps 1 --profile 0 > gpurun_out/${T}_ncu_${p%:*}.out 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_seg_kernel -s 160 -c 1 -o gpurun_out/${T}_step_hex7 -f python scripts/quick_bench.py --game hex --n 7 --games 16384 --rollout 64 --width 512 --blocks 8 --reps 1 --profile 0 > gpurun_out/${T}_ncu_step.out 2>&1
AGPU_SEGMENTS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:^step_kernel -s 160 -c 1 -o gpurun_out/${T}_step_hex7_16384 -f python scripts/quick_bench.py --game hex --n 7 --games 16384 --rollout 64 --width 512 --blocks 8 --reps 1 --profile 0 > gpurun_out/${T}_ncu_step_full.out 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_mlp512 -s 160 -c 1 -o gpurun_out/${T}_mlp512_hex7 -f python scripts/quick_bench.py --game hex --n 7 --games 16384 --rollout 64 --width 512 --blocks 8 --reps 1 --profile 0 > gpurun_out/${T}_ncu_mlp.out 2>&1
ls -la gpurun_out/${T}_*
