#!/bin/bash
# large boards: lane-parallel backup + staged expand sums; block size of the search kernels 256 (default) / 128 / 64
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02w_tests.log 2>&1; tail -5 gpurun_out/r02w_tests.log
for v in default b128 b64; do
  if [ $v != default ]; then export AGPU_LIB=$PWD/alphagpu_b200/libalphagpu_$v.so; else unset AGPU_LIB; fi
  timeout 600 python scripts/quick_bench.py --game hex --n 7 --games 16384 --width 512 --blocks 8 --reps 2 --profile 0 > gpurun_out/r02w_hex_$v.txt 2>&1; tail -1 gpurun_out/r02w_hex_$v.txt
  timeout 600 python scripts/quick_bench.py --game gobang --n 9 --nvict 5 --rollout 128 --games 16384 --width 512 --blocks 8 --reps 1 --profile 0 > gpurun_out/r02w_gobang_$v.txt 2>&1; tail -1 gpurun_out/r02w_gobang_$v.txt
  timeout 600 python scripts/quick_bench.py --game reversi8 --games 32768 --width 512 --blocks 8 --reps 1 --profile 0 > gpurun_out/r02w_reversi8_$v.txt 2>&1; tail -1 gpurun_out/r02w_reversi8_$v.txt
done
unset AGPU_LIB
timeout 600 python scripts/quick_bench.py --game hex --n 7 --games 16384 --width 512 --blocks 8 --reps 1 > gpurun_out/r02w_hex_profile.txt 2>&1; tail -10 gpurun_out/r02w_hex_profile.txt
