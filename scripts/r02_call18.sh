#!/bin/bash
# large boards: lanes per game 32 (default) vs 16 vs 8; parity of the variants (exact-net tests)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in w16 w8; do
  AGPU_LIB=$PWD/alphagpu_b200/libalphagpu_$v.so timeout 900 python -m pytest tests/test_gpu_exact.py -x -q -k "large_board" > gpurun_out/r02t_tests_$v.log 2>&1; tail -3 gpurun_out/r02t_tests_$v.log
done
for v in default w16 w8; do
  if [ $v != default ]; then export AGPU_LIB=$PWD/alphagpu_b200/libalphagpu_$v.so; else unset AGPU_LIB; fi
  timeout 600 python scripts/quick_bench.py --game hex --n 7 --games 16384 --width 512 --blocks 8 --reps 2 --profile 0 > gpurun_out/r02t_hex_$v.txt 2>&1; tail -1 gpurun_out/r02t_hex_$v.txt
  timeout 600 python scripts/quick_bench.py --game gobang --n 9 --nvict 5 --rollout 128 --games 16384 --width 512 --blocks 8 --reps 1 --profile 0 > gpurun_out/r02t_gobang_$v.txt 2>&1; tail -1 gpurun_out/r02t_gobang_$v.txt
  timeout 600 python scripts/quick_bench.py --game reversi8 --games 32768 --width 512 --blocks 8 --reps 1 --profile 0 > gpurun_out/r02t_reversi8_$v.txt 2>&1; tail -1 gpurun_out/r02t_reversi8_$v.txt
done
