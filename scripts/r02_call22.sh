#!/bin/bash
# large boards: network CTA capped at 64 registers (search blocks share its SM) + pipelined epilogue; variant r112 = same kernel, 112 registers
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_exact.py tests/test_gpu_nn.py -x -q > gpurun_out/r02x_tests.log 2>&1; tail -3 gpurun_out/r02x_tests.log
for v in default r112; do
  if [ $v != default ]; then export AGPU_LIB=$PWD/alphagpu_b200/libalphagpu_$v.so; else unset AGPU_LIB; fi
  timeout 600 python scripts/quick_bench.py --game hex --n 7 --games 16384 --width 512 --blocks 8 --reps 2 --profile 0 > gpurun_out/r02x_hex_$v.txt 2>&1; tail -1 gpurun_out/r02x_hex_$v.txt
  timeout 600 python scripts/quick_bench.py --game gobang --n 9 --nvict 5 --rollout 128 --games 16384 --width 512 --blocks 8 --reps 1 --profile 0 > gpurun_out/r02x_gobang_$v.txt 2>&1; tail -1 gpurun_out/r02x_gobang_$v.txt
  timeout 600 python scripts/quick_bench.py --game reversi8 --games 32768 --width 512 --blocks 8 --reps 1 --profile 0 > gpurun_out/r02x_reversi8_$v.txt 2>&1; tail -1 gpurun_out/r02x_reversi8_$v.txt
done
unset AGPU_LIB
timeout 600 python scripts/quick_bench.py --game hex --n 7 --games 16384 --width 512 --blocks 8 --reps 1 > gpurun_out/r02x_hex_profile.txt 2>&1; tail -10 gpurun_out/r02x_hex_profile.txt
