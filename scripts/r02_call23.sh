#!/bin/bash
# can blocks of the search kernels share an SM with a network CTA?  carveout of the search kernels = max shared (same configuration as the network kernel)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in none 100 50; do
for v in default r112; do
  if [ $v != default ]; then export AGPU_LIB=$PWD/alphagpu_b200/libalphagpu_$v.so; else unset AGPU_LIB; fi
  if [ $c != none ]; then export AGPU_SEARCH_CARVEOUT=$c; else unset AGPU_SEARCH_CARVEOUT; fi
  echo "carveout $c lib $v"
  timeout 600 python scripts/quick_bench.py --game hex --n 7 --games 16384 --width 512 --blocks 8 --reps 2 --profile 0 > gpurun_out/r02y_hex_${v}_$c.txt 2>&1; tail -1 gpurun_out/r02y_hex_${v}_$c.txt
done
done
