#!/bin/bash
# Development: build the compile-time variants of libalphagpu.so (here, on CPU) and compare each with the default library on a GPU box.
#   bash scripts/variants_build_and_compare.sh build            # in the container (nvcc cross-compiles; the .so files travel with gpurun)
#   gpurun -- 'bash scripts/variants_build_and_compare.sh run'  # on the B200: identical output + ms per generation, per variant
# Variants: resreg = -DAG_SWAP_RESREG=1 (swapped epilogue keeps its residual values in registers),
#           nhalf = -DAG_NHALF=1 (two MMA chains per trunk layer, first-half epilogue under the second chain),
#           prefetch = -DAG_PREFETCH=1 (L2 prefetch of the children in the backup, of the record's tail in the descent),
#           tree16 / tree48 = -DAG_TREE_SMEM=16|48 (node cache of the small-batch kernel), ld128 = -DAG_DESC_LD128=1,
#           root = -DAG_ROOT_SMEM=1, l2 = -DAG_L2_HINT=1 (the last two measured in round 1: no gain).
set -e
cd "$(dirname "$0")/.."
VARIANTS="prefetch:-DAG_PREFETCH=1 nhalf:-DAG_NHALF=1 resreg:-DAG_SWAP_RESREG=1 tree16:-DAG_TREE_SMEM=16 tree48:-DAG_TREE_SMEM=48 ld128:-DAG_DESC_LD128=1"
case "${1:-build}" in
  build)
    for v in $VARIANTS; do AGPU_VARIANT=${v%%:*} AGPU_EXTRA_NVCC="${v#*:}" python -m alphagpu_b200.build; done ;;
  run)
    mkdir -p gpurun_out
    for v in $VARIANTS; do
      n=${v%%:*}
      timeout 200 python scripts/lib_variant_experiment.py alphagpu_b200/libalphagpu_$n.so > gpurun_out/variant_$n.txt 2> gpurun_out/variant_$n.err || true
      tail -4 gpurun_out/variant_$n.txt
    done ;;
esac
