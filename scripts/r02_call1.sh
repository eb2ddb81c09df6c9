#!/bin/bash
# round 2, GPU call 1: full GPU test suite (incl. the exact-net parity of the per-ply kernel), smoke, bench, the never-run variants, traces
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r02a_gpu.txt 2>&1
python - > gpurun_out/r02a_devattr.txt 2>&1 <<'PY'
import torch
p = torch.cuda.get_device_properties(0)
print(p)
from cuda import cudart
for name in ("cudaDevAttrMaxPersistingL2CacheSize", "cudaDevAttrL2CacheSize", "cudaDevAttrMaxAccessPolicyWindowSize", "cudaDevAttrMaxSharedMemoryPerBlockOptin"):
    print(name, cudart.cudaDeviceGetAttribute(getattr(cudart.cudaDeviceAttr, name), 0))
PY
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_gpu_tests.log
tail -5 gpurun_out/r02a_gpu_tests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r02a_smoke.txt 2>&1; tail -2 gpurun_out/r02a_smoke.txt
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; tail -c 600 gpurun_out/r02a_bench.json
timeout 900 python scripts/lib_variant_experiment.py alphagpu_b200/libalphagpu_{prefetch,nhalf,resreg,tree16,tree48,ld128}.so > gpurun_out/r02a_variants.txt 2> gpurun_out/r02a_variants.err
cat gpurun_out/r02a_variants.txt | cut -c1-400
timeout 200 python scripts/fused_trace.py 32768 16384 4096 1024 > gpurun_out/r02a_trace_default.txt 2>&1
for v in tree48 resreg nhalf prefetch; do
  AGPU_LIB=$PWD/alphagpu_b200/libalphagpu_$v.so timeout 200 python scripts/fused_trace.py 32768 4096 1024 > gpurun_out/r02a_trace_$v.txt 2>&1
done
tail -n 3 gpurun_out/r02a_trace_*.txt
