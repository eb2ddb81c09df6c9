#!/bin/bash
# AGPU_STREAM_SAMPLES: rows of a ply copied to the caller's pinned buffers while the next ply searches
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
AGPU_STREAM_SAMPLES=1 timeout 900 python -m pytest tests/test_gpu_exact.py tests/test_gpu_parity.py -x -q -k "selfplay or config" > gpurun_out/r03g_tests_stream.log 2>&1; tail -3 gpurun_out/r03g_tests_stream.log
for s in 0 1 0 1; do
  AGPU_STREAM_SAMPLES=$s timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stream $s', 'value', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), 'ms', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2))"
done
