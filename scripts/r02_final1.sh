#!/bin/bash
# round 2, final validation on one GPU: whole GPU suite, smoke, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r03u}
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; tail -4 gpurun_out/${T}_tests.log
timeout 600 python __graft_entry__.py --smoke > gpurun_out/${T}_smoke.txt 2>&1; tail -2 gpurun_out/${T}_smoke.txt
timeout 1200 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['roofline']['frac'], d['roofline']['traffic'], d['cpu_baseline']['value'])
PY
