#!/bin/bash
# repeatability of the device-resident loop of bench.py
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out
for i in 1 2 3; do python bench.py --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])"; done | tee $O/r02_bench_repeat.log
