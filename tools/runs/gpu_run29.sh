#!/bin/bash
# Round-2 run 29 on one B200: HEAD validation — full GPU suite, smoke(), one short bench line.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q --tb=short -x -p no:cacheprovider 2>&1 | tail -5 > $O/r02_t29.log
tail -3 $O/r02_t29.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['parity']['units_match_cpu_twin'], d['modexp']['value'])"
