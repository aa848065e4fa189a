#!/bin/bash
# Round-2 run 30 on two B200s: bench.py at N = 2 at HEAD (poller change, 4-blocks shapes).
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-modexp \
    > $O/r02_bench_n2_head.json 2> $O/r02_bench_n2_head.err
python -c "import json; d=json.loads(open('$O/r02_bench_n2_head.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['parity']['gather_consistent'], d['clocks'])"
tail -2 $O/r02_bench_n2_head.err
