#!/bin/bash
# Round-2 run 25 on one B200: the two bench arms as the driver launches them (after the poller fix and the 4-blocks-per-SM defaults).
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
python bench.py --impl reference > $O/r02_ref_n1.json 2> $O/r02_ref_n1.err
python bench.py > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err
head -c 300 $O/r02_bench_n1.json; echo; tail -c 300 $O/r02_ref_n1.json
