#!/bin/bash
# Round-2 run 15 on one B200: full GPU suite, smoke(), then the two bench arms exactly as the driver launches them.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -15 > $O/r02_t15.log
tail -6 $O/r02_t15.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --impl reference > $O/r02_ref_n1.json 2> $O/r02_ref_n1.err
python bench.py > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err
tail -c 600 $O/r02_ref_n1.json; echo; head -c 700 $O/r02_bench_n1.json; echo
