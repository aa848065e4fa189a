#!/bin/bash
# Round-2 run 18 on one B200: Lindell-2017 key generation test, Lindell-2017 signing throughput.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_other_protocols.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -25 > $O/r02_t18.log
tail -25 $O/r02_t18.log
python tools/l17_throughput.py 16384 2>&1 | tail -2 > $O/r02_l17_throughput.json
cat $O/r02_l17_throughput.json
