#!/bin/bash
# Round-2 run 20 on one B200: the 2-lanes-per-operand variant of the 2048-bit modexp kernel (254 registers, 8 warps per SM) against 4 lanes.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
: > $O/r02_tpi2.log
for t in 2 4 2 4; do echo "tpi=$t" >> $O/r02_tpi2.log; python tools/prof_modexp.py 2048 $t 65536 3 2>&1 | tail -2 >> $O/r02_tpi2.log; done
python -m pytest tests/test_modexp_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2 >> $O/r02_tpi2.log
cat $O/r02_tpi2.log
