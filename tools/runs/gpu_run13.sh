#!/bin/bash
# Round-2 run 13 on one B200: the size-generic GG20 driver against its oracle.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_gg20_general.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -40 > $O/r02_t13.log
tail -40 $O/r02_t13.log
