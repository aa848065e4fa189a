#!/bin/bash
# Round-2 run 28 on one B200: GG18 key generation + signing test and the rest of the other-protocol tests.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_other_protocols.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -25 > $O/r02_t28.log
tail -25 $O/r02_t28.log
