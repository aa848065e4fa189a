#!/bin/bash
# Round-2 run 11 on one B200: GG18 whole-signing test, then the full GPU suite.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_other_protocols.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -40 > $O/r02_t11a.log
tail -40 $O/r02_t11a.log
python -m pytest tests -m gpu -q --tb=short -x -p no:cacheprovider 2>&1 | tail -12 > $O/r02_t11.log
tail -6 $O/r02_t11.log
