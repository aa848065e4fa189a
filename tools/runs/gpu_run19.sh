#!/bin/bash
# Round-2 run 19 on one B200: Lindell-2017 key generation test, round-message documents, the tests that consume them.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
python tools/emit_wire_fixtures.py $O/wire 2>&1 | tail -3
python -m pytest tests/test_other_protocols.py tests/test_wire.py tests/test_gg20_general.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -25 > $O/r02_t19.log
tail -25 $O/r02_t19.log
