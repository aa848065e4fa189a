#!/bin/bash
# Round-2 run 16 on one B200: wire documents for the Rust pin test (incl. Lindell-2017), the tests that consume them, then compute-sanitizer
# over every kernel family including the section 8(f) rank-4 kernels.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
python tools/emit_wire_fixtures.py $O/wire 2>&1 | tail -2
python -m pytest tests/test_wire.py tests/test_other_protocols.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -12 > $O/r02_t16.log
tail -5 $O/r02_t16.log
: > $O/r02_san2_summary.log
for tool in memcheck racecheck synccheck; do
    timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_small.py > $O/r02_san2_$tool.log 2>&1
    echo "$tool rc=$?" >> $O/r02_san2_summary.log
    grep -h "ERROR SUMMARY\|RACECHECK SUMMARY" $O/r02_san2_$tool.log >> $O/r02_san2_summary.log
done
cat $O/r02_san2_summary.log
tail -12 $O/r02_san2_memcheck.log
