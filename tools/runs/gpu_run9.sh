#!/bin/bash
# Round-2 evidence run 9 on one B200: ncu --set full of the dominant N-adic launch (template-qualified kernel filter), modexp timing
# after the persistent variant was reverted.  Outputs: gpurun_out/.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
TECDSA_SPLIT=0 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:nadic_jobs_kernel<.int.64' -s 8 -c 1 \
    -o $O/r02_nadic64 -f python tools/offline_throughput.py 4096 > $O/r02_ncu_nadic64.log 2>&1
tail -3 $O/r02_ncu_nadic64.log
ls -la $O/r02_nadic64.ncu-rep
