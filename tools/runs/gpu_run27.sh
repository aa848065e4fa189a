#!/bin/bash
# Round-2 run 27 on one B200: ncu --set full of the largest launch of the second and third kernels of the offline stage
# (p-adic job list, 2048-bit job list) in a 4096-session batch on one stream.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
TECDSA_SPLIT=0 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:nadic_jobs_kernel<.int.32' -s 4 -c 1 \
    -o $O/r02_nadic32 -f python tools/offline_throughput.py 4096 > $O/r02_ncu_nadic32.log 2>&1
TECDSA_SPLIT=0 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:exp_jobs_kernel<.int.64' -s 8 -c 1 \
    -o $O/r02_expjobs64 -f python tools/offline_throughput.py 4096 > $O/r02_ncu_expjobs64.log 2>&1
ls -la $O/r02_nadic32.ncu-rep $O/r02_expjobs64.ncu-rep
